"""Light descriptions of the scene API (drop-in for RL/light.py:1-37).

The GPU evaluates these rules inside ``terminalRGB`` (csrc/rt_trace.cuh,
``terminal_rgb``); the Python methods below exist so user code that calls them
directly keeps working.
"""
from .colour import Colour  # noqa: F401  (re-exported like the reference module)

__all__ = ["incidence", "GlobalLight", "PointLight"]


def incidence(angle, max_angle):
    """Linear fall-off of a light with the angle to the surface normal (RL/light.py:3-9)."""
    if angle > max_angle:
        return 0
    return 1 if angle == 0 else (max_angle - angle) / max_angle


class GlobalLight:
    """Directional light: ``vector`` points towards the light (RL/light.py:11-21)."""

    def __init__(self, vector, colour, strength, max_angle, func=0):
        self.vector, self.colour, self.strength = vector, colour, strength
        self.max_angle, self.func = max_angle, func

    def relativeStrength(self, angle):
        if self.func == 0:
            return self.colour.scaleRGB(incidence(angle, self.max_angle) * self.strength)


class PointLight:
    """Point light tied to the sphere with the same ``id`` (RL/light.py:24-37).

    ``func=-1``: no distance fall-off; ``func=0``: divided by distance."""

    def __init__(self, id, position, colour, strength, max_angle, func=0):
        self.id, self.position, self.colour = id, position, colour
        self.strength, self.max_angle, self.func = strength, max_angle, func

    def relativeStrength(self, angle, distance):
        k = incidence(angle, self.max_angle) * self.strength
        if self.func == -1:
            return self.colour.scaleRGB(k)
        if self.func == 0:
            return self.colour.scaleRGB(k / distance)
