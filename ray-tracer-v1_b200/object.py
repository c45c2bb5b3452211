"""``Sphere`` -- the only primitive of the scene API (drop-in for RL/object.py:1-9)."""
from .colour import Colour

__all__ = ["Sphere"]


class Sphere:

    def __init__(self, centre, radius, material, colour=Colour(128, 128, 128), id=0):
        self.id = id
        self.centre = centre
        self.radius = radius
        self.material = material
        self.colour = colour

    def __repr__(self):
        return f"Sphere(id={self.id!r}, centre={self.centre!r}, radius={self.radius!r})"
