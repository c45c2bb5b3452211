"""``Material`` -- four scalars per sphere (drop-in for RL/material.py:1-23).

``reflective`` / ``transparent`` / ``emitive`` are used both as booleans
(``Material(reflective=True)``) and as 0-1 floats (0.95, 0.1) by the reference's
scenes; the tracer's rules (``== True`` for Algorithm A, ``> threshold`` for
Algorithm B, truthiness for ``emitive``) are applied on the GPU to the float
value ``float(x)``.
"""

__all__ = ["Material", "matte"]


class Material:

    def __init__(self, reflective=0, transparent=0, emitive=0, refractive_index=1):
        self.reflective = reflective
        self.transparent = transparent
        self.emitive = emitive
        self.refractive_index = refractive_index

    def __repr__(self):
        return (f"Material(reflective={self.reflective!r}, transparent={self.transparent!r}, "
                f"emitive={self.emitive!r}, refractive_index={self.refractive_index!r})")


matte = Material()
