"""``Colour`` -- RGB triple on the 0-255 float scale (drop-in for RL/colour.py:1-29)."""

__all__ = ["Colour"]


class Colour:
    __slots__ = ("r", "g", "b")

    def __init__(self, r, g, b):
        self.r, self.g, self.b = r, g, b

    def __repr__(self):
        return f"Colour({self.r!r}, {self.g!r}, {self.b!r})"

    def getList(self):
        return [self.r, self.g, self.b]

    def addColour(self, colour):
        return Colour(self.r + colour.r, self.g + colour.g, self.b + colour.b)

    def scaleRGB(self, scale, return_type=None):
        scaled = [c * scale for c in self.getList()]
        if return_type is None:
            return Colour(*scaled)
        if return_type == "list":
            return [round(c) for c in scaled]
        if return_type == "Colour":
            return Colour(*(round(c) for c in scaled))

    def illuminate(self, light):
        """Surface colour under ``light``: per channel round(c * l/255), half-to-even (RL/colour.py:21-29)."""
        return Colour(*(round(c * (l / 255)) for c, l in zip(self.getList(), light.getList())))
