"""``Vector`` / ``Angle`` value types of the scene API.

Drop-in for the reference's ``vector.py`` (RL/vector.py:5-139): same class
names, constructor signatures, public ``x/y/z`` attributes and method names, so
scene-building code written against the reference runs unchanged.  These are
host-side scene-description helpers; the tracing arithmetic itself runs on the
GPU (csrc/rt_trace.cuh) and never calls back into this module.
"""
import math

import numpy as np

__all__ = ["Vector", "Angle"]


class Vector:
    __slots__ = ("x", "y", "z")

    def __init__(self, x, y, z):
        self.x, self.y, self.z = x, y, z

    # -- construction / conversion (RL/vector.py:7-23)
    @staticmethod
    def fromNpArray(array):
        return Vector(array[0], array[1], array[2])

    def toNpArray(self):
        return np.array(self.getXYZ())

    def getXYZ(self):
        return self.x, self.y, self.z

    def describe(self, caption=""):
        print(f"{caption}x: {self.x}, y: {self.y}, z: {self.z}")

    def __repr__(self):
        return f"Vector({self.x!r}, {self.y!r}, {self.z!r})"

    def __iter__(self):
        return iter(self.getXYZ())

    def _assign(self, x, y, z, inplace):
        if inplace == True:  # noqa: E712  (reference semantics: only the literal True mutates)
            self.x, self.y, self.z = x, y, z
            return self
        return Vector(x, y, z)

    # -- arithmetic (RL/vector.py:25-56)
    def addVector(self, B, inplace=False):
        return self._assign(self.x + B.x, self.y + B.y, self.z + B.z, inplace)

    def subtractVector(self, B, inplace=False):
        return self._assign(self.x - B.x, self.y - B.y, self.z - B.z, inplace)

    def invert(self, inplace=False):
        return self._assign(-self.x, -self.y, -self.z, inplace)

    def scaleByLength(self, l, inplace=False):
        return self._assign(self.x * l, self.y * l, self.z * l, inplace)

    # -- metrics (RL/vector.py:58-62, 94-112)
    def dotProduct(self, B):
        return self.x * B.x + self.y * B.y + self.z * B.z

    def crossProduct(self, B):
        ax, ay, az = self.getXYZ()
        bx, by, bz = B.x, B.y, B.z
        return Vector(ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx)

    def magnitude(self):
        return math.sqrt(self.dotProduct(self))

    def normalise(self):
        m = self.magnitude()
        return Vector(self.x / m, self.y / m, self.z / m)

    def distanceFrom(self, B):
        return B.subtractVector(self).magnitude()

    def angleBetween(self, B):
        return np.arccos(self.dotProduct(B) / (self.magnitude() * B.magnitude()))

    # -- optics (RL/vector.py:64-92)
    def reflectInVector(self, B):
        v, n = self.normalise(), B.normalise()
        return v.subtractVector(n.scaleByLength(2 * v.dotProduct(n))).normalise()

    def refractInVector(self, B, r_index_a, r_index_b):
        """Snell refraction; ``False`` on total internal reflection."""
        v, n = self.normalise(), B.normalise()
        eta = r_index_a / r_index_b
        cos_i = abs(min(1, max(-1, v.dotProduct(n))))
        k = 1 - eta ** 2 * (1 - cos_i ** 2)
        if k < 0:
            return False
        return v.scaleByLength(eta).addVector(n.scaleByLength(eta * cos_i - math.sqrt(k))).normalise()

    # -- camera rotation (RL/vector.py:114-127)
    def multiplyByMatrix(self, T):
        return Vector.fromNpArray(np.matmul(self.toNpArray(), T))

    def rotate(self, angle, inplace=False):
        ca, cb, cc = np.cos(angle.x), np.cos(angle.y), np.cos(angle.z)
        sa, sb, sc = np.sin(angle.x), np.sin(angle.y), np.sin(angle.z)
        R = np.array([
            [cc * cb * ca - sc * sa, cc * cb * sa + sc * ca, -cc * sb],
            [-sc * cb * ca - cc * sa, -sc * cb * sa + cc * ca, sc * sb],
            [sb * ca, sb * sa, cb],
        ])
        # like the reference, ``inplace`` never mutates (it rebinds a local): a new Vector is returned
        return self.multiplyByMatrix(R)


class Angle:
    """Euler camera angles (RL/vector.py:131-139)."""
    __slots__ = ("x", "y", "z")

    def __init__(self, x, y, z):
        self.x, self.y, self.z = x, y, z

    def __repr__(self):
        return f"Angle({self.x!r}, {self.y!r}, {self.z!r})"
