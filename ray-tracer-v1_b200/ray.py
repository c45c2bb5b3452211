"""``Ray`` and ``Intersection`` -- drop-ins for RL/ray.py (= FB/ray.py) whose tracing methods run on the GPU.

The scalar methods (``sphereDiscriminant``, ``nearestSphereIntersect``, ``terminalRGB``) keep the reference's
signatures and return types, so notebook-style loops keep working; each call is a batch of ONE through
``rt_sphere_discriminant`` / ``rt_trace_rays`` / ``rt_terminal_rgb`` in the FP64 parity build (the reference computes
in Python floats).  They exist for compatibility -- a per-ray round trip to the device is slow by construction.  The
fast paths are the batched forms: ``trace_rays`` / ``shade_hits`` below, the frame renderers (renderers.py) and
``BatchedRayTracerEnv``.

The scene list is re-flattened on every call (scenes are mutable in the reference); the device copy is reused only
when the flattened bytes are identical.
"""
import numpy as np

from . import _native as nat
from .colour import Colour
from .scene import flatten_scene
from .vector import Vector

__all__ = ["Ray", "Intersection", "trace_rays", "shade_hits"]

_cache = {"key": None, "scene": None, "precision_device": None}


def _device_scene(spheres, global_light_sources=(), point_light_sources=(), background_colour=None, device=0):
    fs = flatten_scene(spheres, global_light_sources, point_light_sources, background_colour)
    key = b"".join(np.ascontiguousarray(getattr(fs, k)).tobytes() for k in (
        "centre", "radius", "material", "colour", "ids", "g_vec", "g_col", "g_strength", "g_max_angle", "g_func", "p_id",
        "p_pos", "p_col", "p_strength", "p_max_angle", "p_func", "bg")) + bytes([device])
    if _cache["key"] != key:
        if _cache["scene"] is not None:
            _cache["scene"].close()
        _cache["scene"] = nat.DeviceScene(fs, device)
        _cache["key"] = key
    return _cache["scene"], fs


def trace_rays(spheres, origins, directions, suppress_ids=None, bounces=None, through_counts=None, max_bounces=1,
               precision="f64", device=0):
    """Batched ``Ray(o, d).nearestSphereIntersect(spheres, ...)``: origins/directions [m,3] ->
    dict(hit [m] bool, index [m] (position in ``spheres``), bounces, through_count, point [m,3], normal [m,3], distance [m])."""
    sc, _ = _device_scene(spheres, device=device)
    rays = np.concatenate([np.asarray(origins, np.float64).reshape(-1, 3), np.asarray(directions, np.float64).reshape(-1, 3)], 1)
    prec = nat.F64 if precision in ("f64", nat.F64) and precision != nat.F32 else nat.F32
    term, _ = sc.trace_rays(rays, suppress=suppress_ids, bounces0=bounces, through0=through_counts, max_bounces=max_bounces,
                            shade=False, precision=prec)
    return {"hit": term[:, 0] == 1, "index": term[:, 1].astype(np.int64), "bounces": term[:, 2].astype(np.int64),
            "through_count": term[:, 3].astype(np.int64), "point": term[:, 4:7], "normal": term[:, 7:10],
            "distance": term[:, 10]}


def shade_hits(spheres, indices, points, normals, background_colour=Colour(0, 0, 0), global_light_sources=(),
               point_light_sources=(), max_bounces=0, precision="f64", device=0):
    """Batched ``Intersection.terminalRGB``: -> rgb [m,3] (0-255 scale floats, not clamped)."""
    sc, _ = _device_scene(spheres, global_light_sources, point_light_sources, background_colour, device)
    hits = np.concatenate([np.asarray(indices, np.float64).reshape(-1, 1), np.asarray(points, np.float64).reshape(-1, 3),
                           np.asarray(normals, np.float64).reshape(-1, 3)], 1)
    prec = nat.F64 if precision in ("f64", nat.F64) and precision != nat.F32 else nat.F32
    return sc.shade_hits(hits, max_bounces, prec)


class Intersection:
    """RL/ray.py:8-65."""

    @staticmethod
    def nearestIntersection(intersections):
        nearest = None
        for intersection in intersections:
            if intersection.intersects == True:      # noqa: E712  (the reference's comparison)
                if nearest is None or intersection.distance < nearest.distance:
                    nearest = intersection
        return nearest

    def __init__(self, intersects=False, distance=None, point=None, normal=None, object=None, bounces=0, through_count=0):
        self.intersects = intersects
        self.distance = distance
        self.point = point
        self.normal = normal
        self.object = object
        self.bounces = bounces
        self.through_count = through_count

    def terminalRGB(self, spheres, background_colour=Colour(0, 0, 0), global_light_sources=[], point_light_sources=[],
                    max_bounces=0):
        """Colour of the surface this intersection landed on (RL/ray.py:37-65), evaluated on the GPU."""
        spheres = list(spheres)
        index = next((i for i, s in enumerate(spheres) if s is self.object), None)
        if index is None:
            # the reference only reads self.object's own fields plus its id: shade it as an extra sphere at the end
            spheres = spheres + [self.object]
            index = len(spheres) - 1
        rgb = shade_hits(spheres, [index], [self.point.getXYZ()], [self.normal.getXYZ()], background_colour,
                         global_light_sources, point_light_sources, max_bounces)[0]
        return Colour(float(rgb[0]), float(rgb[1]), float(rgb[2]))


class Ray:
    """RL/ray.py:68-231."""

    def __init__(self, origin, D):
        self.origin = origin
        self.D = D.normalise()

    def sphereDiscriminant(self, sphere, point=0):
        out = nat.sphere_discriminant([[*self.origin.getXYZ(), *self.D.getXYZ()]],
                                      [[*sphere.centre.getXYZ(), float(sphere.radius)]], point, nat.F64)[0]
        if out[0] != 1:
            return Intersection()
        return Intersection(intersects=True, distance=float(out[1]), point=Vector(*out[2:5]), normal=Vector(*out[5:8]),
                            object=sphere)

    def nearestSphereIntersect(self, spheres, suppress_ids=[], bounces=0, max_bounces=1, through_count=0):
        spheres = list(spheres)
        if len(suppress_ids) > 1:
            raise NotImplementedError("the reference only ever suppresses one id (the sphere the ray leaves)")
        sup = None if not suppress_ids else [int(suppress_ids[0])]
        r = trace_rays(spheres, [self.origin.getXYZ()], [self.D.getXYZ()], suppress_ids=sup, bounces=[int(bounces)],
                       through_counts=[int(through_count)], max_bounces=max_bounces)
        if not r["hit"][0]:
            return None
        return Intersection(intersects=True, distance=float(r["distance"][0]), point=Vector(*r["point"][0]),
                            normal=Vector(*r["normal"][0]), object=spheres[int(r["index"][0])],
                            bounces=int(r["bounces"][0]), through_count=int(r["through_count"][0]))
