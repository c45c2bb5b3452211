"""Scene flattening: Python scene graph -> SoA host buffers for the C-ABI.

The reference has no device boundary; its callers build ``list[Sphere]`` plus
light lists and hand them to ``Ray.nearestSphereIntersect`` /
``Intersection.terminalRGB`` / ``TraditionalRenderer``.  Here
that object graph is flattened, on every render / ``reset`` (scenes are mutable
lists in the reference, e.g. FB/train_complex_only.py:184-228, so nothing is
cached by identity), into contiguous float64 / int32 arrays which
``rt_scene_create`` (include/rt_b200.h) uploads once per call into HBM.

Objects are read by attribute (duck typing): both this package's classes and the
reference's own ``Sphere`` / ``Material`` / ... instances flatten identically.
"""
from dataclasses import dataclass, field

import numpy as np

__all__ = ["FlatScene", "flatten_scene"]

_F = np.float64
_I = np.int32


def _xyz(v):
    return (float(v.x), float(v.y), float(v.z))


def _rgb(c):
    return (float(c.r), float(c.g), float(c.b))


@dataclass
class FlatScene:
    """Structure-of-arrays scene.  Shapes: n spheres, nG global lights, nP point
    lights, nL Algorithm-B light spheres."""
    centre: np.ndarray          # [n,3] f64
    radius: np.ndarray          # [n]   f64
    material: np.ndarray        # [n,4] f64  reflective, transparent, emitive, refractive_index
    colour: np.ndarray          # [n,3] f64
    ids: np.ndarray             # [n]   i32
    g_vec: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), _F))
    g_col: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), _F))
    g_strength: np.ndarray = field(default_factory=lambda: np.zeros(0, _F))
    g_max_angle: np.ndarray = field(default_factory=lambda: np.zeros(0, _F))
    g_func: np.ndarray = field(default_factory=lambda: np.zeros(0, _I))
    p_id: np.ndarray = field(default_factory=lambda: np.zeros(0, _I))
    p_pos: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), _F))
    p_col: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), _F))
    p_strength: np.ndarray = field(default_factory=lambda: np.zeros(0, _F))
    p_max_angle: np.ndarray = field(default_factory=lambda: np.zeros(0, _F))
    p_func: np.ndarray = field(default_factory=lambda: np.zeros(0, _I))
    bg: np.ndarray = field(default_factory=lambda: np.zeros(3, _F))
    l_centre: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), _F))
    l_colour: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), _F))
    l_index: np.ndarray = field(default_factory=lambda: np.zeros(0, _I))
    small: np.ndarray = None    # [n] u8

    @property
    def n(self):
        return int(self.radius.shape[0])

    def index_of_id(self, sphere_id):
        """First scene index carrying ``sphere_id`` (-1 if none)."""
        hit = np.nonzero(self.ids == int(sphere_id))[0]
        return int(hit[0]) if hit.size else -1


def flatten_scene(spheres, global_light_sources=None, point_light_sources=None, background_colour=None,
                  light_sources=None, small_lights=None, path_lights=True, lights_like=None):
    """Flatten a scene graph.

    spheres               list of Sphere-like (``.centre .radius .material .colour .id``)
    global_light_sources  list of GlobalLight-like   (Algorithm A, RL/ray.py:43-45)
    point_light_sources   list of PointLight-like    (Algorithm A, RL/ray.py:47-62)
    background_colour     Colour-like or None (= black)
    light_sources         Algorithm B: spheres evaluated as lights
                          (``TraditionalRenderer.light_sources``); default = the emissive spheres, in scene
                          order, as the reference's drivers build it (FB/fb_vs_traditional_chandelier.py:801)
    small_lights          Algorithm B: spheres counted in ``small_light_hits``; default radius < 0.5 lights
    lights_like           a FlatScene whose (already flattened) global / point light arrays and background are reused
                          as they are: for callers whose lights are constants (the output5 entries)
    path_lights           False: skip the Algorithm-B light arrays (callers that only run Algorithm A / the env; a
                          sub-millisecond frame should not pay for them on every call)
    """
    spheres = list(spheres)
    n = len(spheres)
    # one pass over the objects, one array construction: this runs on EVERY render / reset (scenes are mutable lists)
    rows = []
    for s in spheres:
        c, m, k = s.centre, s.material, s.colour
        rows.append((c.x, c.y, c.z, s.radius, m.reflective, m.transparent, m.emitive, m.refractive_index, k.r, k.g, k.b))
    a = np.array(rows, _F).reshape(n, 11)
    centre = np.ascontiguousarray(a[:, 0:3])
    radius = np.ascontiguousarray(a[:, 3])
    material = np.ascontiguousarray(a[:, 4:8])
    colour = np.ascontiguousarray(a[:, 8:11])
    ids = np.array([int(s.id) for s in spheres], _I).reshape(n)
    fs = FlatScene(centre, radius, material, colour, ids)

    if lights_like is not None:
        for name in ("g_vec", "g_col", "g_strength", "g_max_angle", "g_func", "p_id", "p_pos", "p_col", "p_strength",
                     "p_max_angle", "p_func", "bg"):
            setattr(fs, name, getattr(lights_like, name))
        global_light_sources = point_light_sources = background_colour = None
    gl = list(global_light_sources or [])
    if gl:
        fs.g_vec = np.array([_xyz(g.vector) for g in gl], _F)
        fs.g_col = np.array([_rgb(g.colour) for g in gl], _F)
        fs.g_strength = np.array([float(g.strength) for g in gl], _F)
        fs.g_max_angle = np.array([float(g.max_angle) for g in gl], _F)
        fs.g_func = np.array([int(g.func) for g in gl], _I)
    pl = list(point_light_sources or [])
    if pl:
        fs.p_id = np.array([int(p.id) for p in pl], _I)
        fs.p_pos = np.array([_xyz(p.position) for p in pl], _F)
        fs.p_col = np.array([_rgb(p.colour) for p in pl], _F)
        fs.p_strength = np.array([float(p.strength) for p in pl], _F)
        fs.p_max_angle = np.array([float(p.max_angle) for p in pl], _F)
        fs.p_func = np.array([int(p.func) for p in pl], _I)
    if background_colour is not None:
        fs.bg = np.array(_rgb(background_colour), _F)

    if not path_lights:
        fs.small = np.zeros(n, np.uint8)
        return fs
    if light_sources is None:
        light_sources = [s for s in spheres if s.material.emitive]
    light_sources = list(light_sources)
    if small_lights is None:
        small_lights = [s for s in light_sources if s.radius < 0.5]
    where = {id(s): i for i, s in reversed(list(enumerate(spheres)))}   # object identity, like ``light == sphere``
    nL = len(light_sources)
    fs.l_centre = np.array([_xyz(s.centre) for s in light_sources], _F).reshape(nL, 3)
    fs.l_colour = np.array([_rgb(s.colour) for s in light_sources], _F).reshape(nL, 3)
    fs.l_index = np.array([where.get(id(s), -1) for s in light_sources], _I).reshape(nL)
    small_ids = {id(s) for s in small_lights}
    fs.small = np.array([1 if id(s) in small_ids else 0 for s in spheres], np.uint8).reshape(n)
    return fs
