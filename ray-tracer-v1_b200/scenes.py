"""Scene builders for the benchmark / parity configurations (BASELINE.json configs).

Each ``build_*`` function takes a *namespace* ``ns`` providing the scene classes
(``Vector Colour Material Sphere GlobalLight PointLight``) and defaults to this
package's own.  oracle/gen_golden.py passes the reference's classes instead, so
the very same table of numbers drives both the reference run that produced
tests/golden/ and the CUDA runs that are compared with it.

Scene data (positions, radii, colours) restates the reference's scene
definitions -- it is input data of the hot path, cited per builder.  The
"complex" scene's source module is absent from the reference (SURVEY.md
section 8c), so ``build_complex`` is a seeded synthetic scene matching its known
census.
"""
import math
import random
from types import SimpleNamespace

import numpy as np

from . import colour as _colour, light as _light, material as _material, object as _object, vector as _vector
from .scene import FlatScene

__all__ = ["SceneSpec", "default_ns", "build_balls_in_space", "build_marbles4", "build_planets2", "build_chandelier",
           "build_complex", "build_optimized_env_scene", "build_many_spheres_flat", "notebook_grid", "custom_scene_grid"]

default_ns = SimpleNamespace(Vector=_vector.Vector, Angle=_vector.Angle, Colour=_colour.Colour,
                             Material=_material.Material, Sphere=_object.Sphere,
                             GlobalLight=_light.GlobalLight, PointLight=_light.PointLight)


class SceneSpec(SimpleNamespace):
    """spheres, global_lights, point_lights, background, camera (x,y,z) + free-form extras."""


def notebook_grid(ray_count, ray_step):
    """Ray grid of the notebooks and ``render_true_original`` (RL/output5.py:432-433): 2*ray_count+1 per axis."""
    xs = [r * ray_step for r in range(-ray_count, 0, 1)] + [r * ray_step for r in range(0, ray_count + 1)]
    ys = [r * ray_step for r in range(ray_count, 0, -1)] + [-r * ray_step for r in range(0, ray_count + 1)]
    return np.array(xs, np.float64), np.array(ys, np.float64)


def custom_scene_grid(width, height):
    """Ray grid of ``render_custom_scene`` (RL/output5.py:1437-1450): +-k on BOTH axes, k = int(100*min(W,H)/601)*0.01."""
    k = int(100 * (min(width, height) / 601)) * 0.01
    return np.linspace(-k, k, width), np.linspace(k, -k, height)


# --------------------------------------------------------------------------- C1
def build_balls_in_space(ns=default_ns, as_rendered=True):
    """"balls_in_space" = ``create_custom_scene()`` (RL/output5.py:165-262; FB/output6.py:45-83).

    as_rendered=True applies what ``render_true_original`` / ``_trace_custom_traditional`` do before tracing
    (RL/output5.py:447-486, :543-578): the sun (id 7) is replaced by an identical sphere with id 0 appended LAST,
    one GlobalLight and one PointLight(func=-1) tied to id 0, background (2,2,5), camera (0,0,1).
    as_rendered=False is the raw 7-sphere list (sun id 7) the FB-flavour env steps in."""
    V, C, M, S = ns.Vector, ns.Colour, ns.Material, ns.Sphere
    base, mirror = M(reflective=False), M(reflective=True)
    glass = M(reflective=False, transparent=True, refractive_index=1.52)
    table = [
        (1, (-0.8, 0.6, 0), 0.3, glass, (255, 100, 100)),
        (2, (0.8, -0.8, -10), 2.2, base, (204, 204, 255)),
        (3, (0.3, 0.34, 0.1), 0.2, base, (0, 51, 204)),
        (4, (5.6, 3, -2), 5, mirror, (153, 51, 153)),
        (5, (-0.8, -0.8, -0.2), 0.25, base, (153, 204, 0)),
        (6, (-3, 10, -75), 30, base, (255, 204, 102)),
    ]
    spheres = [S(id=i, centre=V(*c), radius=r, material=m, colour=C(*col)) for i, c, r, m, col in table]
    sun = S(id=0 if as_rendered else 7, centre=V(-0.6, 0.2, 6), radius=0.1, material=M(emitive=True),
            colour=C(255, 255, 204))
    spheres.append(sun)
    gl = [ns.GlobalLight(vector=V(3, 1, -0.75), colour=C(20, 20, 255), strength=1, max_angle=np.radians(90), func=0)]
    pl = [ns.PointLight(id=sun.id, position=sun.centre, colour=sun.colour, strength=1, max_angle=np.radians(90),
                        func=-1)]
    return SceneSpec(spheres=spheres, global_lights=gl, point_lights=pl, background=C(2, 2, 5), miss=C(2, 2, 5),
                     camera=(0.0, 0.0, 1.0), sun=sun)


# --------------------------------------------------------------------------- C2
def build_marbles4(ns=default_ns):
    """"marbles" scene: RL/Marbles 4.ipynb cell 0 (8 spheres, 3 PointLight func=0, glass n=2, 2 mirrors)."""
    V, C, M, S = ns.Vector, ns.Colour, ns.Material, ns.Sphere
    base, emit, mirror = M(), M(emitive=True), M(reflective=True)
    glass = M(reflective=False, transparent=True, refractive_index=2)
    lights_tbl = [(200, (5, 0.5, 1.5), (179, 230, 255), 5), (201, (-5, 0.5, 2.5), (255, 153, 194), 5),
                  (202, (1, 1, 6), (255, 218, 179), 3)]
    spheres, pl = [], []
    for i, c, col, strength in lights_tbl:
        s = S(id=i, centre=V(*c), radius=0.05, material=emit, colour=C(*col))
        spheres.append(s)
        pl.append(ns.PointLight(id=s.id, position=s.centre, colour=s.colour, strength=strength,
                                max_angle=np.radians(90), func=0))
    for c, r, m, col in [((0, 0, 2), 0.5, glass, (100, 100, 100)), ((0.5, 0.5, -1), 1, base, (153, 102, 255)),
                         ((-0.5, -0.5, 1), 0.3, base, (204, 51, 0)), ((0.5, 0.3, 4), 0.3, mirror, (194, 194, 214)),
                         ((-1, -0.1, -6), 1.5, mirror, (255, 214, 153))]:
        spheres.append(S(id=len(spheres), centre=V(*c), radius=r, material=m, colour=C(*col)))
    gl = [ns.GlobalLight(vector=V(0.1, 1, -0.2), colour=C(255, 255, 255), strength=0.5, max_angle=np.radians(180),
                         func=0)]
    return SceneSpec(spheres=spheres, global_lights=gl, point_lights=pl, background=C(0, 0, 1), miss=C(230, 230, 255),
                     camera=(0.0, 0.0, 10.0), ray_step=0.002)


def build_planets2(ns=default_ns):
    """"shadows" scene: RL/Planets 2.ipynb cell 0 (10 spheres, 3 PointLight func=0 strengths 3/1/2)."""
    V, C, M, S = ns.Vector, ns.Colour, ns.Material, ns.Sphere
    base, emit, mirror = M(), M(emitive=True), M(reflective=True)
    glass = M(reflective=False, transparent=True, refractive_index=1.52)
    tbl = [
        (0, (0.2, 0, 0), 0.1, emit, (255, 255, 204)),
        (1, (-1, 0.5, -2), 1, base, (255, 153, 102)),
        (2, (1, -0.5, 0.5), 0.4, base, (255, 0, 0)),
        (3, (-10, 5, -20), 14, base, (102, 204, 255)),
        (4, (0, 0.4, -0.8), 0.2, base, (204, 0, 204)),
        (5, (0.45, -0.25, 0.2), 0.1, base, (50, 255, 25)),
        (6, (1.5, 1, -2.5), 1, mirror, (24, 24, 35)),
        (7, (-5, -5, 5), 0.2, emit, (255, 0, 0)),
        (8, (5, 0, -2.5), 0.2, emit, (0, 255, 0)),
        (10, (-0.25, -0.2, 0.7), 0.3, glass, (100, 100, 100)),
    ]
    spheres = [S(id=i, centre=V(*c), radius=r, material=m, colour=C(*col)) for i, c, r, m, col in tbl]
    by_id = {s.id: s for s in spheres}
    pl = [ns.PointLight(id=i, position=by_id[i].centre, colour=by_id[i].colour, strength=k, max_angle=np.radians(90),
                        func=0) for i, k in [(0, 3), (7, 1), (8, 2)]]
    gl = [ns.GlobalLight(vector=V(1, 0.1, -0.2), colour=C(0, 0, 255), strength=0.1, max_angle=np.radians(90), func=0)]
    return SceneSpec(spheres=spheres, global_lights=gl, point_lights=pl, background=C(0, 0, 1), miss=C(0, 0, 1),
                     camera=(0.0, 0.0, 5.0), ray_step=0.005)


# --------------------------------------------------------------------------- C4
def build_chandelier(ns=default_ns):
    """``generate_chandelier_scene()`` (FB/fb_vs_traditional_chandelier.py:275-387): 29 spheres, 21 emissive.
    Camera (0,2,0) (:807); Algorithm B with mirror rule ``reflective > 0`` (:481)."""
    V, C, M, S = ns.Vector, ns.Colour, ns.Material, ns.Sphere
    matte_white = M(reflective=0.1, transparent=0, emitive=0)
    mirror = M(reflective=0.95, transparent=0, emitive=0)
    glass = M(reflective=0.1, transparent=0.9, emitive=0, refractive_index=1.5)
    emit = M(reflective=0, transparent=0, emitive=1)
    base = 1000
    tbl = [(1, (0, -100, 0), 99, mirror, (220, 220, 230)), (2, (0, 100, 0), 99, mirror, (240, 240, 255)),
           (3, (0, 0, -100), 99, matte_white, (210, 210, 230)), (4, (-100, 0, 0), 99, matte_white, (200, 200, 220)),
           (5, (100, 0, 0), 99, matte_white, (220, 200, 200)), (6, (0, 10, 5), 1.2, emit, (255, 255, 240))]
    spheres = [S(id=base + i, centre=V(*c), radius=r, material=m, colour=C(*col)) for i, c, r, m, col in tbl]
    cx, cy, cz, rad = 0, 4, 8, 2.0
    for i in range(20):
        theta = (i * 137.5) % 360 * math.pi / 180
        phi = (i * 90) % 360 * math.pi / 180
        pos = V(cx + rad * math.sin(phi) * math.cos(theta), cy + rad * math.sin(phi) * math.sin(theta),
                cz + rad * math.cos(phi))
        rgb = [int(200 + 55 * math.sin(theta)), int(200 + 55 * math.cos(phi)), int(200 + 55 * math.sin(phi + theta))]
        rgb = [max(180, min(255, c)) for c in rgb]
        spheres.append(S(id=base + 10 + i, centre=pos, radius=0.1, material=emit, colour=C(*rgb)))
    for i, c, r, m, col in [(40, (1.5, 3, 7), 0.6, glass, (255, 255, 255)),
                            (41, (-1.5, -1.2, 6), 0.7, mirror, (200, 200, 220)),
                            (42, (0, 1, 4), 0.5, glass, (255, 240, 240))]:
        spheres.append(S(id=base + i, centre=V(*c), radius=r, material=m, colour=C(*col)))
    return SceneSpec(spheres=spheres, global_lights=[], point_lights=[], background=C(2, 2, 5),
                     camera=(0.0, 2.0, 0.0), mirror_threshold=0.0, max_bounces=8)


# --------------------------------------------------------------------------- C3
def build_complex(ns=default_ns, seed=0):
    """Synthetic restatement of the missing ``complex_scene.create_complex_scene()``.

    Census known from the reference's artefacts (traditional_renders/*_stats.txt:14-22,
    complex_scene_layout.png legend, FB/train_complex_only.py:196): 54 spheres, 3 lights (one big,
    two small r 0.08-0.15), six r=99 wall spheres ids 1-6, one big mirror, glass spheres, matte rows and
    an interlocked cluster; camera (0,0,12), fov 60; Algorithm B, mirror rule ``reflective > 0.9``
    (FB/fb_vs_traditional_complex.py:349), max_bounces 5.  Matte colours ~ U[100,255] from ``seed``."""
    V, C, M, S = ns.Vector, ns.Colour, ns.Material, ns.Sphere
    rng = random.Random(seed)
    matte = M(reflective=0, transparent=0, emitive=0)
    wall = M(reflective=0.05, transparent=0, emitive=0)
    mirror = M(reflective=0.95, transparent=0, emitive=0)
    glass = M(reflective=0.1, transparent=0.9, emitive=0, refractive_index=1.5)
    emit = M(reflective=0, transparent=0, emitive=1)

    def colr():
        return C(*(rng.randint(100, 255) for _ in range(3)))

    spheres = []
    # room x in [-5,5], y in [-1,6], z in [-2,14] bounded by six r=1000 spheres; camera and lights lie
    # OUTSIDE every wall sphere (with the chandelier's r=99 @ +-100 walls they would be inside the ceiling/front
    # spheres, every path would die at the depth limit with zero light hits, contradicting the reference's own
    # stats file: 5.79 rays/sample, 0.96 % light hits)
    R = 1000.0
    walls = [((0, -1 - R, 0), (200, 200, 210)), ((0, 6 + R, 0), (230, 230, 240)), ((0, 0, -2 - R), (210, 210, 230)),
             ((-5 - R, 0, 0), (220, 180, 180)), ((5 + R, 0, 0), (180, 220, 180)), ((0, 0, 14 + R), (200, 200, 200))]
    for i, (c, col) in enumerate(walls):
        spheres.append(S(id=i + 1, centre=V(*c), radius=R, material=wall, colour=C(*col)))
    nid = 7
    for c, r, col in [((0, 4, 4), 1.0, (255, 250, 235)), ((-1, 2, 7), 0.15, (255, 220, 180)),
                      ((3, 1.5, 4), 0.08, (200, 220, 255))]:
        spheres.append(S(id=nid, centre=V(*c), radius=r, material=emit, colour=C(*col))); nid += 1
    spheres.append(S(id=nid, centre=V(-2.5, -1, 3), radius=2.0, material=mirror, colour=C(235, 235, 245))); nid += 1
    spheres.append(S(id=nid, centre=V(2.5, -1, 3), radius=1.5, material=glass, colour=C(255, 255, 255))); nid += 1
    for k in range(4):
        spheres.append(S(id=nid, centre=V(-1.5 + k, -0.6, 2), radius=0.3, material=glass, colour=C(240, 250, 255))); nid += 1
    for z in (8, 9.5):
        for k in range(5):
            spheres.append(S(id=nid, centre=V(-2 + k, -0.6, z), radius=0.4, material=matte, colour=colr())); nid += 1
    for k in range(6):   # interlocked cluster: neighbours overlap (centre spacing < 2r)
        a = k * math.pi / 3
        spheres.append(S(id=nid, centre=V(0.45 * math.cos(a), 0.5 + 0.45 * math.sin(a), 5 + 0.1 * (k % 2)),
                         radius=0.35, material=matte, colour=colr())); nid += 1
    while len(spheres) < 54:
        spheres.append(S(id=nid, centre=V(rng.uniform(-3.5, 3.5), rng.uniform(-0.8, 3.0), rng.uniform(1.5, 10.0)),
                         radius=rng.uniform(0.12, 0.3), material=matte, colour=colr())); nid += 1
    return SceneSpec(spheres=spheres, global_lights=[], point_lights=[], background=C(2, 2, 5),
                     camera=(0.0, 0.0, 12.0), mirror_threshold=0.9, max_bounces=5)


# --------------------------------------------------------------------------- C5
def build_optimized_env_scene(ns=default_ns):
    """``create_optimized_scene()`` (RL/train_raytracer_improved.py:52-93): the RL-flavour env scene
    (320x240, fov 80, max_bounces 6, camera origin, black background)."""
    V, C, M, S = ns.Vector, ns.Colour, ns.Material, ns.Sphere
    matte = M(reflective=0, transparent=0, emitive=0.1, refractive_index=1)
    mirror = M(reflective=1, transparent=0, emitive=0, refractive_index=1)
    lamp = M(reflective=0, transparent=0, emitive=1, refractive_index=1)
    spheres = [S(V(0, -100, -3), 99, matte, C(100, 100, 100), id=1), S(V(0, 0, -3), 0.7, mirror, C(255, 255, 255), id=2),
               S(V(-1.8, 0.3, -3), 0.5, mirror, C(200, 200, 255), id=3), S(V(0, 2, -3), 0.5, lamp, C(255, 255, 200), id=99),
               S(V(-2, 1.5, -3), 0.4, lamp, C(200, 255, 200), id=100)]
    pl = [ns.PointLight(id=99, position=V(0, 2, -3), colour=C(255, 255, 200), strength=12.0, max_angle=np.pi, func=0),
          ns.PointLight(id=100, position=V(-2, 1.5, -3), colour=C(200, 255, 200), strength=8.0, max_angle=np.pi, func=0)]
    return SceneSpec(spheres=spheres, global_lights=[], point_lights=pl, background=C(0, 0, 0),
                     camera=(0.0, 0.0, 0.0), width=320, height=240, fov=80, max_bounces=6)


# ------------------------------------------------------------- C4 scaled (LBVH)
def build_many_spheres_flat(n_small, seed=0, emissive_fraction=0.01):
    """Chandelier room + ``n_small`` spheres r~U(0.02,0.1) uniform in [-4,4]x[-1,6]x[2,10], 1% emissive
    (``emissive_fraction``)
    (SURVEY.md section 8d, C4 scaled variant for the on-device LBVH).  Returns a FlatScene directly (no
    per-sphere Python objects: n_small reaches 1e5)."""
    from .scene import flatten_scene
    room = build_chandelier().spheres[:6]
    fs0 = flatten_scene(room)
    rs = np.random.RandomState(seed)
    lo, hi = np.array([-4.0, -1.0, 2.0]), np.array([4.0, 6.0, 10.0])
    centre = lo + (hi - lo) * rs.random_sample((n_small, 3))
    radius = 0.02 + 0.08 * rs.random_sample(n_small)
    emissive = rs.random_sample(n_small) < emissive_fraction
    mirror = (~emissive) & (rs.random_sample(n_small) < 0.05)
    colour = np.floor(100 + 156 * rs.random_sample((n_small, 3)))
    material = np.zeros((n_small, 4))
    material[:, 0] = np.where(mirror, 0.95, 0.0)
    material[:, 2] = emissive.astype(np.float64)
    material[:, 3] = 1.0
    fs = FlatScene(centre=np.concatenate([fs0.centre, centre]), radius=np.concatenate([fs0.radius, radius]),
                   material=np.concatenate([fs0.material, material]), colour=np.concatenate([fs0.colour, colour]),
                   ids=np.concatenate([fs0.ids, 2000 + np.arange(n_small)]).astype(np.int32))
    lights = np.nonzero(fs.material[:, 2] != 0)[0].astype(np.int32)
    fs.l_index, fs.l_centre, fs.l_colour = lights, fs.centre[lights].copy(), fs.colour[lights].copy()
    fs.small = ((fs.material[:, 2] != 0) & (fs.radius < 0.5)).astype(np.uint8)
    fs.bg = np.array([2.0, 2.0, 5.0])
    return fs
