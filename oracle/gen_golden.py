#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference Python code.

Runs only in the build container (needs /root/reference, which does not exist on the
GPU box).  The reference modules are imported from where they lie -- nothing is
copied -- with permissive stubs for the packages the container lacks
(matplotlib, seaborn, gymnasium, stable_baselines3) and for the reference's own
missing ``complex_scene`` module.  Randomness: the reference draws from the global
``np.random.random``; here it is monkey-patched with the Philox stream the CUDA
path uses (oracle/rt_oracle.c ``rng_pair``), keyed by (pixel, sample, slot), so the
stochastic renders become deterministic and comparable sample for sample.

    python oracle/gen_golden.py            # writes tests/golden/*.npz (about a minute)
"""
import contextlib
import importlib
import importlib.util
import io
import os
import sys
import tempfile
import types
from pathlib import Path
from unittest import mock

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("RT_REFERENCE", "/root/reference"))
OUT = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))

from oracle import oracle as orc  # noqa: E402  (only for the Philox stream)
import ray_tracer_v1_b200 as rtb  # noqa: E402
from ray_tracer_v1_b200 import scenes  # noqa: E402


# ----------------------------------------------------------------- reference import
def _install_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "seaborn", "stable_baselines3",
                 "stable_baselines3.common", "stable_baselines3.common.callbacks", "stable_baselines3.common.env_util",
                 "stable_baselines3.common.vec_env", "stable_baselines3.common.monitor",
                 "stable_baselines3.common.evaluation", "stable_baselines3.common.env_checker"):
        if name not in sys.modules:
            m = mock.MagicMock(name=name)
            m.__path__ = []
            sys.modules[name] = m
    gym = types.ModuleType("gymnasium")

    class Env:
        def reset(self, seed=None, options=None):
            return None

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.dtype = np.asarray(low, dtype), np.asarray(high, dtype), dtype
            self.shape = self.low.shape

    gym.Env = Env
    gym.spaces = types.ModuleType("gymnasium.spaces")
    gym.spaces.Box = Box
    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.spaces"] = gym.spaces


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


def load_reference():
    _install_stubs()
    sys.path.insert(0, str(REF / "RL"))
    ref = types.SimpleNamespace()
    for m in ("vector", "colour", "material", "object", "light", "ray"):
        setattr(ref, m, importlib.import_module(m))
    ref.ns = types.SimpleNamespace(Vector=ref.vector.Vector, Angle=ref.vector.Angle, Colour=ref.colour.Colour,
                                   Material=ref.material.Material, Sphere=ref.object.Sphere,
                                   GlobalLight=ref.light.GlobalLight, PointLight=ref.light.PointLight)
    ref.env_rl = _load("ref_env_rl", REF / "RL" / "ray_tracer_env.py")
    ref.env_fb = _load("ref_env_fb", REF / "FB" / "ray_tracer_env.py")
    ref.output5 = _load("ref_output5", REF / "RL" / "output5.py")
    ref.chandelier = _load("ref_chandelier", REF / "FB" / "fb_vs_traditional_chandelier.py")
    cs = types.ModuleType("complex_scene")       # the reference's own missing module
    cs.create_complex_scene = lambda: scenes.build_complex(ref.ns).spheres
    cs.create_camera_for_scene = lambda: (ref.ns.Vector(0, 0, 12), None)
    cs.create_lights_for_scene = lambda: ([], [])
    sys.modules["complex_scene"] = cs
    ref.complex = _load("ref_complex", REF / "FB" / "fb_vs_traditional_complex.py")
    ref.improved = _load("ref_improved", REF / "RL" / "train_raytracer_improved.py")
    sys.modules["ray_tracer_env"] = ref.env_rl          # train_raytracer_optimized.py: `from ray_tracer_env import RayTracerEnv`
    ref.optimized = _load("ref_optimized", REF / "RL" / "train_raytracer_optimized.py")
    return ref


def flat(spec, **kw):
    return rtb.flatten_scene(spec.spheres, spec.global_lights, spec.point_lights, spec.background, **kw)


def flat_dict(fs, prefix="scene_"):
    return {prefix + k: getattr(fs, k) for k in ("centre", "radius", "material", "colour", "ids", "g_vec", "g_col",
            "g_strength", "g_max_angle", "g_func", "p_id", "p_pos", "p_col", "p_strength", "p_max_angle", "p_func",
            "bg", "l_centre", "l_colour", "l_index", "small")}


def assert_same_scene(a, b, what):
    fa, fb_ = rtb.flatten_scene(a), rtb.flatten_scene(b)
    for k in ("centre", "radius", "material", "colour", "ids"):
        assert np.array_equal(getattr(fa, k), getattr(fb_, k)), f"{what}: builder differs from the reference in {k}"


# ----------------------------------------------------------------- unit KATs
def gen_kat(ref):
    V, S, M, Ray = ref.ns.Vector, ref.ns.Sphere, ref.ns.Material, ref.ray.Ray
    rs = np.random.RandomState(1234)
    rows = []
    for i in range(400):
        o = rs.uniform(-3, 3, 3)
        c = rs.uniform(-3, 3, 3)
        r = float(rs.uniform(0.2, 2.5))
        if i % 5 == 0:      # origin inside the sphere -> negative t0 (ray.py:93-96)
            o = c + rs.uniform(-0.5, 0.5, 3) * r
        d = (c - o) + rs.uniform(-1, 1, 3) * r * (1.5 if i % 3 else 0.3)
        if i % 7 == 0:
            d = -d          # behind the origin -> tca < 0
        point = i % 2
        it = Ray(V(*o), V(*d)).sphereDiscriminant(S(V(*c), r, M()), point)
        if it.intersects:
            rows.append([*o, *d, *c, r, point, 1, it.distance, *it.point.getXYZ(), *it.normal.getXYZ()])
        else:
            rows.append([*o, *d, *c, r, point, 0] + [0.0] * 7)
    disc = np.array(rows, np.float64)
    refl, refr = [], []
    for i in range(200):
        v, n = rs.uniform(-1, 1, 3), rs.uniform(-1, 1, 3)
        out = V(*v).reflectInVector(V(*n))
        refl.append([*v, *n, *out.getXYZ()])
        ra, rb = (1.0, float(rs.uniform(1.1, 2.2))) if i % 2 else (float(rs.uniform(1.1, 2.2)), 1.0)
        out = V(*v).refractInVector(V(*n), ra, rb)
        refr.append([*v, *n, ra, rb, 0, 0, 0, 0] if out is False else [*v, *n, ra, rb, 1, *out.getXYZ()])
    # RL/Marbles 1.ipynb cell 7 (stored notebook output; the one valid KAT the reference ships)
    hit = Ray(V(0.1, 0, 5), V(0, 0, -1)).sphereDiscriminant(S(V(0, 0, 0), 1, M()))
    nb7 = np.array([*hit.point.getXYZ(), *V(0, 0, -1).refractInVector(hit.normal, 1, 1.5).getXYZ()])
    np.savez_compressed(OUT / "kat.npz", disc=disc, reflect=np.array(refl), refract=np.array(refr), notebook7=nb7)
    print("kat.npz", disc.shape, int(disc[:, 11].sum()), "hits;", int(np.array(refr)[:, 8].sum()), "refractions")


# ----------------------------------------------------------------- Algorithm A frames
def ref_whitted_frame(ref, spec, X, Y, max_bounces, prenorm):
    """The notebooks' / render_true_original's pixel loop (RL/output5.py:491-518), reference objects only."""
    V, Ray = ref.ns.Vector, ref.ray.Ray
    cam = V(*spec.camera)
    H, W = len(Y), len(X)
    rgb = np.zeros((H, W, 3))
    hit = np.full((H, W), -1, np.int32)
    index = {id(s): i for i, s in enumerate(spec.spheres)}
    for yi, Yv in enumerate(Y):
        for xi, Xv in enumerate(X):
            d = V(float(Xv), float(Yv), -1)
            if prenorm:
                d = d.normalise()
            t = Ray(cam, d).nearestSphereIntersect(spec.spheres, max_bounces=max_bounces)
            if t is None:
                rgb[yi, xi] = spec.miss.getList()
            else:
                hit[yi, xi] = index[id(t.object)]
                rgb[yi, xi] = t.terminalRGB(spheres=spec.spheres, background_colour=spec.background,
                                            global_light_sources=spec.global_lights,
                                            point_light_sources=spec.point_lights).getList()
    return rgb, hit


def gen_whitted(ref):
    # C1: balls_in_space through the reference's own render_custom_scene('traditional') entry point
    spec = scenes.build_balls_in_space(ref.ns, as_rendered=True)
    raw = scenes.build_balls_in_space(ref.ns, as_rendered=False)
    with contextlib.redirect_stdout(io.StringIO()):
        theirs = ref.output5.create_custom_scene()["custom_scene"]
    assert_same_scene(raw.spheres, theirs, "balls_in_space")
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()), \
            contextlib.redirect_stderr(io.StringIO()):
        exp = ref.output5.CustomSceneExperiment(output_dir=tmp)
        exp.config.update(image_width=320, image_height=240, samples_per_pixel=1, max_bounces=1)
        _, img = exp.render_custom_scene(theirs, "traditional", Path(tmp) / "x.png")
    X, Y = scenes.custom_scene_grid(320, 240)
    rgb, hit = ref_whitted_frame(ref, spec, X, Y, 1, True)
    assert np.array_equal(np.minimum(1.0, np.floor(rgb) / 255.0).astype(np.float32), img), "driver loop != render_custom_scene"
    np.savez_compressed(OUT / "whitted_c1_balls_320x240.npz", rgb=rgb.astype(np.float32), hit=hit.astype(np.int16),
                        image=img, X=X, Y=Y, cam=np.array(spec.camera), max_bounces=1, prenorm=1,
                        miss=np.array(spec.miss.getList(), float), **flat_dict(flat(spec)))
    print("whitted_c1", rgb.shape, "hit px", int((hit >= 0).sum()))

    # same scene, spp 4 with Philox-fed jitter through render_custom_scene (RL/output5.py:1463-1470)
    W, H, spp, seed = 80, 60, 4, 7
    state = {"k": 0}

    def fake_random():
        k = state["k"]; state["k"] += 1
        pix, sm, w = k // (2 * spp), (k // 2) % spp, k % 2
        return orc.rng_pair(seed, pix, sm, 0)[w]

    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()), \
            contextlib.redirect_stderr(io.StringIO()), mock.patch.object(np.random, "random", fake_random):
        exp = ref.output5.CustomSceneExperiment(output_dir=tmp)
        exp.config.update(image_width=W, image_height=H, samples_per_pixel=spp, max_bounces=6)
        _, img4 = exp.render_custom_scene(theirs, "traditional", Path(tmp) / "x.png")
    X4, Y4 = scenes.custom_scene_grid(W, H)
    np.savez_compressed(OUT / "whitted_balls_spp4_80x60.npz", image=img4, X=X4, Y=Y4, cam=np.array(spec.camera),
                        max_bounces=6, prenorm=1, spp=spp, seed=seed, miss=np.array(spec.miss.getList(), float),
                        **flat_dict(flat(spec)))
    print("whitted spp4", img4.shape)

    # render_true_original geometry (notebook grid, depth 5, no pre-normalisation), reduced to 121x121
    X, Y = scenes.notebook_grid(60, 0.01 / 3 * 5)
    rgb, hit = ref_whitted_frame(ref, spec, X, Y, 5, False)
    np.savez_compressed(OUT / "whitted_balls_true_original_121.npz", rgb=rgb.astype(np.float32), hit=hit.astype(np.int16),
                        X=X, Y=Y, cam=np.array(spec.camera), max_bounces=5, prenorm=0,
                        miss=np.array(spec.miss.getList(), float), **flat_dict(flat(spec)))
    print("whitted true_original", int((hit >= 0).sum()))

    # C2: marbles / shadows scenes (notebook loops), two depths each
    for name, builder in (("marbles4", scenes.build_marbles4), ("planets2", scenes.build_planets2)):
        spec = builder(ref.ns)
        X, Y = scenes.notebook_grid(60, spec.ray_step * 100 / 60)
        for depth in (4, 8 if name == "marbles4" else 10):
            rgb, hit = ref_whitted_frame(ref, spec, X, Y, depth, False)
            np.savez_compressed(OUT / f"whitted_{name}_d{depth}_121.npz", rgb=rgb.astype(np.float32),
                                hit=hit.astype(np.int16), X=X, Y=Y, cam=np.array(spec.camera), max_bounces=depth,
                                prenorm=0, miss=np.array(spec.miss.getList(), float), **flat_dict(flat(spec)))
            print(f"whitted_{name}_d{depth}", int((hit >= 0).sum()), "max", rgb.max())


# ----------------------------------------------------------------- Algorithm B frames
def ref_path_frame(module, spheres, camera, W, H, spp, depth, seed):
    """TraditionalRenderer.render (the reference's own loops) with the Philox stream patched in."""
    ctl = types.SimpleNamespace(n=0, phase=0, bounce=0, word={})

    class Hooked(module.TraditionalRenderer):
        def trace_ray_traditional(self, ray, bounce_count=0):
            if bounce_count == 0:
                ctl.phase, ctl.word = 1, {}
            ctl.bounce = bounce_count
            out = super().trace_ray_traditional(ray, bounce_count)
            if bounce_count == 0:
                sums[ctl.n // spp] += [out.r, out.g, out.b]
                ctl.n += 1
                ctl.phase, ctl.word = 0, {}
            return out

    def fake_random():
        pix, sm = ctl.n // spp, ctl.n % spp
        slot = 0 if ctl.phase == 0 else ctl.bounce + 1
        w = ctl.word.get(slot, 0)
        ctl.word[slot] = w + 1
        assert w < 2
        return orc.rng_pair(seed, pix, sm, slot)[w]

    sums = np.zeros((W * H, 3))
    r = Hooked()
    r.scene = spheres
    r.light_sources = [s for s in spheres if s.material.emitive]
    r.small_lights = [s for s in r.light_sources if s.radius < 0.5]
    r.camera_position = camera
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), \
            mock.patch.object(np.random, "random", fake_random), \
            mock.patch.object(module, "tqdm", lambda it, **k: it):
        img = r.render(W, H, spp, depth)
    stats = np.array([r.stats[k] for k in ("total_rays", "total_intersections", "light_hits", "small_light_hits")],
                     np.int64)
    return img, sums.reshape(H, W, 3), stats


def gen_path(ref):
    V = ref.ns.Vector
    spec = scenes.build_chandelier(ref.ns)
    assert_same_scene(spec.spheres, ref.chandelier.generate_chandelier_scene(), "chandelier")
    W, H, spp, depth, seed = 48, 27, 4, 8, 11
    img, sums, stats = ref_path_frame(ref.chandelier, spec.spheres, V(*spec.camera), W, H, spp, depth, seed)
    np.savez_compressed(OUT / "path_chandelier_48x27.npz", image=img, sums=sums.astype(np.float32), stats=stats,
                        W=W, H=H, spp=spp, max_bounces=depth, seed=seed, mirror_threshold=0.0,
                        cam=np.array(spec.camera), **flat_dict(flat(spec)))
    print("path_chandelier", stats, "rays/sample", stats[0] / (W * H * spp))

    spec = scenes.build_complex(ref.ns)
    W, H, spp, depth, seed = 48, 27, 4, 5, 12
    img, sums, stats = ref_path_frame(ref.complex, spec.spheres, V(*spec.camera), W, H, spp, depth, seed)
    np.savez_compressed(OUT / "path_complex_48x27.npz", image=img, sums=sums.astype(np.float32), stats=stats,
                        W=W, H=H, spp=spp, max_bounces=depth, seed=seed, mirror_threshold=0.9,
                        cam=np.array(spec.camera), **flat_dict(flat(spec)))
    print("path_complex", stats, "rays/sample", stats[0] / (W * H * spp))


def gen_path_native(ref):
    """The UNMODIFIED reference with its OWN random numbers: TraditionalRenderer.render of the chandelier scene with
    numpy's global MT19937 (seeded, not patched).  The CUDA path cannot reproduce that stream; this golden pins the
    statistical agreement of the mean image (tests/test_gpu_parity.py::test_mean_image_agrees_with_the_reference_rng)."""
    V = ref.ns.Vector
    spec = scenes.build_chandelier(ref.ns)
    W, H, spp, depth = 40, 24, 48, 8
    r = ref.chandelier.TraditionalRenderer()
    r.scene = spec.spheres
    r.light_sources = [s for s in spec.spheres if s.material.emitive]
    r.small_lights = [s for s in r.light_sources if s.radius < 0.5]
    r.camera_position = V(*spec.camera)
    np.random.seed(20240229)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), \
            mock.patch.object(ref.chandelier, "tqdm", lambda it, **k: it):
        img = r.render(W, H, spp, depth)
    stats = np.array([r.stats[k] for k in ("total_rays", "total_intersections", "light_hits", "small_light_hits")], np.int64)
    np.savez_compressed(OUT / "path_chandelier_native_rng_40x24.npz", image=np.asarray(img, np.float32), stats=stats, W=W, H=H,
                        spp=spp, max_bounces=depth, mirror_threshold=0.0, cam=np.array(spec.camera), **flat_dict(flat(spec)))
    print("path_chandelier_native_rng", stats, "rays/sample", stats[0] / (W * H * spp))


# ----------------------------------------------------------------- env rollouts
REASON = {None: 0, "ray_missed": 1, "ray_escaped": 2, "max_bounces": 3, "hit_sun": 4, "already_on_sun": 5}


def build_env_demo(ns):
    """The ``__main__`` demo scene of both env files (RL/ray_tracer_env.py:429-480, FB/ray_tracer_env.py:542-590)."""
    V, C, M, S = ns.Vector, ns.Colour, ns.Material, ns.Sphere
    matte = M(reflective=0, transparent=0, emitive=0.1, refractive_index=1)
    mirror = M(reflective=1, transparent=0, emitive=0, refractive_index=1)
    glass = M(reflective=0, transparent=1, emitive=0, refractive_index=1.5)
    spheres = [S(V(0, -100.5, -3), 100, matte, C(200, 200, 200), id=1), S(V(0, 0, -3), 0.5, mirror, C(255, 255, 255), id=2),
               S(V(-1.2, 0, -3), 0.5, mirror, C(200, 200, 255), id=3), S(V(1.2, 0, -3), 0.5, glass, C(255, 200, 200), id=4),
               S(V(0, 2, -3), 0.3, M(reflective=0, transparent=0, emitive=1, refractive_index=1), C(255, 255, 200), id=99)]
    gl = [ns.GlobalLight(vector=V(0, -1, -0.5).normalise(), colour=C(200, 200, 255), strength=0.3, max_angle=np.pi / 3)]
    pl = [ns.PointLight(id=99, position=V(0, 2, -3), colour=C(255, 255, 200), strength=5.0, max_angle=np.pi, func=0)]
    return scenes.SceneSpec(spheres=spheres, global_lights=gl, point_lights=pl, background=C(0, 0, 0),
                            camera=(0.0, 0.0, 0.0), width=400, height=300, fov=90, max_bounces=10)


def gen_env(ref):
    V, A, C = ref.ns.Vector, ref.ns.Angle, ref.ns.Colour
    cases = []
    opt = scenes.build_optimized_env_scene(ref.ns)
    theirs, _, their_pl = ref.improved.create_optimized_scene()
    assert_same_scene(opt.spheres, theirs, "optimized env scene")
    cases.append(("rl_optimized", "rl", opt, (0, 0, 0)))
    demo = build_env_demo(ref.ns)
    cases.append(("rl_demo", "rl", demo, (0, 0, 0)))
    cases.append(("fb_demo", "fb", demo, (0, 0, 0)))
    balls = scenes.build_balls_in_space(ref.ns, as_rendered=False)
    balls.point_lights = [ref.ns.PointLight(id=7, position=balls.sun.centre, colour=balls.sun.colour, strength=1,
                                            max_angle=np.radians(90), func=-1)]
    balls.width, balls.height, balls.fov, balls.max_bounces = 160, 120, 60, 5
    balls.camera = (0.0, 0.0, 1.0)
    cases.append(("fb_balls", "fb", balls, (0, 0, 0)))
    cases.append(("rl_balls_rotated", "rl", balls, (0.1, -0.05, 0.02)))
    cases.append(("rl_adaptive", "rl", opt, (0, 0, 0)))            # AdaptiveRewardRayTracerEnv on its own training scene
    B = 96
    only = os.environ.get("GEN_ENV_ONLY")
    for name, flavour, spec, angle in cases:
        if only and name != only:
            continue
        mod = ref.env_fb if flavour == "fb" else ref.env_rl
        cls = ref.optimized.AdaptiveRewardRayTracerEnv if name == "rl_adaptive" else mod.RayTracerEnv
        rs = np.random.RandomState(abs(hash(name)) % 2 ** 31 if False else sum(map(ord, name)))
        pixels = np.stack([rs.randint(0, spec.width, B), rs.randint(0, spec.height, B)], 1).astype(np.int32)
        pixels[0] = (spec.width // 2, spec.height // 2)
        T = spec.max_bounces + 3
        lo, hi = ((-1, -1), (1, 1)) if flavour == "fb" else ((0, 0), (np.pi / 2, 2 * np.pi))
        actions = rs.uniform(lo, hi, (T, B, 2)).astype(np.float32)
        obs0 = np.zeros((B, 18), np.float32)
        obs = np.zeros((T, B, 18), np.float32)
        rew = np.zeros((T, B))
        term = np.zeros((T, B), np.uint8)
        trunc = np.zeros((T, B), np.uint8)
        reason = np.zeros((T, B), np.int32)
        total = np.zeros((T, B))
        for b in range(B):
            with contextlib.redirect_stdout(io.StringIO()):
                env = cls(spheres=spec.spheres, image_width=spec.width, image_height=spec.height,
                          camera_position=V(*spec.camera), camera_angle=A(*angle), fov=spec.fov,
                          max_bounces=spec.max_bounces, background_colour=spec.background,
                          global_light_sources=spec.global_lights, point_light_sources=spec.point_lights)
                obs0[b], _ = env.reset(options={"pixel": (int(pixels[b, 0]), int(pixels[b, 1]))})
                for t in range(T):      # keeps stepping after termination, like a careless caller would
                    o, r, te, tr, info = env.step(actions[t, b])
                    obs[t, b], rew[t, b], term[t, b], trunc[t, b] = o, r, te, tr
                    reason[t, b] = REASON[info.get("reason")]
                    total[t, b] = info["total_reward"]
        np.savez_compressed(OUT / f"env_{name}.npz", flavour=flavour, pixels=pixels, actions=actions, obs0=obs0, obs=obs,
                            reward=rew, terminated=term, truncated=trunc, reason=reason, total_reward=total,
                            width=spec.width, height=spec.height, fov=spec.fov, max_bounces=spec.max_bounces,
                            cam=np.array(spec.camera), cam_angle=np.array(angle, float), **flat_dict(flat(spec)))
        print(f"env_{name}", "first-hit rate", float((np.abs(obs0).sum(1) > 0).mean()), "reasons",
              np.bincount(reason.ravel(), minlength=6))


# ----------------------------------------------------------------- output6 ("Algorithm C")
def gen_simple(ref):
    """FB/output6.py SimplifiedFBRenderer.render_original_style in traditional mode (fb_usage_prob = 0, no model).
    ``fb_ray_tracing`` is absent from the reference, so ``__init__`` would raise: the instance is made with
    ``object.__new__`` and given exactly the attributes ``__init__`` sets (output6.py:96-126); every method that runs
    (render_original_style, trace_ray_simple, calculate_lighting_exact_original) is the reference's own."""
    mod = _load("ref_output6", REF / "FB" / "output6.py")
    spec = scenes.build_balls_in_space(ref.ns, as_rendered=False)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_scene = mod.create_your_custom_scene()
    assert_same_scene(spec.spheres, ref_scene, "output6 custom scene")
    ctl = types.SimpleNamespace(pixel=-1, bounce=0, word=0)

    class Hooked(mod.SimplifiedFBRenderer):
        def trace_ray_simple(self, ray):
            ctl.pixel += 1
            ctl.bounce = -1
            out = super().trace_ray_simple(ray)
            colours.append([out.r, out.g, out.b])
            return out

        def calculate_lighting_exact_original(self, intersection):     # once per bounce, before any draw
            ctl.bounce += 1
            ctl.word = 0
            out = super().calculate_lighting_exact_original(intersection)
            if len(lit) < 4096:          # the helper on its own: (intersection -> Colour) pairs of the reference
                k = next(j for j, sph in enumerate(self.scene) if sph is intersection.object)
                p_, n_ = intersection.point, intersection.normal
                lit.append([p_.x, p_.y, p_.z, n_.x, n_.y, n_.z, float(k), out.r, out.g, out.b])
            return out

    def fake_random():
        w = ctl.word
        ctl.word += 1
        assert w < 2
        return orc.rng_pair(seed, ctl.pixel, 0, ctl.bounce + 1)[w]

    lit = []
    for tag, W, H, seed, depth in (("a", 64, 48, 21, 5), ("b", 40, 30, 22, 8)):
        colours = []
        ctl.pixel = -1
        r = object.__new__(Hooked)
        r.scene = ref_scene
        r.sun_position, r.sun_radius, r.sun_color = ref.ns.Vector(-0.6, 0.2, 6), 0.1, ref.ns.Colour(255, 255, 204)
        r.agent, r.fb_model_loaded = None, False
        r.max_bounces, r.samples_per_pixel, r.fb_usage_prob = depth, 100, 0.0
        r.stats = {}
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), \
                mock.patch.object(np.random, "random", fake_random), mock.patch.object(mod, "tqdm", lambda it, **k: it), \
                tempfile.TemporaryDirectory() as tmp:
            img, _ = r.render_original_style(W, H, os.path.join(tmp, "x.png"))
        rgb = np.array(colours, np.float64).reshape(H, W, 3)
        stats = np.array([r.stats["total_rays"], r.stats["sun_hits"]], np.int64)
        np.savez_compressed(OUT / f"simple_balls_{tag}_{W}x{H}.npz", image=img, rgb=rgb.astype(np.float32), stats=stats,
                            W=W, H=H, seed=seed, max_bounces=depth, **flat_dict(rtb.flatten_scene(spec.spheres)))
        print(f"simple_balls_{tag}", stats, "mean", rgb.mean(axis=(0, 1)))
    lit = np.array(lit, np.float64)
    np.savez_compressed(OUT / "simple_lighting_balls.npz", hits=lit[:, :7], rgb=lit[:, 7:].astype(np.float32),
                        **flat_dict(rtb.flatten_scene(spec.spheres)))
    print("simple_lighting", lit.shape, "sun rows", int((lit[:, 6] == 7).sum()), "mean", lit[:, 7:].mean(axis=0))


# ----------------------------------------------------------------- FB-guided Algorithm B (f-4)
def test_policy(obs):
    """The stand-in for ``fb_agent.choose_direction``: IEEE basic operations only (products with powers of two, adds,
    clip), so numpy on the CPU and torch on the GPU give the same float32 bits.  Returned as float64 -- see
    trace_path_fb in rt_oracle.c for why."""
    o = np.asarray(obs, np.float32)
    a0 = np.clip(o[6] * np.float32(0.5) + o[7] * np.float32(0.25) - np.float32(0.125), np.float32(-1), np.float32(1))
    a1 = np.clip(o[8] * np.float32(0.5) + o[3] * np.float32(0.25) + o[16] * np.float32(0.5), np.float32(-1), np.float32(1))
    return np.array([a0, a1], dtype=np.float64)


def gen_fb(ref):
    """WorkingFBRenderer.render (FB/fb_vs_traditional_complex.py:425-640), the reference's own class, with a stand-in
    agent (the trained checkpoints are not in the reference) and np.random.random patched to the Philox streams."""
    V = ref.ns.Vector
    mod = ref.complex
    spec = scenes.build_complex(ref.ns)
    W, H, spp, depth, seed, prob = 40, 24, 3, 5, 41, 0.6
    ctl = types.SimpleNamespace(n=0, phase=0, bounce=0, draws=0)

    class Hooked(mod.WorkingFBRenderer):
        def trace_ray_fb(self, ray, bounce_count=0, accumulated_color=ref.ns.Colour(0, 0, 0)):
            if bounce_count == 0:
                ctl.phase = 1
            ctl.bounce, ctl.draws = bounce_count, 0
            out = super().trace_ray_fb(ray, bounce_count, accumulated_color)
            if bounce_count == 0:
                sums[ctl.n // spp] += [out.r, out.g, out.b]
                ctl.n += 1
                ctl.phase, ctl.draws = 0, 0
            return out

    def fake_random():
        pix, sm = ctl.n // spp, ctl.n % spp
        k = ctl.draws
        ctl.draws += 1
        if ctl.phase == 0:                         # jitter x, y
            assert k < 2
            return orc.rng_pair(seed, pix, sm, 0)[k]
        if k == 0:                                 # the use-FB decision
            return (orc.philox([pix, sm, ctl.bounce, 0x52544642], [seed & 0xffffffff, seed >> 32])[0] >> 8) / 16777216.0
        assert k < 3
        return orc.rng_pair(seed, pix, sm, ctl.bounce + 1)[k - 1]

    sums = np.zeros((W * H, 3))
    r = Hooked(model_path=None, scene_small_lights=[s for s in spec.spheres if s.material.emitive and s.radius < 0.5],
               camera_position=V(*spec.camera))
    r.scene = spec.spheres
    r.light_sources = [s for s in spec.spheres if s.material.emitive]
    r.fb_agent = types.SimpleNamespace(choose_direction=test_policy)
    r.fb_loaded, r.fb_usage_prob = True, prob
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), \
            mock.patch.object(np.random, "random", fake_random), mock.patch.object(mod, "tqdm", lambda it, **k: it):
        img = r.render(W, H, spp, depth)
    stats = np.array([r.stats[k] for k in ("total_rays", "total_intersections", "light_hits", "small_light_hits", "fb_used")],
                     np.int64)
    np.savez_compressed(OUT / "path_fb_complex_40x24.npz", image=img, sums=sums.reshape(H, W, 3).astype(np.float32), stats=stats,
                        W=W, H=H, spp=spp, max_bounces=depth, seed=seed, mirror_threshold=0.9, fb_usage_prob=prob,
                        cam=np.array(spec.camera), **flat_dict(flat(spec)))
    print("path_fb_complex", stats)


# ----------------------------------------------------------------- FB training trajectories (f-2)
def gen_traj(ref):
    """RayTracedComplexTrainer.generate_trajectory (FB/train_complex_only.py:254-334), the reference's own method, with
    ``random.choice / uniform / random`` patched to the Philox stream.  ``fb_ray_tracing`` and
    ``fb_multi_scene_trainer`` are absent from the reference: stub modules with empty classes stand in for the
    imports, and the trainer is made with ``object.__new__`` (only ``self.max_bounces`` is read)."""
    import random as pyrandom
    for name, attrs in (("fb_ray_tracing", ("FBResearchAgent", "FBConfig")), ("fb_multi_scene_trainer", ("MultiSceneFBTrainer",))):
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, type(a, (), {}))
        sys.modules[name] = m
    mod = _load("ref_train_complex", REF / "FB" / "train_complex_only.py")
    spec = scenes.build_complex(ref.ns)
    seed, n_traj, max_steps = 31, 256, 8
    ctl = types.SimpleNamespace(j=-1, draws=0)

    def nxt():
        k = ctl.draws
        ctl.draws += 1
        # draw order: choice, theta | phi, - | then (r1, r2) pairs from slot 2
        slot, word = (0, k) if k < 2 else (1, 0) if k == 2 else (2 + (k - 3) // 2, (k - 3) % 2)
        return orc.rng_pair(seed, ctl.j, 0, slot)[word]

    def fake_choice(seq):
        return seq[min(len(seq) - 1, int(nxt() * len(seq)))]

    def fake_uniform(a, b):
        return a + (b - a) * nxt()

    tr = object.__new__(mod.RayTracedComplexTrainer)
    tr.max_bounces = 8
    rows = {k: [] for k in ("obs", "action", "next_obs", "reward", "hit")}
    length, hit_light = [], []
    with contextlib.redirect_stdout(io.StringIO()), mock.patch.object(pyrandom, "choice", fake_choice), \
            mock.patch.object(pyrandom, "uniform", fake_uniform), mock.patch.object(pyrandom, "random", nxt):
        for j in range(n_traj):
            ctl.j, ctl.draws = j, 0
            transitions, hit = tr.generate_trajectory(spec.spheres, max_steps=max_steps)
            length.append(len(transitions)); hit_light.append(int(hit))
            pad = {"obs": np.zeros((max_steps, 22), np.float32), "action": np.zeros((max_steps, 2), np.float32),
                   "next_obs": np.zeros((max_steps, 22), np.float32), "reward": np.zeros(max_steps, np.float32),
                   "hit": np.zeros(max_steps, np.uint8)}
            for t, (o, a, no, r, h) in enumerate(transitions):
                pad["obs"][t], pad["action"][t], pad["next_obs"][t], pad["reward"][t], pad["hit"][t] = o, a, no, r, h
            for k in rows:
                rows[k].append(pad[k])
    np.savez_compressed(OUT / "traj_complex_256.npz", seed=seed, max_steps=max_steps, max_bounces=8,
                        length=np.array(length, np.int32), hit_light=np.array(hit_light, np.uint8),
                        **{k: np.stack(v) for k, v in rows.items()}, **flat_dict(rtb.flatten_scene(spec.spheres)))
    print("traj_complex", "transitions", int(np.sum(length)), "light hits", int(np.sum(hit_light)))


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    ref = load_reference()
    which = sys.argv[1:] or ["kat", "whitted", "path", "path_native", "env", "simple", "traj", "fb"]
    for w in which:
        {"kat": gen_kat, "whitted": gen_whitted, "path": gen_path, "path_native": gen_path_native, "env": gen_env,
         "simple": gen_simple, "traj": gen_traj, "fb": gen_fb}[w](ref)


if __name__ == "__main__":
    main()
