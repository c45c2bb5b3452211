"""ctypes front-end of the CPU oracle (oracle/rt_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, ``__graft_entry__.smoke()`` and
bench.py's ``cpu_baseline`` / ``--impl reference`` legs as the *checker*.  The
product package (ray-tracer-v1_b200/) never imports this module.

Every entry takes a flat scene (any object with the attribute names of
``ray_tracer_v1_b200.scene.FlatScene``) plus plain numpy arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librt_oracle.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)
c_fp = C.POINTER(C.c_float)
c_u64p = C.POINTER(C.c_uint64)
NO_ID = -(2 ** 31)

REASONS = {0: None, 1: "ray_missed", 2: "ray_escaped", 3: "max_bounces", 4: "hit_sun", 5: "already_on_sun"}


class _Scene(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("centre", c_dp), ("radius", c_dp), ("material", c_dp), ("colour", c_dp), ("ids", c_ip),
        ("nG", C.c_int32), ("g_vec", c_dp), ("g_col", c_dp), ("g_strength", c_dp), ("g_max_angle", c_dp), ("g_func", c_ip),
        ("nP", C.c_int32), ("p_id", c_ip), ("p_pos", c_dp), ("p_col", c_dp), ("p_strength", c_dp), ("p_max_angle", c_dp),
        ("p_func", c_ip),
        ("bg", C.c_double * 3),
        ("nL", C.c_int32), ("l_centre", c_dp), ("l_colour", c_dp), ("l_index", c_ip), ("small", c_u8p),
    ]


class _EnvCfg(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("max_bounces", C.c_int32), ("flavour", C.c_int32),
                ("cam", C.c_double * 3), ("cam_angle", C.c_double * 3), ("fov", C.c_double), ("sun_id", C.c_int32),
                ("adaptive", C.c_int32), ("light_ids", C.c_int32 * 2)]


class _SimpleCfg(C.Structure):
    _fields_ = [("cam", C.c_double * 3), ("fov", C.c_double), ("sun_pos", C.c_double * 3), ("sun_col", C.c_double * 3),
                ("sun_id", C.c_int32), ("max_bounces", C.c_int32)]


POLICY_FN = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)


def build(force=False):
    """Compile oracle/rt_oracle.c with the committed Makefile (gcc, a second or two)."""
    src = os.path.join(_HERE, "rt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "librt_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_sphere_discriminant.restype = C.c_int
        _lib.orc_sphere_discriminant.argtypes = [c_dp, c_dp, c_dp, C.c_double, C.c_int, c_dp]
        _lib.orc_refract.restype = C.c_int
        _lib.orc_refract.argtypes = [c_dp, c_dp, C.c_double, C.c_double, c_dp]
        _lib.orc_reflect.argtypes = [c_dp, c_dp, c_dp]
        _lib.orc_rng_pair.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, c_dp]
        _lib.orc_philox.argtypes = [C.POINTER(C.c_uint32)] * 3
        _lib.orc_trace_rays.argtypes = [C.POINTER(_Scene), C.c_int, c_dp, c_ip, c_ip, C.c_int, C.c_int, c_dp, c_dp, c_dp]
        _lib.orc_shade_hits.argtypes = [C.POINTER(_Scene), C.c_int, c_dp, C.c_int, c_dp]
        _lib.orc_render_whitted.argtypes = [C.POINTER(_Scene), c_dp, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_int, c_dp, C.c_uint64, C.c_int, c_dp, c_ip, c_u64p,
                                            C.c_int]
        _lib.orc_render_path.argtypes = [C.POINTER(_Scene), c_dp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, c_dp, c_u64p, C.c_int]
        _lib.orc_env_reset.argtypes = [C.POINTER(_Scene), C.POINTER(_EnvCfg), C.c_int, c_ip, C.c_void_p, c_fp]
        _lib.orc_env_step.argtypes = [C.POINTER(_Scene), C.POINTER(_EnvCfg), C.c_int, c_fp, C.c_void_p, c_fp, c_dp,
                                      c_u8p, c_u8p, c_ip]
        _lib.orc_simple_lighting.argtypes = [C.POINTER(_Scene), C.POINTER(_SimpleCfg), C.c_int, c_dp, c_dp,
                                             C.POINTER(C.c_uint64)]
        _lib.orc_simple_lighting.restype = None
        _lib.orc_render_simple.argtypes = [C.POINTER(_Scene), C.POINTER(_SimpleCfg), C.c_int, C.c_int, C.c_uint64, c_dp, c_dp,
                                           c_u64p, C.c_int]
        _lib.orc_generate_trajectories.argtypes = [C.POINTER(_Scene), C.c_int, C.c_int, C.c_int, C.c_uint64, c_fp, c_fp, c_fp,
                                                   c_fp, c_u8p, c_ip, c_u8p]
        _lib.orc_render_path_fb.argtypes = [C.POINTER(_Scene), c_dp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                            C.c_double, C.c_double, C.c_uint64, POLICY_FN, C.c_void_p, c_dp, c_u64p]
        _lib.orc_sizeof_env.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _d(a):
    return np.ascontiguousarray(a, np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


class OracleScene:
    """Keeps the numpy buffers alive next to the C struct that points into them."""

    def __init__(self, fs):
        k = self._keep = {}
        for name in ("centre", "radius", "material", "colour", "g_vec", "g_col", "g_strength", "g_max_angle",
                     "p_pos", "p_col", "p_strength", "p_max_angle", "l_centre", "l_colour"):
            k[name] = _d(getattr(fs, name))
        for name in ("ids", "g_func", "p_id", "p_func", "l_index"):
            k[name] = np.ascontiguousarray(getattr(fs, name), np.int32)
        n = k["radius"].shape[0]
        small = getattr(fs, "small", None)
        k["small"] = np.ascontiguousarray(small if small is not None else np.zeros(n), np.uint8)
        s = self.c = _Scene()
        s.n, s.nG, s.nP, s.nL = n, k["g_strength"].shape[0], k["p_strength"].shape[0], k["l_index"].shape[0]
        for name in ("centre", "radius", "material", "colour", "g_vec", "g_col", "g_strength", "g_max_angle",
                     "p_pos", "p_col", "p_strength", "p_max_angle", "l_centre", "l_colour"):
            setattr(s, name, _p(k[name], c_dp))
        for name in ("ids", "g_func", "p_id", "p_func", "l_index"):
            setattr(s, name, _p(k[name], c_ip))
        s.small = _p(k["small"], c_u8p)
        s.bg[:] = [float(x) for x in np.asarray(fs.bg).reshape(3)]
        self.n = n

    @property
    def ref(self):
        return C.byref(self.c)


def _scene(fs):
    return fs if isinstance(fs, OracleScene) else OracleScene(fs)


# ------------------------------------------------------------------ unit-level
def sphere_discriminant(origin, direction, centre, radius, point=0):
    """Ray(origin, direction).sphereDiscriminant(Sphere(centre, radius), point) -> (hit, t, p[3], n[3])."""
    out = np.zeros(8)
    hit = lib().orc_sphere_discriminant(_p(_d(origin), c_dp), _p(_d(direction), c_dp), _p(_d(centre), c_dp),
                                        float(radius), int(point), _p(out, c_dp))
    return bool(hit), out[0], out[1:4].copy(), out[4:7].copy()


def reflect(v, n):
    out = np.zeros(3)
    lib().orc_reflect(_p(_d(v), c_dp), _p(_d(n), c_dp), _p(out, c_dp))
    return out


def refract(v, n, ra, rb):
    out = np.zeros(3)
    ok = lib().orc_refract(_p(_d(v), c_dp), _p(_d(n), c_dp), float(ra), float(rb), _p(out, c_dp))
    return out if ok else False


def rng_pair(seed, pixel, sample, slot):
    out = np.zeros(2)
    lib().orc_rng_pair(int(seed), int(pixel), int(sample), int(slot), _p(out, c_dp))
    return float(out[0]), float(out[1])


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    u32p = C.POINTER(C.c_uint32)
    lib().orc_philox(_p(c, u32p), _p(k, u32p), _p(o, u32p))
    return o


def trace_rays(fs, rays, suppress=None, bounces0=None, max_bounces=1, shadow_max_bounces=0, miss=(0, 0, 0), shade=True):
    """Batch of ``Ray.nearestSphereIntersect`` (+ ``terminalRGB``).

    rays [m,6] (origin, raw direction).  Returns (term [m,10], rgb [m,3] or None);
    term = hit, scene index, bounces, through_count, point(3), normal(3), distance."""
    sc = _scene(fs)
    rays = _d(rays).reshape(-1, 6)
    m = rays.shape[0]
    term = np.zeros((m, 11))
    rgb = np.zeros((m, 3)) if shade else None
    sup = None if suppress is None else np.ascontiguousarray(suppress, np.int32)
    b0 = None if bounces0 is None else np.ascontiguousarray(bounces0, np.int32)
    lib().orc_trace_rays(sc.ref, m, _p(rays, c_dp), None if sup is None else _p(sup, c_ip),
                         None if b0 is None else _p(b0, c_ip), int(max_bounces), int(shadow_max_bounces),
                         _p(_d(miss), c_dp), _p(term, c_dp), None if rgb is None else _p(rgb, c_dp))
    return term, rgb


def shade_hits(fs, hits, shadow_max_bounces=0):
    """``Intersection.terminalRGB`` at given hits [m,7] = scene index, point(3), normal(3) -> rgb [m,3]."""
    sc = _scene(fs)
    hits = _d(hits).reshape(-1, 7)
    rgb = np.zeros((hits.shape[0], 3))
    lib().orc_shade_hits(sc.ref, hits.shape[0], _p(hits, c_dp), int(shadow_max_bounces), _p(rgb, c_dp))
    return rgb


# ------------------------------------------------------------------ frames
def render_whitted(fs, cam, X, Y, spp=1, max_bounces=1, shadow_max_bounces=0, miss=None, seed=0, prenorm=False,
                   rows=None, nthreads=0):
    """Algorithm A frame.  Returns (sum [H,W,3] f64, hit [H,W] i32, queries)."""
    sc = _scene(fs)
    X, Y = _d(X), _d(Y)
    W, H = X.shape[0], Y.shape[0]
    y0, y1 = rows if rows is not None else (0, H)
    miss = _d(sc.c.bg[:] if miss is None else miss)
    out = np.zeros((H, W, 3))
    hit = np.full((H, W), -1, np.int32)
    q = C.c_uint64(0)
    lib().orc_render_whitted(sc.ref, _p(_d(cam), c_dp), _p(X, c_dp), _p(Y, c_dp), W, H, int(y0), int(y1), int(spp),
                             int(max_bounces), int(shadow_max_bounces), _p(miss, c_dp), int(seed), int(bool(prenorm)),
                             _p(out, c_dp), _p(hit, c_ip), C.byref(q), int(nthreads))
    return out, hit, int(q.value)


def render_path(fs, cam, W, H, spp, max_bounces, mirror_threshold, seed=0, fov=60.0, rows=None, samples=None,
                nthreads=0):
    """Algorithm B frame.  Returns (sum [H,W,3] f64 over the sample range, stats dict)."""
    sc = _scene(fs)
    y0, y1 = rows if rows is not None else (0, H)
    s0, s1 = samples if samples is not None else (0, spp)
    out = np.zeros((H, W, 3))
    st = (C.c_uint64 * 5)()
    lib().orc_render_path(sc.ref, _p(_d(cam), c_dp), int(W), int(H), float(fov), int(y0), int(y1), int(s0), int(s1),
                          int(max_bounces), float(mirror_threshold), int(seed), _p(out, c_dp), st, int(nthreads))
    stats = {"total_rays": int(st[0]), "total_intersections": int(st[1]), "light_hits": int(st[2]),
             "small_light_hits": int(st[3]), "queries": int(st[4])}
    return out, stats


def render_simple(fs, W, H, cam=(0, 0, 1), fov=np.pi / 3, sun_pos=(-0.6, 0.2, 6), sun_col=(255, 255, 204), sun_id=7,
                  max_bounces=5, seed=0, rays=None, nthreads=0):
    """FB/output6.py ``render_original_style`` / ``trace_ray_simple`` (traditional mode).  Returns (rgb [H,W,3] f64
    integer-valued colours, stats dict); with ``rays`` [m,6] the m explicit rays are traced instead (W = m, H = 1)."""
    sc = _scene(fs)
    cfg = _SimpleCfg()
    cfg.cam[:] = [float(x) for x in cam]
    cfg.fov = float(fov)
    cfg.sun_pos[:] = [float(x) for x in sun_pos]
    cfg.sun_col[:] = [float(x) for x in sun_col]
    cfg.sun_id, cfg.max_bounces = int(sun_id), int(max_bounces)
    rp = None
    if rays is not None:
        rays = _d(rays).reshape(-1, 6)
        W, H, rp = rays.shape[0], 1, _p(rays, c_dp)
    out = np.zeros((H, W, 3))
    st = (C.c_uint64 * 2)()
    lib().orc_render_simple(sc.ref, C.byref(cfg), int(W), int(H), int(seed), rp, _p(out, c_dp), st, int(nthreads))
    return out, {"total_rays": int(st[0]), "sun_hits": int(st[1])}


def simple_lighting(fs, hits, sun_pos=(-0.6, 0.2, 6), sun_col=(255, 255, 204), sun_id=7):
    """FB/output6.py ``calculate_lighting_exact_original`` on given intersections: hits [m,7] = point, normal, scene
    index -> (rgb [m,3] f64 integer-valued, sun_hits)."""
    sc = _scene(fs)
    cfg = _SimpleCfg()
    cfg.sun_pos[:] = [float(x) for x in sun_pos]
    cfg.sun_col[:] = [float(x) for x in sun_col]
    cfg.sun_id = int(sun_id)
    hits = _d(hits).reshape(-1, 7)
    out = np.zeros((hits.shape[0], 3))
    st = (C.c_uint64 * 2)()
    lib().orc_simple_lighting(sc.ref, C.byref(cfg), int(hits.shape[0]), _p(hits, c_dp), _p(out, c_dp), st)
    return out, int(st[1])


def generate_trajectories(fs, n_traj, max_steps=8, max_bounces=8, seed=0):
    """FB/train_complex_only.py ``generate_trajectory`` x n_traj -> dict of obs/action/next_obs/reward/hit (padded to
    max_steps), length, hit_light."""
    sc = _scene(fs)
    n, m = int(n_traj), int(max_steps)
    out = {"obs": np.zeros((n, m, 22), np.float32), "action": np.zeros((n, m, 2), np.float32),
           "next_obs": np.zeros((n, m, 22), np.float32), "reward": np.zeros((n, m), np.float32),
           "hit": np.zeros((n, m), np.uint8), "length": np.zeros(n, np.int32), "hit_light": np.zeros(n, np.uint8)}
    lib().orc_generate_trajectories(sc.ref, n, m, int(max_bounces), int(seed), _p(out["obs"], c_fp), _p(out["action"], c_fp),
                                    _p(out["next_obs"], c_fp), _p(out["reward"], c_fp), _p(out["hit"], c_u8p),
                                    _p(out["length"], c_ip), _p(out["hit_light"], c_u8p))
    return out


def render_path_fb(fs, cam, W, H, spp, max_bounces, mirror_threshold, policy, fb_usage_prob=1.0, seed=0, fov=60.0,
                   samples=None):
    """WorkingFBRenderer.render (FB/fb_vs_traditional_complex.py:487-640): Algorithm B with ``policy(obs22 float32
    array) -> action (2,)`` choosing the diffuse direction with probability ``fb_usage_prob`` (policy None = the
    traditional renderer).  Returns (sum [H,W,3], stats dict incl. fb_used)."""
    sc = _scene(fs)
    s0, s1 = samples if samples is not None else (0, spp)
    out = np.zeros((H, W, 3))
    st = (C.c_uint64 * 6)()

    def cb(obs, act, _user):
        a = policy(np.ctypeslib.as_array(obs, shape=(22,)).copy())
        act[0], act[1] = float(a[0]), float(a[1])

    fn = POLICY_FN(cb) if policy is not None else C.cast(None, POLICY_FN)
    lib().orc_render_path_fb(sc.ref, _p(_d(cam), c_dp), int(W), int(H), float(fov), int(s0), int(s1), int(max_bounces),
                             float(mirror_threshold), float(fb_usage_prob), int(seed), fn, None, _p(out, c_dp), st)
    return out, {"total_rays": int(st[0]), "total_intersections": int(st[1]), "light_hits": int(st[2]),
                 "small_light_hits": int(st[3]), "queries": int(st[4]), "fb_used": int(st[5])}


def resolve(sum_rgb, spp):
    """``pixel // spp`` then ``min(1, /255)`` -> float32 image (chandelier.py:540-549; output5.py:1500-1512)."""
    q = np.floor(np.asarray(sum_rgb, np.float64) / spp)
    return np.minimum(1.0, q / 255.0).astype(np.float32)


# ------------------------------------------------------------------ env
class OracleEnv:
    """Batched RayTracerEnv (flavour 'rl' = RL/ray_tracer_env.py, 'fb' = FB/ray_tracer_env.py)."""

    def __init__(self, fs, B, width, height, camera=(0, 0, 0), camera_angle=(0, 0, 0), fov=90, max_bounces=5,
                 flavour="rl", sun_id=7, adaptive=False, light_ids=(99, 100)):
        self.sc = _scene(fs)
        self.B = int(B)
        cfg = self.cfg = _EnvCfg()
        cfg.W, cfg.H, cfg.max_bounces, cfg.flavour = int(width), int(height), int(max_bounces), int(flavour == "fb")
        cfg.cam[:] = [float(x) for x in camera]
        cfg.cam_angle[:] = [float(x) for x in camera_angle]
        cfg.fov, cfg.sun_id = float(fov), int(sun_id)
        cfg.adaptive = int(bool(adaptive))
        cfg.light_ids[:] = [int(light_ids[0]), int(light_ids[1])]
        self.state = np.zeros(self.B * lib().orc_sizeof_env(), np.uint8)

    def reset(self, pixels):
        pixels = np.ascontiguousarray(pixels, np.int32).reshape(self.B, 2)
        obs = np.zeros((self.B, 18), np.float32)
        lib().orc_env_reset(self.sc.ref, C.byref(self.cfg), self.B, _p(pixels, c_ip), self.state.ctypes.data, _p(obs, c_fp))
        return obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.B, 2)
        obs = np.zeros((self.B, 18), np.float32)
        rew = np.zeros(self.B)
        term = np.zeros(self.B, np.uint8)
        trunc = np.zeros(self.B, np.uint8)
        reason = np.zeros(self.B, np.int32)
        lib().orc_env_step(self.sc.ref, C.byref(self.cfg), self.B, _p(a, c_fp), self.state.ctypes.data, _p(obs, c_fp),
                           _p(rew, c_dp), _p(term, c_u8p), _p(trunc, c_u8p), _p(reason, c_ip))
        return obs, rew, term.astype(bool), trunc.astype(bool), reason


def max_threads():
    return int(lib().orc_max_threads())
