#!/usr/bin/env python
"""Every configuration of BASELINE.json on ONE B200, device-timed, with the CPU port (oracle, all host threads) timed
beside it on a bounded sample of the same workload.  One JSON line per configuration on stdout.

    C1  balls_in_space 320x240, spp 1, Algorithm A (render_custom_scene grid), depth 1
    C2  marbles4 / planets2 1280x720, 16 spp, Algorithm A, depth 4
    C3  complex 1920x1080, spp 1/16/64/256, Algorithm B, depth 5
    C4  chandelier 1920x1080 64 spp depth 8 (brute force and LBVH) + the 1e3/1e4/1e5-sphere LBVH variants
    C5  batched RayTracerEnv.step, 65,536 envs, FB flavour on balls_in_space and RL flavour on the optimized scene

bench.py stays the headline (C3 @ 64 spp); this is the per-shape table of SURVEY.md 8(d).  A ray = one nearest-hit or
occlusion query.  FLOP per query: brute force 20 N + shade; LBVH 20 <sphere tests> + 24 <box tests> + shade,
from the device counters.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ray_tracer_v1_b200 as rtb  # noqa: E402
from ray_tracer_v1_b200 import _native as nat, scenes  # noqa: E402
from ray_tracer_v1_b200.ray_tracer_env import BatchedRayTracerEnv  # noqa: E402


def flat(spec):
    return rtb.flatten_scene(spec.spheres, spec.global_lights, spec.point_lights, background_colour=spec.background)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def emit(name, **kw):
    print(json.dumps({"config": name, **kw}), flush=True)


def host_threads():
    from oracle import oracle as orc
    return max(1, orc.max_threads())


# ------------------------------------------------------------------------------------------------ Algorithm A
def bench_whitted(name, spec, W, H, spp, depth, grid, peak, reps, cpu):
    fs = flat(spec)
    sc = nat.DeviceScene(fs)
    X, Y = grid
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    p = sc.whitted_params(spec.camera, X, Y, spp=spp, max_bounces=depth, miss=[spec.miss.r, spec.miss.g, spec.miss.b], seed=0)
    ms = timed(lambda: sc.render_whitted(p, accum, nat.F32, stats=stats), reps)
    stats.zero_()
    sc.render_whitted(p, accum, nat.F32, stats=stats)
    st = stats.cpu().numpy()
    q = int(st[4])
    n, nG, nP = fs.radius.shape[0], fs.g_strength.shape[0], fs.p_strength.shape[0]
    fpq = 20 * n + 15 + (25 * nG + 30 * nP) * (W * H * spp) / max(q, 1)      # shading is paid once per primary hit
    out = {"width": W, "height": H, "spp": spp, "max_bounces": depth, "spheres": n, "ms_per_frame": ms,
           "Mrays_per_s": q / ms / 1e3, "queries_per_pixel_sample": q / (W * H * spp),
           "roofline_frac": q * fpq / (ms * 1e-3) / 1e12 / peak, "flop_per_query": fpq}
    if cpu:
        from oracle import oracle as orc
        cw, ch, cs = cpu
        Xc, Yc = X[:: max(1, W // cw)], Y[:: max(1, H // ch)]
        t0 = time.perf_counter()
        _, _, qc = orc.render_whitted(fs, spec.camera, Xc, Yc, spp=cs, max_bounces=depth,
                                      miss=[spec.miss.r, spec.miss.g, spec.miss.b], seed=0, nthreads=host_threads())
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"Mrays_per_s": qc / dt / 1e6, "cores": host_threads(), "kind": "port",
                               "sample": f"{len(Xc)}x{len(Yc)} x {cs} spp, {qc} queries, {dt:.2f} s"}
    sc.close()
    emit(name, **out)


# ------------------------------------------------------------------------------------------------ Algorithm B
def bench_path(name, fs, cam, W, H, spp, depth, thr, peak, reps, lbvh=False, cpu=None):
    sc = nat.DeviceScene(fs)
    if lbvh:
        sc.build_lbvh(50.0)
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    ms = 0.0
    st = np.zeros(8, np.int64)
    chunk = 64                       # uint32 per-launch accumulators: <= 65536 samples per launch; 64 keeps launches short
    p_list = [sc.path_params(cam, W, H, spp, depth, thr, seed=0, samples=(s0, min(spp, s0 + chunk)), accumulate=s0 > 0)
              for s0 in range(0, spp, chunk)]

    def frame():
        for p in p_list:
            sc.render_path(p, accum, nat.F32, stats=stats)
    ms = timed(frame, reps)
    stats.zero_()
    frame()
    st = stats.cpu().numpy()
    q, tests, boxes = int(st[4]), int(st[5]), int(st[6])
    n, nL = fs.radius.shape[0], fs.l_index.shape[0]
    fpq = (20 * tests + 24 * boxes) / max(q, 1) + 15 + 25 * nL + 70
    out = {"width": W, "height": H, "spp": spp, "max_bounces": depth, "spheres": n, "lights": nL, "lbvh": bool(lbvh),
           "ms_per_frame": ms, "Mrays_per_s": q / ms / 1e3, "rays_ref_compatible_per_pixel_sample": int(st[0]) / (W * H * spp),
           "sphere_tests_per_query": tests / max(q, 1), "aabb_tests_per_query": boxes / max(q, 1),
           "roofline_frac": q * fpq / (ms * 1e-3) / 1e12 / peak, "flop_per_query": fpq}
    if cpu:
        from oracle import oracle as orc
        cw, ch, cs = cpu
        t0 = time.perf_counter()
        _, stc = orc.render_path(fs, cam, cw, ch, cs, depth, thr, seed=0, nthreads=host_threads())
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"Mrays_per_s": stc["queries"] / dt / 1e6, "cores": host_threads(), "kind": "port",
                               "sample": f"{cw}x{ch} x {cs} spp (brute force, as the reference), {stc['queries']} queries, {dt:.2f} s"}
    sc.close()
    emit(name, **out)


# ------------------------------------------------------------------------------------------------ env
def bench_env(name, spec, flavour, B, W, H, fov, depth, cam, reps, cpu_B):
    kw = dict(image_width=W, image_height=H, camera_position=rtb.Vector(*cam), fov=fov, max_bounces=depth,
              background_colour=spec.background, global_light_sources=spec.global_lights,
              point_light_sources=spec.point_lights, flavour=flavour)
    env = BatchedRayTracerEnv(spec.spheres, B, **kw)
    lo = torch.as_tensor(env.action_space.low, device="cuda")
    hi = torch.as_tensor(env.action_space.high, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = [lo + (hi - lo) * torch.rand((B, 2), device="cuda", generator=g) for _ in range(depth + 1)]

    def rollout():
        env.reset(seed=0)
        for a in acts:
            env.step(a)
    ms = timed(rollout, reps)
    env.stats.zero_()
    rollout()
    q = int(env.stats.cpu()[4])
    steps = B * len(acts)
    out = {"envs": B, "flavour": flavour, "steps_per_rollout": len(acts), "ms_per_rollout": ms,
           "env_steps_per_s": steps / (ms * 1e-3), "Mrays_per_s": q / ms / 1e3, "launches_per_rollout": 1 + len(acts)}
    # steady state of a vectorised trainer: every step is followed by a masked reset of the finished episodes, so all
    # B environments stay busy (2 launches per step)
    def busy(n):
        for k in range(n):
            _, _, te, tr, _ = env.step(acts[k % len(acts)])
            env.reset(mask=(te | tr).to(torch.uint8))
    env.reset(seed=1)
    busy(20)
    torch.cuda.synchronize()
    env.stats.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); busy(200); b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b)
    out["steady_state_two_launches"] = {"us_per_step_with_auto_reset": 1e3 * t / 200, "env_steps_per_s": B * 200 / (t * 1e-3)}
    # the same steady state through rt_env_step_auto: step + restart of the finished episodes in ONE launch, replayed from
    # a CUDA graph, actions written in place
    env.reset(seed=1)
    for k in range(20):
        env.step_auto(acts[k % len(acts)])
    env.actions.copy_(acts[0])
    for k in range(10):
        env.step_auto(None, graph=True)
    torch.cuda.synchronize()
    env.stats.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(200):
        env.step_auto(None, graph=True)
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b)
    out["steady_state"] = {"us_per_step_with_auto_reset": 1e3 * t / 200, "env_steps_per_s": B * 200 / (t * 1e-3),
                           "Mrays_per_s": int(env.stats.cpu()[4]) / t / 1e3, "launches_per_step": 1}
    env.close()
    if cpu_B:
        from oracle import oracle as orc
        fs = flat(spec)
        oe = orc.OracleEnv(fs, cpu_B, W, H, camera=cam, fov=fov, max_bounces=depth, flavour=flavour)
        rs = np.random.RandomState(0)
        pix = np.stack([rs.randint(0, W, cpu_B), rs.randint(0, H, cpu_B)], axis=1)
        a_np = [a[:cpu_B].cpu().numpy() for a in acts]
        t0 = time.perf_counter()
        oe.reset(pix)
        for a in a_np:
            oe.step(a)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"env_steps_per_s": cpu_B * len(acts) / dt, "cores": 1, "kind": "port",
                               "sample": f"{cpu_B} envs x {len(acts)} steps, {dt:.2f} s (serial, as the reference steps one env)"}
    emit(name, **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    want = lambda k: not args.only or any(k.startswith(o) for o in args.only.split(","))   # noqa: E731
    cpu = not args.no_cpu
    nat.lib()
    peak, _ = nat.measure_fp32_peak(0, 5)
    emit("fp32_peak", TFLOP_per_s=peak, source="rt_measure_fp32_peak (FFMA issue rate)")
    R = args.reps

    if want("C1"):
        bench_whitted("C1 balls_in_space 320x240 spp1 depth1 (Algorithm A)", scenes.build_balls_in_space(), 320, 240, 1, 1,
                      scenes.custom_scene_grid(320, 240), peak, R, (320, 240, 1) if cpu else None)
    if want("C2"):
        for nm, spec in (("marbles4", scenes.build_marbles4()), ("planets2", scenes.build_planets2())):
            k = 640 * spec.ray_step          # the notebooks' +-ray_count*ray_step window, widened to 16:9
            X, Y = np.linspace(-k * 16 / 9, k * 16 / 9, 1280), np.linspace(k, -k, 720)
            bench_whitted(f"C2 {nm} 1280x720 spp16 depth4 (Algorithm A)", spec, 1280, 720, 16, 4, (X, Y), peak, R,
                          (320, 180, 4) if cpu else None)
    if want("C3"):
        spec = scenes.build_complex()
        fs = rtb.flatten_scene(spec.spheres, background_colour=spec.background)
        for spp in (1, 16, 64, 256):
            bench_path(f"C3 complex 1920x1080 spp{spp} depth5 (Algorithm B)", fs, spec.camera, 1920, 1080, spp, 5, 0.9, peak,
                       R if spp <= 64 else 1, cpu=(480, 270, 8) if cpu and spp == 64 else None)
    if want("C4"):
        spec = scenes.build_chandelier()
        fs = rtb.flatten_scene(spec.spheres, background_colour=spec.background)
        bench_path("C4 chandelier 1920x1080 spp64 depth8 (Algorithm B, brute force)", fs, spec.camera, 1920, 1080, 64, 8, 0.0,
                   peak, R, cpu=(480, 270, 8) if cpu else None)
        bench_path("C4 chandelier 1920x1080 spp64 depth8 (Algorithm B, LBVH)", fs, spec.camera, 1920, 1080, 64, 8, 0.0, peak,
                   R, lbvh=True)
        for n_small in (1000, 10000, 100000):
            fsm = scenes.build_many_spheres_flat(n_small, seed=0)
            bench_path(f"C4 scaled {n_small} small spheres 1920x1080 spp16 depth8 (Algorithm B, LBVH)", fsm, (0.0, 2.0, 0.0), 1920,
                       1080, 16, 8, 0.0, peak, 1, lbvh=True, cpu=(96, 54, 2) if cpu and n_small == 1000 else None)
    if want("C5"):
        balls = scenes.build_balls_in_space(as_rendered=False)
        bench_env("C5 FB env 65536 envs balls_in_space", balls, "fb", 65536, 800, 600, 90, 5, (0.0, 0.0, 1.0), R,
                  65536 if cpu else 0)
        opt = scenes.build_optimized_env_scene()
        bench_env("C5 RL env 65536 envs optimized scene", opt, "rl", 65536, 320, 240, 80, 6, (0.0, 0.0, 0.0), R,
                  65536 if cpu else 0)


if __name__ == "__main__":
    main()
