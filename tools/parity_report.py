#!/usr/bin/env python
"""FP32 product build against the FP64 parity build (which tests/ hold bit-exact to the oracle and the reference-rendered
goldens), on frames far larger than the goldens.  Same Philox stream in both, so the two differ only where FP32 geometry
flips a hit or an int() truncation; beside it the seed-to-seed noise of the estimator itself (RMSE between two seeds of
the FP64 build), which is the scale the north star's "RMSE bound shrinking as 1/sqrt(spp)" refers to.

    python tools/parity_report.py > profiles/parity_r1.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import ray_tracer_v1_b200 as rtb
from ray_tracer_v1_b200 import _native as nat, scenes


def flat(spec):
    return rtb.flatten_scene(spec.spheres, spec.global_lights, spec.point_lights, background_colour=spec.background)


def path_rows(name, spec, W, H):
    sc = nat.DeviceScene(flat(spec))
    print(f"## Algorithm B, {name}, {W}x{H}, depth {spec.max_bounces} -- colour levels 0-255, mean over spp")
    print("spp | pixels differing > 1 level (FP32 vs FP64) | RMSE FP32 vs FP64 | max diff | RMSE seed A vs seed B (FP64) | "
          "rays FP32 / FP64")
    for spp in (1, 4, 16, 64):
        a = (spec.camera, W, H, spp, spec.max_bounces, spec.mirror_threshold)
        _, s64, st64 = sc.render_path_host(sc.path_params(*a, seed=7), nat.F64)
        _, s32, st32 = sc.render_path_host(sc.path_params(*a, seed=7), nat.F32)
        _, s64b, _ = sc.render_path_host(sc.path_params(*a, seed=8), nat.F64)
        m64, m32, m64b = s64[..., :3] / spp, s32[..., :3] / spp, s64b[..., :3] / spp
        d = np.abs(m32 - m64)
        print(f"{spp:3d} | {100.0 * (d.max(axis=2) > 1.0).mean():6.3f} % | {np.sqrt(np.mean(d ** 2)):7.4f} | {d.max():6.1f} | "
              f"{np.sqrt(np.mean((m64 - m64b) ** 2)):7.3f} | {int(st32[0])} / {int(st64[0])}")
    sc.close()
    print()


def whitted_rows():
    from ray_tracer_v1_b200.renderers import CustomSceneExperiment  # noqa: F401  (import check of the drop-in entry)
    print("## Algorithm A (deterministic, spp 1): FP32 vs FP64, 8-bit quantised image")
    print("scene | size | pixels differing by > 1/255 | pixels differing at all | hit index flips")
    for name, spec in (("balls_in_space", scenes.build_balls_in_space()), ("marbles4", scenes.build_marbles4()),
                       ("planets2", scenes.build_planets2())):
        sc = nat.DeviceScene(flat(spec))
        for (W, H) in ((320, 240), (1280, 720)):
            k = int(100 * min(W, H) / 601) * 0.01
            X, Y = np.linspace(-k, k, W), np.linspace(k, -k, H)
            depth = 1 if name == "balls_in_space" else 4
            p = sc.whitted_params(spec.camera, X, Y, spp=1, max_bounces=depth, miss=[spec.miss.r, spec.miss.g, spec.miss.b], seed=0)
            _, s64, h64, _ = sc.render_whitted_host(p, nat.F64)
            _, s32, h32, _ = sc.render_whitted_host(p, nat.F32)
            q64, q32 = np.clip(np.rint(s64[..., :3]), 0, 255), np.clip(np.rint(s32[..., :3]), 0, 255)
            d = np.abs(q64 - q32).max(axis=2)
            print(f"{name} | {W}x{H} | {100.0 * (d > 1).mean():6.3f} % | {100.0 * (d > 0).mean():6.3f} % | "
                  f"{100.0 * (h64 != h32).mean():6.3f} %")
        sc.close()
    print()


if __name__ == "__main__":
    nat.lib()
    print("# FP32 product build vs FP64 parity build on the B200 (tools/parity_report.py)\n")
    path_rows("complex scene (54 spheres, 3 lights)", scenes.build_complex(), 480, 270)
    path_rows("chandelier (29 spheres, 21 lights)", scenes.build_chandelier(), 480, 270)
    try:
        whitted_rows()
    except Exception as e:      # noqa: BLE001 - the report is best effort for Algorithm A helpers
        print(f"(Algorithm A section skipped: {e})")
