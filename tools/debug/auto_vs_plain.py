#!/usr/bin/env python
"""FP32 fused step (rt_env_step_auto) against step + masked reset: which outputs part ways, and by how much."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from ray_tracer_v1_b200 import scenes, flatten_scene
from ray_tracer_v1_b200.ray_tracer_env import BatchedRayTracerEnv

spec = scenes.build_optimized_env_scene()
fs = flatten_scene(spec.spheres, spec.global_lights, spec.point_lights, spec.background)
kw = dict(image_width=320, image_height=240, camera_position=(0, 0, 0), fov=80, max_bounces=6, flavour="rl")
lo, hi = (0.0, 0.0), (np.pi / 2, 2 * np.pi)
B, T = 5000, 14
rs = np.random.RandomState(2)
acts = rs.uniform(lo, hi, (T, B, 2)).astype(np.float32)
fused = BatchedRayTracerEnv(fs, B, precision="f32", seed=9, **kw)
plain = BatchedRayTracerEnv(fs, B, precision="f32", seed=9, **kw)
fused.reset(seed=9)
plain.reset(options={"pixels": fused.pixels.clone()})
ok = torch.ones(B, dtype=torch.bool, device="cuda")
for t in range(T):
    obs, rew, term, trunc, info = fused.step_auto(acts[t])
    po, pr, pt, pu, pi = plain.step(acts[t])
    done = pt | pu
    flags = (term == pt) & (trunc == pu) & (info["reason"] == pi["reason"])
    dr = (rew.double() - pr.double()).abs()
    dto = (info["terminal_observation"] - po).abs().max(dim=1).values
    if bool(done.any()):
        plain.reset(mask=done.to(torch.uint8), options={"pixels": info["pixels"].clone()})
    do = (obs - plain.obs).abs().max(dim=1).values
    tol = float(os.environ.get("TOL", "2e-4"))
    bad = ok & (~flags | (dr > 1e-5) | (done & (dto > tol)) | (do > tol))
    print(f"step {t}: new drop-outs {int(bad.sum())}: flags {int((ok & ~flags).sum())} reward {int((ok & flags & (dr > 1e-5)).sum())} "
          f"(max {float(dr[ok].max()):.2e}) terminal_obs {int((ok & done & (dto > tol)).sum())} obs {int((ok & (do > tol)).sum())} "
          f"(max {float(do[ok].max()):.2e})")
    if int(bad.sum()):
        i = int(torch.nonzero(bad)[0])
        print("   e.g. env", i, "reward", float(rew[i]), float(pr[i]), "reason", int(info["reason"][i]), int(pi["reason"][i]),
              "\n   obs  ", obs[i].cpu().numpy().round(5), "\n   plain", plain.obs[i].cpu().numpy().round(5))
    ok &= ~bad
print("following", float(ok.float().mean()))
