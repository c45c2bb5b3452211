#!/usr/bin/env python
"""Steady-state launches of the fused env step (rt_env_step_auto) of one flavour, for an ncu capture:
   ncu --set full --import-source on -k regex:env_step_kernel --launch-skip 30 --launch-count 1 python tools/debug/env_one.py rl"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from ray_tracer_v1_b200 import scenes, flatten_scene
from ray_tracer_v1_b200.ray_tracer_env import BatchedRayTracerEnv

flavour = sys.argv[1] if len(sys.argv) > 1 else "rl"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
if flavour == "fb":
    spec = scenes.build_balls_in_space(as_rendered=False)
    fs = flatten_scene(spec.spheres, spec.global_lights, [], spec.background)
    kw = dict(image_width=800, image_height=600, camera_position=(0, 0, 1), fov=90, max_bounces=5, flavour="fb")
else:
    spec = scenes.build_optimized_env_scene()
    fs = flatten_scene(spec.spheres, spec.global_lights, spec.point_lights, spec.background)
    kw = dict(image_width=320, image_height=240, camera_position=(0, 0, 0), fov=80, max_bounces=6, flavour="rl")
env = BatchedRayTracerEnv(fs, B, seed=1, **kw)
lo = torch.as_tensor(env.action_space.low, device="cuda")
hi = torch.as_tensor(env.action_space.high, device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)
acts = [lo + (hi - lo) * torch.rand((B, 2), device="cuda", generator=g) for _ in range(8)]
env.reset(seed=1)
for k in range(40):
    env.step_auto(acts[k % 8])
torch.cuda.synchronize()
print("ok", flavour, B)
