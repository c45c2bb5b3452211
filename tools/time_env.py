#!/usr/bin/env python
"""Device- and wall-time the batched env step at BASELINE config 5 (65,536 envs): the two-launch protocol (step +
masked reset), the fused single launch (rt_env_step_auto) and its CUDA-graph replay.  Development aid."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ray_tracer_v1_b200 import scenes, flatten_scene
from ray_tracer_v1_b200.ray_tracer_env import BatchedRayTracerEnv


def case(flavour, B=65536, steps=300):
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

    def two_launch(n):
        for k in range(n):
            _, _, te, tr, _ = env.step(acts[k % 8])
            env.reset(mask=(te | tr).to(torch.uint8))

    def fused(n, graph=False):
        for k in range(n):
            env.actions.copy_(acts[k % 8]) if False else None
            env.step_auto(acts[k % 8] if not graph else None, graph=graph)

    out = {}
    env.step_auto(acts[0])
    for name, fn in (("step + masked reset (2 launches)", two_launch), ("step_auto (1 launch + action copy)", fused),
                     ("step_auto graph replay (actions in place)", lambda n: fused(n, True))):
        fn(20)
        torch.cuda.synchronize()
        env.stats.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(); fn(steps); b.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev = a.elapsed_time(b)
        q = int(env.stats.cpu()[4])
        print(f"{flavour} B={B} {name}: {1e3 * dev / steps:.2f} us/step device, {1e6 * wall / steps:.2f} us/step wall, "
              f"{B * steps / (dev * 1e-3) / 1e9:.2f} G env-steps/s, {q / dev / 1e3:.0f} Mrays/s", flush=True)
    env.close()


if __name__ == "__main__":
    for f in ("fb", "rl"):
        case(f)
