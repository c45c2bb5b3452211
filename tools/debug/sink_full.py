#!/usr/bin/env python
"""The full headline frame through the fused sink (world 1, local image) against the accumulate + resolve path, and the
sum of the 8 / 2 interleaved shares: is it the sink variant or the striping that costs 3-4 % at 8 GPUs?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import ray_tracer_v1_b200 as rtb
from ray_tracer_v1_b200 import _native as nat, scenes

spec = scenes.build_complex()
fs = rtb.flatten_scene(spec.spheres, background_colour=spec.background)
sc = nat.DeviceScene(fs)
W, H = 1920, 1080
stats = torch.zeros(8, dtype=torch.int64, device="cuda")
image = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
accum = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")


def timed(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[n // 2]


p = sc.path_params(spec.camera, W, H, 64, spec.max_bounces, spec.mirror_threshold, seed=1)
print(f"accumulate kernel, full frame: {timed(lambda: sc.render_path(p, accum, nat.F32, stats=stats)):.4f} ms", flush=True)
for world in (1, 2, 4, 8):
    tot, parts = 0.0, []
    for r in range(world):
        sink = nat.PathSink()
        sink.mode, sink.tile_first, sink.tile_step, sink.world = nat.SINK_IMAGE, r, world, world
        sink.image = image.data_ptr()
        t = timed(lambda: sc.render_path_sink(p, sink, stats=stats), 6)
        tot += t; parts.append(t)
    print(f"sink, {world} interleaved share(s): sum {tot:.4f} ms, max {max(parts):.4f}, shares {' '.join(f'{v:.3f}' for v in parts)}", flush=True)
for world in (2, 4, 8):
    tot, parts = 0.0, []
    for r in range(world):
        sink = nat.PathSink()
        sink.mode, sink.tile_first, sink.tile_step, sink.world, sink.col_split = nat.SINK_IMAGE, r, world, world, 1
        sink.image = image.data_ptr()
        t = timed(lambda: sc.render_path_sink(p, sink, stats=stats), 6)
        tot += t; parts.append(t)
    print(f"sink, {world} shares, 2-D interleave (col_split): sum {tot:.4f} ms, max {max(parts):.4f}, shares {' '.join(f'{v:.3f}' for v in parts)}", flush=True)
