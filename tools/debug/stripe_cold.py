#!/usr/bin/env python
"""One GPU, 1/8 of the headline frame (an interleaved share is emulated by a 135-row band at the image centre): the frame
back to back against the frame after an L2 flush, each frame inside its own event pair -- what a flush between steps costs
a 2.3-ms step."""
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
accum = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
stats = torch.zeros(8, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rows in ((0, 1080), (472, 607), (0, 135)):
    p = sc.path_params(spec.camera, W, H, 64, spec.max_bounces, spec.mirror_threshold, seed=1, rows=rows)
    for name, pre in (("back to back", lambda: None), ("after flush.zero_()", lambda: flush.zero_()),
                      ("after a 4 MB memset", lambda: flush[:4 << 20].zero_())):
        for _ in range(3):
            sc.render_path(p, accum, nat.F32, stats=stats)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in ev:
            pre()
            a.record(); sc.render_path(p, accum, nat.F32, stats=stats); b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        print(f"rows {rows}: {name}: median {ts[10]:.4f} ms, min {ts[0]:.4f}, max {ts[-1]:.4f}", flush=True)

# the share of rank r of 8 (interleaved 8-row stripes) through the fused sink, stores into a LOCAL image, no protocol:
# what the per-rank kernel of an 8-GPU frame costs before any NVLink traffic or flag is involved
image = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
for world in (8, 2):
    for r in (0, world - 1):
        p = sc.path_params(spec.camera, W, H, 64, spec.max_bounces, spec.mirror_threshold, seed=1)
        sink = nat.PathSink()
        sink.mode, sink.tile_first, sink.tile_step, sink.world = nat.SINK_IMAGE, r, world, world
        sink.image = image.data_ptr()
        for _ in range(3):
            sc.render_path_sink(p, sink, stats=stats)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in ev:
            flush.zero_()
            a.record(); sc.render_path_sink(p, sink, stats=stats); b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        stripes = len(range(r, 135, world))
        print(f"stripes of rank {r} of {world} ({stripes} of 135 = {17.27 * stripes / 135:.3f} ms of the 17.27-ms frame), local image, "
              f"no protocol: median {ts[10]:.4f} ms, min {ts[0]:.4f}", flush=True)
