#!/usr/bin/env python
"""Rank 0's share of an 8-GPU frame (17 interleaved stripes, local image, no protocol) for different sample splits:
is the fixed ~80 us of a 2.2-ms launch the drain of the last work units?"""
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
for spp in (64, 8, 1, 0):
    for ks in (-1, 4, 8, 16, 32):
        if spp < 8 and ks != -1:
            continue
        p = sc.path_params(spec.camera, W, H, max(spp, 1), spec.max_bounces, spec.mirror_threshold, seed=1, ksplit=ks)
        if spp == 0:
            p.s1 = p.s0
        sink = nat.PathSink()
        sink.mode, sink.tile_first, sink.tile_step, sink.world = nat.SINK_IMAGE, 0, 8, 8
        sink.image = image.data_ptr()
        for _ in range(3):
            sc.render_path_sink(p, sink, stats=stats)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in ev:
            a.record(); sc.render_path_sink(p, sink, stats=stats); b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        print(f"17 stripes of 135, spp {spp}, ksplit {ks}: median {ts[10] * 1e3:.1f} us (work share of the full frame {17.27e3 * 17 / 135 * spp / 64:.1f} us)", flush=True)
