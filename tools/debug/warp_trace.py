#!/usr/bin/env python
"""When do the 4,736 persistent warps of a path-kernel launch enter and leave their work loop?  Needs the development
build `tools/build_variant.sh trace -DRT_TRACE_WARPS` (RT_B200_LIB=build/ab/librt_trace.so): every warp writes
(%globaltimer at loop entry, at loop exit) through PathSink.timed_out.  Prints the ramp and the drain of a launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import ray_tracer_v1_b200 as rtb
from ray_tracer_v1_b200 import _native as nat, scenes

spec = scenes.build_complex()
fs = rtb.flatten_scene(spec.spheres, background_colour=spec.background)
sc = nat.DeviceScene(fs)
W, H = 1920, 1080
stats = torch.zeros(8, dtype=torch.int64, device="cuda")
image = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
trace = torch.zeros((592 * 8, 4), dtype=torch.int64, device="cuda")
for world, spp in ((8, 64), (1, 64), (8, 8)):
    p = sc.path_params(spec.camera, W, H, spp, spec.max_bounces, spec.mirror_threshold, seed=1)
    sink = nat.PathSink()
    sink.mode, sink.tile_first, sink.tile_step, sink.world = nat.SINK_IMAGE, 0, world, world
    sink.image = image.data_ptr()
    sink.timed_out = trace.data_ptr()
    for _ in range(3):
        sc.render_path_sink(p, sink, stats=stats)
    torch.cuda.synchronize()
    trace.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); sc.render_path_sink(p, sink, stats=stats); b.record()
    torch.cuda.synchronize()
    t = trace.cpu().numpy().astype(np.int64)
    t = t[t[:, 0] > 0]
    t0, t1 = t[:, 0], t[:, 1]
    base, end = t0.min(), t1.max()
    busy = (t1 - t0).sum() / 1e3
    span = (end - base) / 1e3
    print(f"share 1/{world}, spp {spp}: {len(t)} warps, event time {a.elapsed_time(b) * 1e3:.1f} us, first entry -> last exit {span:.1f} us; "
          f"entries spread over {(t0.max() - base) / 1e3:.1f} us; exits: first {(t1.min() - base) / 1e3:.1f}, median {(np.median(t1) - base) / 1e3:.1f}, "
          f"p90 {(np.percentile(t1, 90) - base) / 1e3:.1f}, p99 {(np.percentile(t1, 99) - base) / 1e3:.1f}, last {span:.1f} us; "
          f"warp-time inside the loop {100 * busy / (span * len(t)):.1f} % of warps x span", flush=True)
    last_len = (t1 - t[:, 2]) / 1e3                      # duration of every warp's LAST unit
    units = t[:, 3]
    n_units = units.max() + 1
    for name, sel in (("coarse", units < 0.88 * n_units), ("late (fine if any)", units >= 0.88 * n_units)):
        if sel.any():
            print(f"   last unit {name}: {int(sel.sum())} warps, duration median {np.median(last_len[sel]):.1f} us, p90 {np.percentile(last_len[sel], 90):.1f}, "
                  f"max {last_len[sel].max():.1f}; exit median {(np.median(t1[sel]) - base) / 1e3:.1f}")
    late = np.sort(end - t1)[::-1]
    print("   idle before the end, per warp (us): mean", round(float((end - t1).mean()) / 1e3, 1), " of entry:", round(float((t0 - base).mean()) / 1e3, 1))
