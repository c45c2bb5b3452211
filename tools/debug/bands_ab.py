#!/usr/bin/env python
"""A/B of the read-back band layout of FrameContext (BANDS / BAND_WEIGHTS) through ComplexTraditionalRenderer.render
(1920x1080, 64 spp, depth 5): wall clock per frame, host objects in, host image out.  Development aid."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import ray_tracer_v1_b200 as rtb
from ray_tracer_v1_b200 import scenes
from ray_tracer_v1_b200.frames import FrameContext
from ray_tracer_v1_b200.renderers import ComplexTraditionalRenderer

spec = scenes.build_complex()
r = ComplexTraditionalRenderer(seed=1)
r.scene = spec.spheres
r.light_sources = [s for s in spec.spheres if s.material.emitive]
r.small_lights = [s for s in r.light_sources if s.radius < 0.5]
r.camera_position = rtb.Vector(*spec.camera)
ref = r.render(1920, 1080, 64, 5).copy()
layouts = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]] or [(1, 1, 1, 1), (11, 10, 8, 3), (6, 5, 2), (7, 1), (15, 1), (1,)]
for rnd in range(2):
    for w in layouts:
        FrameContext.BANDS, FrameContext.BAND_WEIGHTS = len(w), w
        img = r.render(1920, 1080, 64, 5)
        assert np.array_equal(img, ref), w
        ts = []
        for _ in range(8):
            t = time.perf_counter(); r.render(1920, 1080, 64, 5); ts.append(time.perf_counter() - t)
        print(f"round {rnd} weights {w}: best {min(ts) * 1e3:.3f} ms, median {sorted(ts)[4] * 1e3:.3f} ms", flush=True)
