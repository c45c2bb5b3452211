#!/usr/bin/env python
"""Wall-clock the class-level drop-in calls a user of the reference makes (host objects in, host image out):
TraditionalRenderer.render (Algorithm B, BASELINE config 3), CustomSceneExperiment.render_custom_scene (Algorithm A,
BASELINE config 1), SimplifiedFBRenderer.render_original_style (output6).  Development aid."""
import cProfile
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import ray_tracer_v1_b200 as rtb
from ray_tracer_v1_b200 import scenes
from ray_tracer_v1_b200.renderers import ComplexTraditionalRenderer, CustomSceneExperiment, SimplifiedFBRenderer


def best(fn, n):
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return min(ts), sorted(ts)[len(ts) // 2]


spec = scenes.build_complex()
r = ComplexTraditionalRenderer(seed=1)
r.scene = spec.spheres
r.light_sources = [s for s in spec.spheres if s.material.emitive]
r.small_lights = [s for s in r.light_sources if s.radius < 0.5]
r.camera_position = rtb.Vector(*spec.camera)
r.render(1920, 1080, 64, 5)
lo, med = best(lambda: r.render(1920, 1080, 64, 5), 5)
print(f"ComplexTraditionalRenderer.render(1920,1080,64,5): best {lo * 1e3:.2f} ms, median {med * 1e3:.2f} ms (view of the pinned ring)")
r.reuse_output = False
lo, med = best(lambda: r.render(1920, 1080, 64, 5), 5)
print(f"  reuse_output=False (fresh array per render): best {lo * 1e3:.2f} ms, median {med * 1e3:.2f} ms")
lo, med = best(lambda: r.render(320, 240, 4, 5), 20)
print(f"ComplexTraditionalRenderer.render(320,240,4,5): best {lo * 1e6:.0f} us, median {med * 1e6:.0f} us")

balls = scenes.build_balls_in_space(as_rendered=False).spheres
exp = CustomSceneExperiment(output_dir=tempfile.mkdtemp())
exp.config.update(image_width=320, image_height=240, samples_per_pixel=1, max_bounces=1)
exp.render_custom_scene(balls, 'traditional')
lo, med = best(lambda: exp.render_custom_scene(balls, 'traditional'), 200)
print(f"CustomSceneExperiment.render_custom_scene(balls_in_space, 320x240 spp 1) [C1]: best {lo * 1e6:.0f} us, median {med * 1e6:.0f} us")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    exp.render_custom_scene(balls, 'traditional')
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

fb = SimplifiedFBRenderer(seed=1)
fb.render_original_style(400, 300, output_path="")
lo, med = best(lambda: fb.render_original_style(400, 300, output_path=""), 50)
print(f"SimplifiedFBRenderer.render_original_style(400,300): best {lo * 1e6:.0f} us, median {med * 1e6:.0f} us")
