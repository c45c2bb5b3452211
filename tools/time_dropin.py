import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import ray_tracer_v1_b200 as rtb
from ray_tracer_v1_b200 import scenes
from ray_tracer_v1_b200.renderers import TraditionalRenderer
spec = scenes.build_complex()
r = TraditionalRenderer()
r.scene = spec.spheres
r.light_sources = [s for s in spec.spheres if s.material.emitive]
r.small_lights = [s for s in r.light_sources if s.radius < 0.5]
r.camera_position = rtb.Vector(*spec.camera)
if hasattr(r, 'mirror_threshold'): r.mirror_threshold = spec.mirror_threshold
for i in range(4):
    t = time.perf_counter(); img = r.render(1920, 1080, 64, 5); dt = time.perf_counter() - t
    print(f"TraditionalRenderer.render(1920,1080,64,5): {dt*1e3:.2f} ms", img.shape, img.dtype, float(img.mean()), r.stats.get('total_rays'))
for i in range(3):
    t = time.perf_counter(); img = r.render(320, 240, 4, 5); dt = time.perf_counter() - t
    print(f"TraditionalRenderer.render(320,240,4,5): {dt*1e3:.2f} ms")
