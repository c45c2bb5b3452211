import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import ray_tracer_v1_b200 as pkg
from ray_tracer_v1_b200 import scenes, _native as nat
spec = scenes.build_complex()
fs = pkg.flatten_scene(spec.spheres, background_colour=spec.background)
sc = nat.DeviceScene(fs)
W, H, spp, world = 256, 144, 4, 3
dev = torch.device("cuda", 0)
images = [torch.zeros((H, W, 3), dtype=torch.float32, device=dev) for _ in (0, 1)]
flags = [torch.zeros(nat.FLAG_WORDS, dtype=torch.int32, device=dev) for _ in range(world)]
timed_out = torch.zeros(1, dtype=torch.int32, device=dev)
streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
print("streams", [s.cuda_stream for s in streams])
torch.cuda.synchronize()
evs = []
def launch(rank, e):
    sink = nat.PathSink()
    sink.mode, sink.tile_first, sink.tile_step = nat.SINK_IMAGE, rank, world
    sink.world, sink.rank, sink.sync, sink.epoch = world, rank, 1, e
    sink.go_epoch = e - 1 if rank == 0 else 0
    sink.image = images[e & 1].data_ptr()
    sink.timed_out, sink.timeout_ms, sink.max_ctas = timed_out.data_ptr(), 1500, 148
    for k in range(world):
        sink.flags[k] = flags[k].data_ptr()
    sc.render_path_sink(sc.path_params(spec.camera, W, H, spp, 5, 0.9, seed=300 + e), sink, stream=streams[rank].cuda_stream)
    ev = torch.cuda.Event(enable_timing=True); ev.record(streams[rank]); evs.append((rank, e, ev))
t0 = torch.cuda.Event(enable_timing=True); t0.record()
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
if mode == "all":
    for rank in (1, 2):
        for e in range(1, 5):
            launch(rank, e)
    for e in range(1, 5):
        launch(0, e)
        with torch.cuda.stream(streams[0]):
            torch.cuda._sleep(int(0.04 * 1.9e9))
            snap = images[e & 1].clone()
else:
    for rank in (1, 2):
        launch(rank, 1); launch(rank, 2)
    launch(0, 1)
    for e in range(1, 5):
        with torch.cuda.stream(streams[0]):
            torch.cuda._sleep(int(0.04 * 1.9e9))
            snap = images[e & 1].clone()
        if e + 1 <= 4: launch(0, e + 1)
        if e + 2 <= 4:
            for rank in (1, 2): launch(rank, e + 2)
torch.cuda.synchronize()
print("timed_out", int(timed_out.item()))
for k in range(world): print("flags", k, flags[k][:3].tolist(), flags[k][16:19].tolist(), int(flags[k][32]))
for rank, e, ev in evs: print(f"rank {rank} frame {e} done at {t0.elapsed_time(ev):9.2f} ms")
