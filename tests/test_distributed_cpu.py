"""world_size-2 `gloo` tests (CPU) of the multi-GPU sharding layer (ray-tracer-v1_b200/distributed.py).

The partition + collective code is exactly what runs over NCCL on the GPUs; here the band renderer is a stand-in (the
CPU oracle renders each rank's rows / sample range), so what is checked is: tiles gathered onto rank 0 == the unsharded
frame, per-rank sample sums reduced onto rank 0 == the unsharded sums, ragged bands, and env slices."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, H_override, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from oracle import oracle as orc
        from ray_tracer_v1_b200.distributed import row_bands, sample_ranges, gather_row_bands, reduce_sample_sums
        z, fs = load_golden("path_complex_48x27")
        W, H, spp = int(z["W"]), H_override or int(z["H"]), int(z["spp"])
        args = dict(max_bounces=int(z["max_bounces"]), mirror_threshold=float(z["mirror_threshold"]), seed=int(z["seed"]))
        # ---- tiles: each rank resolves its band, rank 0 gathers the float32 rows
        bands = row_bands(H, world)
        y0, y1 = bands[rank]
        sums, _ = orc.render_path(fs, z["cam"], W, H, spp, rows=(y0, y1), **args)
        image = torch.zeros((H, W, 3), dtype=torch.float32)
        image[y0:y1] = torch.from_numpy(orc.resolve(sums, spp))[y0:y1]
        gather_row_bands(image, bands)
        # ---- samples: each rank sums its sample range, rank 0 gets the reduced [H,W,4] buffer
        s0, s1 = sample_ranges(spp, world)[rank]
        part, _ = orc.render_path(fs, z["cam"], W, H, spp, samples=(s0, s1), **args)
        accum = torch.zeros((H, W, 4), dtype=torch.float32)
        accum[..., :3] = torch.from_numpy(part.astype(np.float32))
        accum[..., 3] = s1 - s0
        reduce_sample_sums(accum)
        if rank == 0:
            whole, _ = orc.render_path(fs, z["cam"], W, H, spp, **args)
            ok_tiles = np.array_equal(image.numpy(), orc.resolve(whole, spp))
            ok_samples = np.array_equal(accum[..., :3].numpy(), whole.astype(np.float32)) and bool((accum[..., 3] == spp).all())
            q.put((ok_tiles, ok_samples, bands))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:      # surface the failure in the parent
        q.put(("error", repr(e), None))
        raise


@pytest.mark.parametrize("H", [None, 25])      # 27 rows -> bands 13/14 (ragged); 25 rows -> 12/13 (ragged) too
def test_tiles_and_samples_compose_over_gloo(H):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, H, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
    assert res[0] is True and res[1] is True, res
    assert all(p.exitcode == 0 for p in procs)


def test_partitions():
    from ray_tracer_v1_b200.distributed import row_bands, sample_ranges, env_slices
    for total, world in ((1080, 8), (1080, 7), (27, 2), (5, 8), (65536, 8), (64, 3)):
        for f in (row_bands, sample_ranges, env_slices):
            parts = f(total, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert row_bands(1080, 8)[3] == (405, 540)


def test_tile_stripes_cover_the_frame_once():
    """Fused tile mode: the interleaved 8-row stripes of all ranks partition the rows; the owner-band lookup of the
    scatter-add sink (rt_kernels.cuh: start from y*world//H, then walk) lands in the band that holds the row."""
    from ray_tracer_v1_b200.distributed import tile_stripes, row_bands
    for H, world in ((1080, 8), (1080, 3), (27, 2), (7, 4), (600, 16)):
        rows = sorted(y for r in range(world) for a, b in tile_stripes(H, world, r) for y in range(a, b))
        assert rows == list(range(H))
        bands = row_bands(H, world)
        bound = [b[0] for b in bands] + [H]
        for y in range(H):
            k = min(world - 1, y * world // H)
            while y >= bound[k + 1]:
                k += 1
            while y < bound[k]:
                k -= 1
            assert bands[k][0] <= y < bands[k][1]
