"""Multi-GPU sharding of frames: one process per GPU (torchrun), ``torch.distributed`` for the plumbing.

The path shards three ways; each needs at most ONE collective per frame:

* tiles    -- contiguous row bands, one per rank (C3: complex scene).  Every rank resolves its own band and rank 0
              receives the float32 bands with one NCCL gather straight into the rows of the final image (no staging).
* samples  -- sample ranges of every pixel, one per rank (C4: chandelier at high spp).  One NCCL reduce(sum) of the
              FP32 [H,W,4] accumulation buffers to rank 0, then the ``// spp`` resolve there.  Per-sample colours are
              integers, so FP32 sums are exact below 2^24 and the reduced frame is bit-identical to the unsharded one.
* env      -- environment slices (C5): no collective at all, see ray_tracer_env.BatchedRayTracerEnv.shard.

Because the Philox stream is keyed by the GLOBAL (pixel, sample, bounce), a sharded frame equals the unsharded frame
bit for bit for any world size.  The partition/collective helpers take the band renderer as a callable so the
world_size-2 ``gloo`` tests on CPU exercise exactly this code with a stand-in renderer.
"""
import numpy as np

__all__ = ["row_bands", "sample_ranges", "env_slices", "gather_row_bands", "reduce_sample_sums", "ShardedPathRenderer"]


def _split(total, world):
    return [(total * r // world, total * (r + 1) // world) for r in range(world)]


def row_bands(height, world):
    """Contiguous [y0, y1) row bands, sizes differing by at most one row."""
    return _split(int(height), int(world))


def sample_ranges(spp, world):
    """Contiguous [s0, s1) sample ranges (a rank may get an empty range when spp < world)."""
    return _split(int(spp), int(world))


def env_slices(n_envs, world):
    """Contiguous [b0, b1) environment slices."""
    return _split(int(n_envs), int(world))


def _dist():
    import torch.distributed as dist
    return dist


def _world(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gather_row_bands(image, bands, group=None, dst=0):
    """``image`` [H,W,C]: each rank has filled its own band ``bands[rank]``; after the call rank ``dst`` holds all rows.

    Equal bands: one gather whose receive buffers are the row slices of ``image`` itself.  Ragged bands: padded to
    the tallest band and copied in."""
    import torch
    dist = _dist()
    rank, world = _world(group)
    if world == 1:
        return image
    sizes = {b[1] - b[0] for b in bands}
    y0, y1 = bands[rank]
    dst_global = dist.get_global_rank(group, dst) if group is not None else dst
    if len(sizes) == 1:
        mine = image[y0:y1]
        if rank == dst:
            recv = [image[a:b] for a, b in bands]
            recv[rank] = torch.empty_like(mine)          # own band is already in place
            dist.gather(mine, recv, dst=dst_global, group=group)
        else:
            dist.gather(mine, None, dst=dst_global, group=group)
        return image
    tallest = max(sizes)
    pad = image.new_zeros((tallest,) + tuple(image.shape[1:]))
    pad[: y1 - y0] = image[y0:y1]
    if rank == dst:
        recv = [torch.empty_like(pad) for _ in bands]
        dist.gather(pad, recv, dst=dst_global, group=group)
        for r, (a, b) in enumerate(bands):
            if r != rank:
                image[a:b] = recv[r][: b - a]
    else:
        dist.gather(pad, None, dst=dst_global, group=group)
    return image


def reduce_sample_sums(accum, group=None, dst=0):
    """Sum the per-rank [H,W,4] accumulation buffers (r,g,b sums + sample count) onto rank ``dst``."""
    dist = _dist()
    rank, world = _world(group)
    if world == 1:
        return accum
    dst_global = dist.get_global_rank(group, dst) if group is not None else dst
    dist.reduce(accum, dst=dst_global, op=dist.ReduceOp.SUM, group=group)
    return accum


class ShardedPathRenderer:
    """Algorithm B frames on this rank's GPU, sharded over the process group by tiles or by samples.

    Buffers are torch CUDA tensors (so NCCL can move them); the kernels are launched through the C ABI on torch's
    current stream's device with the tensors' ``data_ptr()``."""

    def __init__(self, device=None, group=None, precision="f32"):
        import torch
        from . import _native as nat
        self.torch, self.nat = torch, nat
        self.group = group
        self.rank, self.world = _world(group)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.precision = nat.F64 if precision in ("f64", nat.F64) and precision != nat.F32 else nat.F32
        self.scene = None
        self._key = None
        self.h2d_bytes = self.d2h_bytes = 0
        self.launches = 0

    def set_scene(self, fs, lbvh=None):
        nat = self.nat
        if self.scene is None:
            self.scene = nat.DeviceScene(fs, self.device)
        else:
            self.scene.update(fs)
        n, nG, nP, nL = fs.radius.shape[0], fs.g_strength.shape[0], fs.p_strength.shape[0], fs.l_index.shape[0]
        self.h2d_bytes = (16 + 32) * (3 * n + 2 * (nG + nP + nL)) + 2 * 4 * (n + nG + 2 * nP + nL) + n
        if lbvh or (lbvh is None and self.scene.n > 256):
            self.scene.build_lbvh()

    def _ensure(self, W, H):
        torch = self.torch
        if self._key != (W, H):
            dev = torch.device("cuda", self.device)
            ft = torch.float64 if self.precision == self.nat.F64 else torch.float32
            self.accum = torch.zeros((H, W, 4), dtype=ft, device=dev)
            self.image = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
            self.stats = torch.zeros(8, dtype=torch.int64, device=dev)       # uint64 counters, read as int64
            self.host_image = torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True)
            self._key = (W, H)

    def render(self, cam, W, H, spp, max_bounces, mirror_threshold, seed=0, fov=60.0, mode="tiles", to_host=False):
        """Render one frame cooperatively.  Returns (image, stats) on rank 0 -- image a CUDA tensor [H,W,3] float32, or
        a numpy view of pinned host memory when ``to_host`` -- and (None, stats) elsewhere.  stats = this rank's
        uint64[8] counter block as a CUDA int64 tensor (not reduced: callers sum what they need)."""
        torch, nat, sc = self.torch, self.nat, self.scene
        self._ensure(W, H)
        rows = (0, H)
        samples = (0, spp)
        if mode == "tiles":
            bands = row_bands(H, self.world)
            rows = bands[self.rank]
        elif mode == "samples":
            samples = sample_ranges(spp, self.world)[self.rank]
        else:
            raise ValueError("mode must be 'tiles' or 'samples'")
        self.stats.zero_()
        if mode == "samples" and samples[0] == samples[1]:
            self.accum.zero_()
        p = sc.path_params(cam, W, H, spp, max_bounces, mirror_threshold, seed=seed, fov=fov, rows=rows, samples=samples)
        sc.render_path(p, self.accum, self.precision, stats=self.stats)
        self.launches = 1
        if mode == "tiles":
            sc.resolve(self.accum, W, H, spp, self.image, self.precision, rows=rows)
            self.launches += 1
            gather_row_bands(self.image, bands, self.group)
        else:
            reduce_sample_sums(self.accum, self.group)
            if self.rank == 0:
                sc.resolve(self.accum, W, H, spp, self.image, self.precision)
                self.launches += 1
        if self.rank != 0:
            return None, self.stats
        if to_host:
            self.host_image.copy_(self.image, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self.d2h_bytes = self.host_image.numel() * 4
            return self.host_image.numpy(), self.stats
        return self.image, self.stats

    def close(self):
        if self.scene is not None:
            self.scene.close()
            self.scene = None
