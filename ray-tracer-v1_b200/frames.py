"""Persistent frame state on one GPU: scene handle, accumulation / image buffers in HBM and a pinned host image.

``FrameContext`` is what the render entry points and the multi-GPU sharding layer (distributed.py) share.  A frame is
rendered in three stream-ordered launches -- path/whitted kernel -> resolve kernel -> one D2H copy of the float32
image into page-locked memory -- and nothing else crosses PCIe except the few KB of flattened scene.
"""
import ctypes as C

import numpy as np

from . import _native as nat

__all__ = ["FrameContext"]


class FrameContext:
    HOST_RING = 3

    def __init__(self, device=0):
        self.device = device
        self.scene = None
        self._key = None
        self.accum = self.image = self.stats = self.hit = None
        self.host_image = self.host_stats = None
        self.h2d_bytes = 0          # bytes of the last scene upload
        self.d2h_bytes = 0          # bytes of the last read-back

    # ---- scene -----------------------------------------------------------------------------------------------
    def set_scene(self, fs, lbvh=None, skip_unchanged=False):
        """(Re)upload the flattened scene; allocations are reused when the census is unchanged.  ``skip_unchanged``: a
        scene whose flattened content equals the resident one is not uploaded again (``h2d_bytes`` = 0 for that frame)."""
        uploaded = True
        if self.scene is None:
            self.scene = nat.DeviceScene(fs, self.device)
            if skip_unchanged:
                self.scene._sig = nat.scene_signature(fs)
        else:
            uploaded = self.scene.update(fs, skip_unchanged=skip_unchanged)
        n, nG, nP, nL = fs.radius.shape[0], fs.g_strength.shape[0], fs.p_strength.shape[0], fs.l_index.shape[0]
        self.h2d_bytes = ((16 + 32) * (3 * n + 2 * (nG + nP + nL)) + 2 * 4 * (n + nG + 2 * nP + nL) + n) if uploaded else 0
        if uploaded and (lbvh or (lbvh is None and self.scene.n > 256)):
            self.scene.build_lbvh()
        return self.scene

    # ---- buffers ---------------------------------------------------------------------------------------------
    def ensure(self, W, H, precision=nat.F32, want_hit=False):
        key = (int(W), int(H), precision)
        if self._key != key:
            for b in [self.accum, self.image, self.stats, self.hit, self.host_stats] + list(getattr(self, "host_ring", [])):
                if b is not None:
                    b.free()
            ft = np.float64 if precision == nat.F64 else np.float32
            self.accum = nat.DeviceBuffer((H, W, 4), ft, self.device)
            self.image = nat.DeviceBuffer((H, W, 3), np.float32, self.device)
            self.stats = nat.DeviceBuffer(8, np.uint64, self.device)
            # a ring of pinned host images: a frame handed out by read_back stays untouched until HOST_RING - 1 later
            # frames have been read back (TraditionalRenderer.render returns these views without copying them)
            self.host_ring = [nat.PinnedArray((H, W, 3), np.float32) for _ in range(self.HOST_RING)]
            self._ring_at = 0
            self.host_image = self.host_ring[0]
            self.host_stats = nat.PinnedArray(8, np.uint64)
            self.hit = None
            self._key = key
        if want_hit and self.hit is None:
            self.hit = nat.DeviceBuffer((H, W), np.int32, self.device)

    # ---- frames ----------------------------------------------------------------------------------------------
    def render_path(self, cam, W, H, spp, max_bounces, mirror_threshold, seed=0, fov=60.0, precision=nat.F32, rows=None,
                    samples=None, resolve=True, read_back=True, stream=None):
        """Algorithm B over rows x samples of this rank.  Returns (host image view or None, stats u64[8] or None)."""
        self.ensure(W, H, precision)
        L = nat.lib()
        p = self.scene.path_params(cam, W, H, spp, max_bounces, mirror_threshold, seed=seed, fov=fov, rows=rows,
                                   samples=samples)
        self.stats.fill(0, stream)
        bands = self._bands(W, p.y0, p.y1, p.s1 - p.s0) if (resolve and read_back) else None
        if bands:
            return self._render_path_banded(p, bands, W, H, precision, stream)
        self.scene.render_path(p, self.accum, precision, stats=self.stats, stream=stream)
        self.launches = 1
        if resolve:
            self.scene.resolve(self.accum, W, H, p.s1 - p.s0, self.image, precision, rows=(p.y0, p.y1), stream=stream)
            self.launches += 1
        if not read_back:
            return None, None
        return self.read_back(W, H, (p.y0, p.y1), stream)

    # a frame whose image takes longer to copy out than a launch costs is rendered in row bands: band b is copied to the
    # pinned host image (copy stream) while band b + 1 renders.  Pixels are keyed by (pixel, sample), so the frame is the
    # same bit for bit (tested); every extra launch costs ~30 us of drain, every band hides its share of the copy.  The
    # bands taper (BAND_WEIGHTS): a band's copy hides behind the NEXT band's render, so only the last band's copy is paid.
    # Only the LAST band's copy is exposed and every extra launch costs its drain, so: two bands, the last one 1/16 of the
    # rows.  Measured through ComplexTraditionalRenderer.render(1920, 1080, 64, 5) (tools/debug/bands_ab.py,
    # profiles/bands_ab_r2.txt): one launch 17.53 ms, four equal bands 17.33, (11,10,8,3) 17.25, (6,5,2) 17.23, (7,1) 17.20,
    # (15,1) 17.17, (23,1) 17.16 (kept at 15:1 -- the first band's 23 MB copy still hides at the smallest banded frames).
    BANDS = 2
    BAND_WEIGHTS = (15, 1)
    BAND_MIN_IMAGE_BYTES = 8 << 20
    BAND_MIN_SAMPLES = 24 << 20          # pixel-samples: below ~10 ms of kernel the split is not worth its launches

    def _bands(self, W, y0, y1, ns):
        rows = y1 - y0
        if self.BANDS < 2 or rows * W * 12 < self.BAND_MIN_IMAGE_BYTES or rows * W * ns < self.BAND_MIN_SAMPLES:
            return None
        stripes = (rows + 7) // 8
        w = self.BAND_WEIGHTS if len(self.BAND_WEIGHTS) == self.BANDS else (1,) * self.BANDS
        cuts = [y0 + 8 * ((stripes * sum(w[:k])) // sum(w)) for k in range(self.BANDS)] + [y1]
        return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]

    def _render_path_banded(self, p, bands, W, H, precision, stream):
        L = nat.lib()
        if getattr(self, "_copy_stream", None) is None:
            cs = nat.vp()
            nat.check(L.rt_stream_create(self.device, C.byref(cs)))
            self._copy_stream = cs
        cs = self._copy_stream
        self._ring_at = (self._ring_at + 1) % self.HOST_RING
        self.host_image = self.host_ring[self._ring_at]
        ns, y0, y1 = p.s1 - p.s0, p.y0, p.y1
        self.launches = 0
        for (a, b) in bands:
            p.y0, p.y1 = a, b
            self.scene.render_path(p, self.accum, precision, stats=self.stats, stream=stream)
            self.scene.resolve(self.accum, W, H, ns, self.image, precision, rows=(a, b), stream=stream)
            self.launches += 2
            nat.check(L.rt_stream_wait_stream(self.device, cs, stream))
            off, nbytes = a * W * 12, (b - a) * W * 12
            nat.check(L.rt_memcpy_d2h(self.device, self.host_image.ptr + off, self.image.ptr + off, nbytes, cs))
        p.y0, p.y1 = y0, y1
        nat.check(L.rt_memcpy_d2h(self.device, self.host_stats.ptr, self.stats.ptr, 64, stream))
        nat.check(L.rt_stream_sync(self.device, stream))
        nat.check(L.rt_stream_sync(self.device, cs))
        self.d2h_bytes = (y1 - y0) * W * 12 + 64
        return self.host_image.array, self.host_stats.array

    def render_whitted(self, cam, X, Y, spp=1, max_bounces=1, shadow_max_bounces=0, miss=None, seed=0, prenorm=False,
                       precision=nat.F32, rows=None, samples=None, read_back=True, stream=None):
        W, H = len(X), len(Y)
        self.ensure(W, H, precision)
        p = self.scene.whitted_params(cam, X, Y, spp=spp, max_bounces=max_bounces, shadow_max_bounces=shadow_max_bounces,
                                      miss=miss, seed=seed, prenorm=prenorm, rows=rows, samples=samples)
        self.stats.fill(0, stream)
        self.scene.render_whitted(p, self.accum, precision, stats=self.stats, stream=stream)
        self.scene.resolve(self.accum, W, H, p.s1 - p.s0, self.image, precision, rows=(p.y0, p.y1), stream=stream)
        self.launches = 2
        if not read_back:
            return None, None
        return self.read_back(W, H, (p.y0, p.y1), stream)

    def read_back(self, W, H, rows, stream=None):
        """D2H of the resolved rows (float32) + the stats block into pinned memory; synchronises the stream."""
        L = nat.lib()
        y0, y1 = rows
        self._ring_at = (self._ring_at + 1) % self.HOST_RING
        self.host_image = self.host_ring[self._ring_at]
        off, nbytes = y0 * W * 3 * 4, (y1 - y0) * W * 3 * 4
        if nbytes:
            nat.check(L.rt_memcpy_d2h(self.device, self.host_image.ptr + off, self.image.ptr + off, nbytes, stream))
        nat.check(L.rt_memcpy_d2h(self.device, self.host_stats.ptr, self.stats.ptr, 64, stream))
        nat.check(L.rt_stream_sync(self.device, stream))
        self.d2h_bytes = nbytes + 64
        return self.host_image.array, self.host_stats.array

    def close(self):
        for b in [self.accum, self.image, self.stats, self.hit, self.host_stats] + list(getattr(self, "host_ring", [])):
            if b is not None:
                b.free()
        self.host_ring = []
        if getattr(self, "_copy_stream", None) is not None:
            nat.lib().rt_stream_destroy(self.device, self._copy_stream)
            self._copy_stream = None
        self.accum = self.image = self.stats = self.hit = self.host_image = self.host_stats = None
        self._key = None
        if self.scene is not None:
            self.scene.close()
            self.scene = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
