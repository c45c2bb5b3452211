"""GPU parity tests: the CUDA path (through the C ABI, csrc/librt_b200.so) against the committed golden vectors
(produced by the UNMODIFIED reference, oracle/gen_golden.py) and against the CPU oracle on seeded inputs.

Tolerances (north_star):
  * FP64 parity build: <= 1e-9 relative on float quantities; integer-valued colours / indices / counters exact.
  * FP32 product path, deterministic frames: <= 1/255 per channel after 8-bit quantisation, except the handful of
    silhouette / shadow-edge pixels where an FP32 hit/miss decision flips (bounded below as a fraction of the frame).
  * FP32 stochastic frames: same Philox stream as the oracle, so most pixels agree exactly; the mean image must agree
    within RMSE <= 4/sqrt(spp) levels (it is far smaller in practice).
"""
import numpy as np
from types import SimpleNamespace
import pytest

from conftest import gate, load_golden

pytestmark = pytest.mark.gpu

WHITTED = ["whitted_c1_balls_320x240", "whitted_balls_true_original_121", "whitted_marbles4_d4_121",
           "whitted_marbles4_d8_121", "whitted_planets2_d4_121", "whitted_planets2_d10_121"]


@pytest.fixture(scope="module")
def nat(rt):
    from ray_tracer_v1_b200 import _native
    return _native


def quant8(rgb):
    return np.clip(np.rint(rgb), 0, 255)


# ------------------------------------------------------------------ unit level
def test_sphere_discriminant_kat(nat):
    z, _ = load_golden("kat")
    d = z["disc"]
    for point in (0, 1):
        rows = d[d[:, 10] == point]
        out64 = nat.sphere_discriminant(rows[:, 0:6], rows[:, 6:10], point, nat.F64)
        assert np.array_equal(out64[:, 0], rows[:, 11])
        hit = rows[:, 11] == 1
        np.testing.assert_allclose(out64[hit, 1:8], rows[hit, 12:19], rtol=1e-9, atol=1e-12)
        out32 = nat.sphere_discriminant(rows[:, 0:6], rows[:, 6:10], point, nat.F32)
        agree = out32[:, 0] == rows[:, 11]
        gate("sphere_discriminant FP32 hit/miss flips (grazing rays)", 1 - agree.mean(), 1.5 / agree.size)        # measured 0: one KAT of slack
        both = hit & agree
        np.testing.assert_allclose(out32[both, 1:8], rows[both, 12:19], rtol=2e-3, atol=2e-4)


def test_trace_rays_matches_oracle(nat, orc):
    """Batched Ray.nearestSphereIntersect + terminalRGB on random rays, every scene, suppress ids and initial bounces."""
    rs = np.random.RandomState(5)
    for name, depth in (("whitted_c1_balls_320x240", 3), ("whitted_marbles4_d8_121", 8), ("whitted_planets2_d10_121", 10)):
        z, fs = load_golden(name)
        m = 4096
        org = np.tile(z["cam"], (m, 1)) + rs.uniform(-0.2, 0.2, (m, 3))
        dirs = np.stack([rs.uniform(-0.6, 0.6, m), rs.uniform(-0.6, 0.6, m), -np.ones(m)], 1)
        rays = np.concatenate([org, dirs], 1)
        sup = np.where(rs.rand(m) < 0.3, fs.ids[rs.randint(0, fs.n, m)], nat.NO_ID).astype(np.int32)
        b0 = rs.randint(0, 3, m).astype(np.int32)
        term_o, rgb_o = orc.trace_rays(fs, rays, suppress=sup, bounces0=b0, max_bounces=depth, miss=z["miss"])
        sc = nat.DeviceScene(fs)
        term, rgb = sc.trace_rays(rays, suppress=sup, bounces0=b0, max_bounces=depth, miss=z["miss"], precision=nat.F64)
        assert np.array_equal(term[:, :4], term_o[:, :4])
        np.testing.assert_allclose(term[:, 4:], term_o[:, 4:], rtol=1e-9, atol=1e-12)
        assert np.array_equal(rgb, rgb_o)
        # Intersection.terminalRGB at given hits
        hit = term_o[:, 0] == 1
        hits = np.concatenate([term_o[hit, 1:2], term_o[hit, 4:10]], 1)
        assert np.array_equal(sc.shade_hits(hits, 0, nat.F64), orc.shade_hits(fs, hits, 0))
        term32, rgb32 = sc.trace_rays(rays, suppress=sup, bounces0=b0, max_bounces=depth, miss=z["miss"], precision=nat.F32)
        same = np.all(term32[:, :2] == term_o[:, :2], axis=1)
        gate("trace_rays FP32 terminal object differs", 1 - same.mean(), max(1.5 / same.size, 1e-4))            # measured 0
        gate("trace_rays FP32 colour beyond 1 level", 1 - (np.abs(quant8(rgb32[same]) - quant8(rgb_o[same])).max(axis=1) <= 1).mean(), max(1.5 / same.size, 1e-4))
        sc.close()


# ------------------------------------------------------------------ Algorithm A frames
@pytest.mark.parametrize("name", WHITTED)
def test_whitted_frames_fp64_exact(nat, name):
    z, fs = load_golden(name)
    sc = nat.DeviceScene(fs)
    p = sc.whitted_params(z["cam"], z["X"], z["Y"], spp=1, max_bounces=int(z["max_bounces"]), miss=z["miss"],
                          prenorm=bool(z["prenorm"]))
    image, sums, hit, stats = sc.render_whitted_host(p, nat.F64)
    assert np.array_equal(hit, z["hit"].astype(np.int32))
    assert np.array_equal(sums[..., :3], z["rgb"].astype(np.float64))     # integer-valued colours: exact
    assert np.all(sums[..., 3] == 1)
    if "image" in z:
        assert np.array_equal(image, z["image"])
    H, W = hit.shape
    assert stats[0] == H * W and stats[4] >= H * W
    sc.close()


@pytest.mark.parametrize("name", WHITTED)
def test_whitted_frames_fp32_within_1_of_255(nat, name):
    z, fs = load_golden(name)
    sc = nat.DeviceScene(fs)
    p = sc.whitted_params(z["cam"], z["X"], z["Y"], spp=1, max_bounces=int(z["max_bounces"]), miss=z["miss"],
                          prenorm=bool(z["prenorm"]))
    image, sums, hit, _ = sc.render_whitted_host(p, nat.F32)
    ref = z["rgb"].astype(np.float64)
    diff = np.abs(quant8(sums[..., :3]) - quant8(ref)).max(axis=2)
    flipped = hit != z["hit"].astype(np.int32)
    bad = (diff > 1)
    # every pixel whose terminal object agrees must be within one 8-bit level, bar shadow-edge flips
    gate(f"{name} FP32 pixels beyond 1/255", bad.mean(), max(1.5 / bad.size, 1e-4))              # measured 0 on all six frames
    gate(f"{name} FP32 hit flips", flipped.mean(), max(3.5 / bad.size, 1e-4))                    # measured <= 1 pixel of 14,641
    if name == "whitted_c1_balls_320x240":
        # BASELINE config 1, the reference's own render_custom_scene frame: the north star's "within 1/255 per channel
        # after 8-bit quantisation" holds on every one of the 76,800 pixels (measured: 0; one pixel of slack)
        assert bad.sum() <= 1 and flipped.sum() <= 1, (int(bad.sum()), int(flipped.sum()))
    sc.close()


def test_whitted_jittered_spp4(nat):
    """render_custom_scene with spp > 1: Philox jitter keyed (pixel, sample) reproduces the reference run sample for sample."""
    z, fs = load_golden("whitted_balls_spp4_80x60")
    sc = nat.DeviceScene(fs)
    spp = int(z["spp"])
    p = sc.whitted_params(z["cam"], z["X"], z["Y"], spp=spp, max_bounces=int(z["max_bounces"]), miss=z["miss"],
                          seed=int(z["seed"]), prenorm=True)
    image, sums, _, _ = sc.render_whitted_host(p, nat.F64)
    assert np.array_equal(image, z["image"])
    image32, _, _, _ = sc.render_whitted_host(p, nat.F32)
    gate("whitted spp4 jittered FP32 pixels beyond 1/255", (np.abs(image32 - z["image"]).max(axis=2) > 1.01 / 255).mean(), 2.5 / image32[..., 0].size)   # measured 0
    sc.close()


# ------------------------------------------------------------------ Algorithm B frames
@pytest.mark.parametrize("name", ["path_chandelier_48x27", "path_complex_48x27"])
def test_path_frames(nat, name):
    z, fs = load_golden(name)
    W, H, spp = int(z["W"]), int(z["H"]), int(z["spp"])
    sc = nat.DeviceScene(fs)
    p = sc.path_params(z["cam"], W, H, spp, int(z["max_bounces"]), float(z["mirror_threshold"]), seed=int(z["seed"]))
    image, sums, stats = sc.render_path_host(p, nat.F64)
    assert list(stats[:4].astype(np.int64)) == list(z["stats"])
    assert np.array_equal(sums[..., :3], z["sums"].astype(np.float64))
    assert np.array_equal(image, z["image"])
    assert np.all(sums[..., 3] == spp)
    # FP32 product path: same Philox stream, so it only departs where FP32 geometry flips a hit or a truncation
    image32, sums32, stats32 = sc.render_path_host(p, nat.F32)
    d = np.abs(sums32[..., :3] / spp - z["sums"] / spp)
    gate(f"{name} FP32 pixels beyond one level", (d.max(axis=2) > 1.0).mean(), 6e-3)                     # measured 0 (full size: 2.5e-4 - 7.2e-4)
    rmse = float(np.sqrt(np.mean(d ** 2)))
    gate(f"{name} FP32 RMSE x sqrt(spp)", rmse * np.sqrt(spp), 1.5)                                   # measured 0.04-0.05 (full size 0.41-0.66)
    assert abs(int(stats32[0]) - int(z["stats"][0])) <= 0.01 * int(z["stats"][0])
    sc.close()


@pytest.mark.parametrize("precision", ["F32", "F64"])
def test_path_shards_compose(nat, precision):
    """Row bands x sample ranges accumulate to exactly the whole frame (what tile / sample sharding relies on)."""
    prec = getattr(nat, precision)
    z, fs = load_golden("path_complex_48x27")
    W, H, spp = int(z["W"]), int(z["H"]), 6
    sc = nat.DeviceScene(fs)
    ft = np.float64 if prec == nat.F64 else np.float32
    args = (z["cam"], W, H, spp, int(z["max_bounces"]), float(z["mirror_threshold"]))
    whole = nat.DeviceBuffer((H, W, 4), ft)
    st_whole = nat.DeviceBuffer(8, np.uint64)
    sc.render_path(sc.path_params(*args, seed=3), whole, prec, stats=st_whole)
    parts = nat.DeviceBuffer((H, W, 4), ft)
    st_parts = nat.DeviceBuffer(8, np.uint64)
    for rows in ((0, 11), (11, H)):
        for k, smp in enumerate(((0, 2), (2, 3), (3, spp))):
            sc.render_path(sc.path_params(*args, seed=3, rows=rows, samples=smp, accumulate=k > 0), parts, prec, stats=st_parts)
    assert np.array_equal(whole.download(), parts.download())
    a, b = st_whole.download(), st_parts.download()
    # slot 5 counts the sphere tests actually executed: the FP32 kernel culls the camera rays of a warp tile against the
    # tile's cone of rays, so that (and only that) counter depends on how the frame was cut into launches
    keep = [0, 1, 2, 3, 4, 6, 7] if prec == nat.F32 else list(range(8))
    assert np.array_equal(a[keep], b[keep])
    assert a[5] <= a[4] * len(fs.ids) and b[5] <= b[4] * len(fs.ids)
    sc.close()


def test_primary_candidate_lists_change_nothing(nat):
    """Camera rays are traced against the spheres their warp tile's cone can touch: same frame, fewer sphere tests."""
    z, fs = load_golden("path_complex_48x27")
    sc = nat.DeviceScene(fs)
    for (W, H, spp, ksplit) in ((160, 90, 8, -1), (160, 90, 8, 0), (97, 61, 5, 4)):
        args = (z["cam"], W, H, spp, int(z["max_bounces"]), float(z["mirror_threshold"]))
        _, with_lists, st_a = sc.render_path_host(sc.path_params(*args, seed=11, schedule=0, ksplit=ksplit), nat.F32)
        _, without, st_b = sc.render_path_host(sc.path_params(*args, seed=11, schedule=2, ksplit=ksplit), nat.F32)
        assert np.array_equal(with_lists, without)
        assert np.array_equal(st_a[:5], st_b[:5])
        assert st_b[5] == st_b[4] * len(fs.ids) and st_a[5] < 0.9 * st_b[5]
    sc.close()


def test_fold_table_equals_double_fold(nat):
    """Launches of >= 4 M pixel-samples of a scene whose colours are <= 255 fold through the per-CTA byte table
    int(albedo * (tot / 255.0)) (PathDev::fold_tab), smaller launches through the double product itself: the frame
    rendered whole (table) equals the same frame rendered in row bands below the threshold (double product) bit for bit.
    The complex scene (54 spheres) and a 60-sphere scene, whose table + static arrays need opt-in shared memory; a
    scene with a colour above 255 never uses the table and still folds exactly like the FP64 build's integers allow."""
    from ray_tracer_v1_b200 import scenes
    z, fs_c = load_golden("path_complex_48x27")
    W, H, spp = 512, 512, 16                                   # 4,194,304 pixel-samples: at the threshold
    for fs, cam, thr in ((fs_c, z["cam"], float(z["mirror_threshold"])),
                         (scenes.build_many_spheres_flat(54, seed=5, emissive_fraction=0.1), (0.0, 2.0, 0.0), 0.0)):
        assert float(np.max(fs.colour)) <= 255.0
        sc = nat.DeviceScene(fs)
        _, whole, st_w = sc.render_path_host(sc.path_params(cam, W, H, spp, 5, thr, seed=3), nat.F32)
        banded, st_b = np.zeros_like(whole), np.zeros(8, np.uint64)
        for y0 in range(0, H, 64):
            _, part, st = sc.render_path_host(sc.path_params(cam, W, H, spp, 5, thr, seed=3, rows=(y0, y0 + 64)), nat.F32)
            assert not part[:y0].any() and not part[y0 + 64:].any()
            banded += part
            st_b += st
        assert np.array_equal(whole, banded)
        assert np.array_equal(st_w[:5], st_b[:5])
        assert whole[..., :3].max() > 0 and len(np.unique(whole[..., 0])) > 100
        sc.close()


def test_parameter_block_loop_at_its_limits(nat):
    """64 spheres / 32 lights is the largest scene whose sphere and light pairs ride in the kernel parameter block; one
    more sphere falls back to the shared-memory loop.  Both sizes: FP32 against the FP64 parity build, and (64 spheres)
    the two FP32 loops against each other."""
    from ray_tracer_v1_b200 import scenes
    for n in (64, 65, 8, 57):
        fs = scenes.build_many_spheres_flat(n - 6, seed=5, emissive_fraction=0.45 if n == 64 else 0.1)
        assert len(fs.ids) == n, len(fs.ids)
        if n == 64:
            assert len(fs.l_index) <= 32
        sc = nat.DeviceScene(fs)
        W, H, spp = 96, 54, 4
        p = sc.path_params((0.0, 2.0, 0.0), W, H, spp, 4, 0.0, seed=9)
        _, f64, st64 = sc.render_path_host(p, nat.F64)
        _, f32, st32 = sc.render_path_host(p, nat.F32)
        assert np.abs(st32[:4].astype(np.int64) - st64[:4].astype(np.int64)).max() <= 0.01 * st64[0]
        assert ((np.abs(f32[..., :3] - f64[..., :3]) > spp).any(axis=2)).mean() < 0.03
        if n <= 64:
            p2 = sc.path_params((0.0, 2.0, 0.0), W, H, spp, 4, 0.0, seed=9, schedule=4)
            _, smem, st_s = sc.render_path_host(p2, nat.F32)
            assert (smem != f32).any(axis=2).mean() < 0.01
            assert st_s[5] == st_s[4] * n and st32[5] < st_s[5]       # only the parameter-block kernel has candidate lists
        sc.close()


def test_persistent_launches_rearm_their_counters(nat):
    """The path kernel's warps pull work from device counters that the launch itself re-arms: 150 launches (more than the
    64 counter pairs a scene rotates through), alternating between two streams without host synchronisation, all give
    the frame of a lone launch; so does a CUDA-graph replay (no memset node is needed)."""
    import torch
    z, fs = load_golden("path_complex_48x27")
    sc = nat.DeviceScene(fs)
    W, H, spp = 64, 36, 3
    args = (z["cam"], W, H, spp, int(z["max_bounces"]), float(z["mirror_threshold"]))
    ref = {}
    for seed in (1, 2):
        buf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        sc.render_path(sc.path_params(*args, seed=seed), buf, nat.F32)
        torch.cuda.synchronize()
        ref[seed] = buf.cpu().numpy()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    bufs = [torch.zeros((H, W, 4), dtype=torch.float32, device="cuda") for _ in range(150)]
    for k, buf in enumerate(bufs):
        sc.render_path(sc.path_params(*args, seed=1 + k % 2), buf, nat.F32, stream=streams[k % 2].cuda_stream)
    torch.cuda.synchronize()
    for k, buf in enumerate(bufs):
        assert np.array_equal(buf.cpu().numpy(), ref[1 + k % 2]), k
    g, out = torch.cuda.CUDAGraph(), torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    cap = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=cap):
        sc.render_path(sc.path_params(*args, seed=2), out, nat.F32, stream=torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ref[2])
    sc.close()


def test_path_schedules_agree(nat):
    """Lock-step and path-regeneration schedules are two orders of the same arithmetic: identical sums and counters."""
    for name in ("path_chandelier_48x27", "path_complex_48x27"):
        z, fs = load_golden(name)
        sc = nat.DeviceScene(fs)
        out = []
        for schedule in (0, 1):
            p = sc.path_params(z["cam"], 200, 120, 5, int(z["max_bounces"]), float(z["mirror_threshold"]), seed=77,
                               schedule=schedule)
            _, sums, stats = sc.render_path_host(p, nat.F32)
            out.append((sums, stats))
        # every counter but the executed sphere tests (slot 5: only the lock-step schedule has camera-ray candidate lists)
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1][:5], out[1][1][:5])
        assert out[0][1][5] <= out[1][1][5] == out[1][1][4] * len(fs.ids)
        sc.close()


def test_path_non_integer_colours_use_the_double_fold(nat, orc):
    """A scene whose colours are not integers takes the double-division fold; FP64 still equals the oracle exactly."""
    z, fs = load_golden("path_complex_48x27")
    fs.colour = fs.colour + 0.37
    fs.l_colour = fs.l_colour + 0.37
    W, H, spp = 64, 36, 3
    sums_o, st_o = orc.render_path(fs, z["cam"], W, H, spp, 5, 0.9, seed=4)
    sc = nat.DeviceScene(fs)
    _, sums, stats = sc.render_path_host(sc.path_params(z["cam"], W, H, spp, 5, 0.9, seed=4), nat.F64)
    assert np.array_equal(sums[..., :3], sums_o) and int(stats[0]) == st_o["total_rays"]
    sc.close()


def test_path_matches_oracle_larger(nat, orc):
    """Seeded 160x90 x 8 spp frames of both Algorithm-B scenes: FP64 CUDA == oracle exactly."""
    for name in ("path_chandelier_48x27", "path_complex_48x27"):
        z, fs = load_golden(name)
        W, H, spp = 160, 90, 8
        depth, thr = int(z["max_bounces"]), float(z["mirror_threshold"])
        sums_o, st_o = orc.render_path(fs, z["cam"], W, H, spp, depth, thr, seed=21)
        sc = nat.DeviceScene(fs)
        _, sums, stats = sc.render_path_host(sc.path_params(z["cam"], W, H, spp, depth, thr, seed=21), nat.F64)
        assert np.array_equal(sums[..., :3], sums_o)
        assert [int(x) for x in stats[:4]] == [st_o[k] for k in ("total_rays", "total_intersections", "light_hits", "small_light_hits")]
        sc.close()


# ------------------------------------------------------------------ LBVH
def test_lbvh_equals_brute_force(nat):
    from ray_tracer_v1_b200 import scenes
    fs = scenes.build_many_spheres_flat(3000, seed=1)
    sc = nat.DeviceScene(fs)
    W, H, spp = 128, 72, 2
    for prec, ft in ((nat.F32, np.float32), (nat.F64, np.float64)):
        p = sc.path_params((0.0, 2.0, 0.0), W, H, spp, 4, 0.0, seed=9)
        sc.drop_lbvh()
        _, brute, st_b = sc.render_path_host(p, prec)
        sc.build_lbvh(huge_radius=50.0)
        assert sc.has_lbvh
        _, bvh, st_v = sc.render_path_host(p, prec)
        if prec == nat.F64:
            assert np.array_equal(brute, bvh)
            assert np.array_equal(st_b[:4], st_v[:4])
        else:
            # FP32: the brute-force loop ranks candidates on the packed form of the discriminant with the 3 low
            # mantissa bits of the key holding the in-group index; the hierarchy ranks on the per-sphere robust form.
            # Near-ties (intersecting spheres seen edge-on) may resolve differently: a handful of pixels.
            differ = (brute != bvh).any(axis=2).mean()
            assert differ < 0.01, differ
            assert np.abs(st_b[:4].astype(np.int64) - st_v[:4].astype(np.int64)).max() <= 0.002 * st_b[0]
        assert st_v[5] < st_b[5] / 10 and st_v[6] > 0          # sphere tests culled, boxes tested
    # Algorithm A through the hierarchy too (signed-distance criterion)
    X, Y = np.linspace(-0.5, 0.5, 96), np.linspace(0.8, 0.2, 64)
    fs.p_id = np.array([int(fs.ids[5])], np.int32); fs.p_pos = fs.centre[5:6].copy(); fs.p_col = fs.colour[5:6].copy()
    fs.p_strength = np.array([3.0]); fs.p_max_angle = np.array([np.pi / 2]); fs.p_func = np.array([0], np.int32)
    sc.update(fs)
    pw = sc.whitted_params((0.0, 2.0, 0.0), X, Y, max_bounces=3)
    _, brute, hit_b, _ = sc.render_whitted_host(pw, nat.F64)
    sc.build_lbvh(huge_radius=50.0)
    _, bvh, hit_v, _ = sc.render_whitted_host(pw, nat.F64)
    assert np.array_equal(hit_b, hit_v) and np.array_equal(brute, bvh)
    sc.close()


# ------------------------------------------------------------------ full-size properties (BASELINE.json sizes)
def test_c3_full_size_properties(nat):
    """Complex scene 1920x1080: tile-sharded render == unsharded, ray statistics in the reference's published band."""
    from ray_tracer_v1_b200 import scenes, flatten_scene
    spec = scenes.build_complex()
    fs = flatten_scene(spec.spheres, background_colour=spec.background)
    sc = nat.DeviceScene(fs)
    W, H, spp = 1920, 1080, 4
    whole = nat.DeviceBuffer((H, W, 4), np.float32)
    st = nat.DeviceBuffer(8, np.uint64)
    sc.render_path(sc.path_params(spec.camera, W, H, spp, 5, 0.9, seed=0), whole, nat.F32, stats=st)
    tiles = nat.DeviceBuffer((H, W, 4), np.float32)
    for r in range(8):
        rows = (H * r // 8, H * (r + 1) // 8)
        sc.render_path(sc.path_params(spec.camera, W, H, spp, 5, 0.9, seed=0, rows=rows), tiles, nat.F32)
    a = whole.download()
    assert np.array_equal(a, tiles.download())
    stats = st.download()
    rays_per_sample = stats[0] / (W * H * spp)
    # traditional_renders/complex_spp_1_230923_stats.txt: 5.79 rays per pixel-sample at depth 5 (enclosed room)
    assert 5.0 < rays_per_sample <= 6.0, rays_per_sample
    assert np.all(a[..., 3] == spp) and a[..., :3].min() >= 0 and a[..., :3].max() <= 255 * spp
    img = nat.DeviceBuffer((H, W, 3), np.float32)
    sc.resolve(whole, W, H, spp, img, nat.F32)
    im = img.download()
    assert np.array_equal(im, np.minimum(1.0, np.floor(a[..., :3].astype(np.float64) / spp) / 255.0).astype(np.float32))
    sc.close()


@pytest.mark.parametrize("precision", ["F32", "F64"])
def test_mean_image_agrees_with_the_reference_rng(nat, precision):
    """The stochastic path against the UNMODIFIED reference drawing from numpy's own MT19937 (a stream the device cannot
    reproduce; golden made by oracle/gen_golden.py path_native): the mean images agree like two seeds of our own
    estimator agree with each other, i.e. within the 1/sqrt(spp) noise, and so do the ray statistics."""
    z, fs = load_golden("path_chandelier_native_rng_40x24")
    W, H, spp = int(z["W"]), int(z["H"]), int(z["spp"])
    ref = z["image"].astype(np.float64) * 255.0
    sc = nat.DeviceScene(fs)
    imgs, rays = [], []
    for seed in (1, 2, 3):
        p = sc.path_params(z["cam"], W, H, spp, int(z["max_bounces"]), float(z["mirror_threshold"]), seed=seed)
        image, _, st = sc.render_path_host(p, getattr(nat, precision))
        imgs.append(image.astype(np.float64) * 255.0)
        rays.append(int(st[0]))
    rm = lambda a, b: float(np.sqrt(np.mean((a - b) ** 2)))      # noqa: E731
    own = np.mean([rm(imgs[0], imgs[1]), rm(imgs[0], imgs[2]), rm(imgs[1], imgs[2])])
    for im in imgs:
        assert rm(ref, im) < 1.25 * own, (rm(ref, im), own)
        assert np.abs((ref - im).mean(axis=(0, 1))).max() < 0.25          # colour levels, averaged over the frame
    assert own < 4.0                                                      # levels at 48 spp
    assert abs(np.mean(rays) - int(z["stats"][0])) < 0.01 * int(z["stats"][0])
    sc.close()


def test_sharded_renderer_async_host_frames(nat):
    """ShardedPathRenderer (one rank): to_host="async" hands back PendingFrames whose copies overlap the next render;
    the frames equal the synchronous ones, through the NCCL-path code and through the fused sinks."""
    import ray_tracer_v1_b200 as pkg
    from ray_tracer_v1_b200 import scenes
    from ray_tracer_v1_b200.distributed import ShardedPathRenderer
    spec = scenes.build_complex()
    fs = pkg.flatten_scene(spec.spheres, background_colour=spec.background)
    r = ShardedPathRenderer()
    r.set_scene(fs)
    W, H, spp = 160, 90, 4
    args = (spec.camera, W, H, spp, spec.max_bounces, spec.mirror_threshold)
    for fn in (r.render, r.render_fused):
        want = [np.array(fn(*args, seed=s, to_host=True)[0]) for s in (1, 2, 3, 4)]
        pend = [fn(*args, seed=s, to_host="async")[0] for s in (1, 2)]          # two frames in flight
        got = [np.array(p.result()) for p in pend]
        pend = [fn(*args, seed=s, to_host="async")[0] for s in (3, 4)]
        got += [np.array(p.result()) for p in pend]
        for a, b in zip(want, got):
            assert np.array_equal(a, b)
        assert not np.array_equal(want[0], want[1])
    r.close()


# ------------------------------------------------------------------ fused multi-GPU sinks (one rank: peers = self)
def test_fused_sinks_equal_accumulate_then_resolve(nat):
    """rt_render_path_sink: the image sink (interleaved stripes rendered by separate launches) and the scatter-add
    sink (two sample ranges added into IPC-exported accumulators, then rt_resolve_clear) both reproduce the frame of
    rt_render_path + rt_resolve bit for bit; epoch flags order a signal before a wait."""
    import torch
    import ray_tracer_v1_b200 as pkg
    from ray_tracer_v1_b200 import scenes
    from ray_tracer_v1_b200.distributed import PeerFabric, row_bands
    spec = scenes.build_complex()
    fs = pkg.flatten_scene(spec.spheres, background_colour=spec.background)
    sc = nat.DeviceScene(fs)
    W, H, spp = 200, 117, 6
    p = sc.path_params(spec.camera, W, H, spp, 5, 0.9, seed=4)
    ref, _, st_ref = sc.render_path_host(p, nat.F32)
    fab = PeerFabric(0)
    img = fab.alloc("image", H * W * 12)[0]
    acc = fab.alloc("accum", H * W * 16)[0]
    fab.alloc("flags", 256)
    stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    out = np.zeros((H, W, 3), np.float32)
    # tiles: three "ranks" render stripes 0,3,6.. / 1,4,7.. / 2,5,8.. into the same image
    for r in range(3):
        sink = nat.PathSink()
        sink.mode, sink.tile_first, sink.tile_step, sink.image = nat.SINK_IMAGE, r, 3, img
        sc.render_path_sink(p, sink, stats=stats)
    nat.check(nat.lib().rt_memcpy_d2h(0, out.ctypes.data, img, out.nbytes, None))
    nat.check(nat.lib().rt_stream_sync(0, None))
    assert np.array_equal(out, ref)
    assert np.array_equal(stats.cpu().numpy()[:5].astype(np.uint64), st_ref[:5])
    # samples: two "ranks" add their sample ranges; owner bands of a world of 2 both live in this process
    bands = row_bands(H, 2)
    for s0, s1 in ((0, 4), (4, 6)):
        q = sc.path_params(spec.camera, W, H, spp, 5, 0.9, seed=4, samples=(s0, s1))
        sink = nat.PathSink()
        sink.mode, sink.world = nat.SINK_SCATTER_ADD, 2
        sink.accum[0] = sink.accum[1] = acc
        sink.band_y[0], sink.band_y[1], sink.band_y[2] = 0, bands[0][1], H
        sc.render_path_sink(q, sink)
    fab.signal("flags", 3, [0], 7)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    fab.wait("flags", 3, 1, 7, flag, timeout_ms=500)
    nat.check(nat.lib().rt_memset_dev(0, img, 0, H * W * 12, None))
    nat.check(nat.lib().rt_resolve_clear(0, acc, W, H, 0, H, spp, img, 1, None))
    nat.check(nat.lib().rt_memcpy_d2h(0, out.ctypes.data, img, out.nbytes, None))
    left = np.ones((H, W, 4), np.float32)
    nat.check(nat.lib().rt_memcpy_d2h(0, left.ctypes.data, acc, left.nbytes, None))
    nat.check(nat.lib().rt_stream_sync(0, None))
    assert np.array_equal(out, ref) and not left.any() and int(flag.item()) == 0
    # a wait nobody signals gives up and says so
    fab.wait("flags", 9, 1, 1, flag, timeout_ms=20)
    assert int(flag.item()) == 1
    fab.close()
    sc.close()


@pytest.mark.parametrize("mode", ["tiles", "samples"])
def test_single_launch_frame_protocol(nat, mode):
    """rt_path_sink.sync: the whole frame protocol (wait for free buffers, render, publish, resolve the own band,
    collect on rank 0) inside ONE launch per rank.  Three "ranks" share this GPU (max_ctas keeps their persistent
    launches co-resident, one stream each, rank 0 launched LAST); four frames through the two buffer parities, the
    "consumed" signal of frame e - 1 published by rank 0's launch of frame e; every frame equals the unsharded one."""
    import torch
    import ray_tracer_v1_b200 as pkg
    from ray_tracer_v1_b200 import scenes
    from ray_tracer_v1_b200.distributed import row_bands, sample_ranges
    spec = scenes.build_complex()
    fs = pkg.flatten_scene(spec.spheres, background_colour=spec.background)
    sc = nat.DeviceScene(fs)
    W, H, spp, world = 384, 216, 7, 3
    dev = torch.device("cuda", 0)
    images = [torch.zeros((H, W, 3), dtype=torch.float32, device=dev) for _ in (0, 1)]
    accum = [[torch.zeros((H, W, 4), dtype=torch.float32, device=dev) for _ in range(world)] for _ in (0, 1)]
    flags = [torch.zeros(nat.FLAG_WORDS, dtype=torch.int32, device=dev) for _ in range(world)]
    timed_out = torch.zeros(1, dtype=torch.int32, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    bands = row_bands(H, world)
    torch.cuda.synchronize()
    for e in range(1, 5):
        buf = e & 1
        ref, _, _ = sc.render_path_host(sc.path_params(spec.camera, W, H, spp, 5, 0.9, seed=100 + e), nat.F32)
        for rank in reversed(range(world)):
            p = sc.path_params(spec.camera, W, H, spp, 5, 0.9, seed=100 + e)
            sink = nat.PathSink()
            sink.world, sink.rank, sink.sync, sink.epoch, sink.spp_total = world, rank, 1, e, spp
            sink.go_epoch = e - 1 if rank == 0 else 0
            sink.image = images[buf].data_ptr()
            sink.timed_out, sink.timeout_ms, sink.max_ctas = timed_out.data_ptr(), 4000, 148
            for k in range(world):
                sink.flags[k] = flags[k].data_ptr()
            if mode == "tiles":
                sink.mode, sink.tile_first, sink.tile_step = nat.SINK_IMAGE, rank, world
            else:
                sink.mode = nat.SINK_SCATTER_ADD
                p.s0, p.s1 = sample_ranges(spp, world)[rank]
                for k in range(world):
                    sink.accum[k] = accum[buf][k].data_ptr()
                    sink.band_y[k] = bands[k][0]
                sink.band_y[world] = H
            sc.render_path_sink(p, sink, stream=streams[rank].cuda_stream)
        # rank 0's launch ends when the frame is complete: its stream alone orders the read-back
        streams[0].synchronize()
        assert int(timed_out.item()) == 0, "a flag wait timed out: the launches did not run concurrently"
        assert np.array_equal(images[buf].cpu().numpy(), ref), f"frame {e}"
        torch.cuda.synchronize()
        if mode == "samples":
            assert not any(bool(a.any()) for a in accum[buf]), "bands are cleared for the frame after next"
    assert [int(f[32]) for f in flags] == [3] * world          # go: frames <= 3 consumed, published to every rank
    sc.close()


def test_frame_buffer_is_not_reused_before_its_consumer_ran(nat):
    """ADVICE r1 (distributed.py): in tile mode a peer could start frame e + 2 -- which stores into the image buffer of
    frame e -- before rank 0's consumer of frame e had run.  The "consumed" signal is now published by rank 0's NEXT
    launch on its stream, i.e. after whatever the caller queued on the image.  Here rank 0's stream sleeps 40 ms and only
    then copies frame e out, while the peers' launches of frames e + 1 and e + 2 are already queued on their streams: the
    copy must still be frame e."""
    import torch
    import ray_tracer_v1_b200 as pkg
    from ray_tracer_v1_b200 import scenes
    spec = scenes.build_complex()
    fs = pkg.flatten_scene(spec.spheres, background_colour=spec.background)
    sc = nat.DeviceScene(fs)
    W, H, spp, world = 256, 144, 4, 3
    dev = torch.device("cuda", 0)
    images = [torch.zeros((H, W, 3), dtype=torch.float32, device=dev) for _ in (0, 1)]
    flags = [torch.zeros(nat.FLAG_WORDS, dtype=torch.int32, device=dev) for _ in range(world)]
    timed_out = torch.zeros(1, dtype=torch.int32, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    refs = {e: sc.render_path_host(sc.path_params(spec.camera, W, H, spp, 5, 0.9, seed=300 + e), nat.F32)[0] for e in range(1, 5)}
    snaps = {}
    torch.cuda.synchronize()

    def launch(rank, e):
        sink = nat.PathSink()
        sink.mode, sink.tile_first, sink.tile_step = nat.SINK_IMAGE, rank, world
        sink.world, sink.rank, sink.sync, sink.epoch = world, rank, 1, e
        sink.go_epoch = e - 1 if rank == 0 else 0
        sink.image = images[e & 1].data_ptr()
        sink.timed_out, sink.timeout_ms, sink.max_ctas = timed_out.data_ptr(), 8000, 148
        for k in range(world):
            sink.flags[k] = flags[k].data_ptr()
        sc.render_path_sink(sc.path_params(spec.camera, W, H, spp, 5, 0.9, seed=300 + e), sink, stream=streams[rank].cuda_stream)

    # Issue order matters on ONE device: a launch that spins on a flag must have been issued AFTER the launch that will
    # set it (streams of one process share hardware queues, a later-issued kernel can be held behind a spinning one).
    # Across GPUs -- the real deployment -- every rank has its own queues and this constraint does not exist.
    cycles = int(0.04 * 1.9e9)
    for rank in (1, 2):
        launch(rank, 1); launch(rank, 2)
    launch(0, 1)
    for e in range(1, 5):
        with torch.cuda.stream(streams[0]):  # rank 0: a slow consumer of frame e ...
            torch.cuda._sleep(cycles)
            snaps[e] = images[e & 1].clone()
        if e + 1 <= 4:
            launch(0, e + 1)                 # ... then its next frame, whose prologue publishes "frame e consumed"
        if e + 2 <= 4:
            for rank in (1, 2):
                launch(rank, e + 2)          # the peers' frame e + 2 (same buffer as e) is queued while the consumer sleeps
    torch.cuda.synchronize()
    assert int(timed_out.item()) == 0
    for e in range(1, 5):
        assert np.array_equal(snaps[e].cpu().numpy(), refs[e]), f"frame {e} was overwritten before its consumer ran"
    sc.close()


def test_sample_split_gives_the_same_frame(nat):
    """ksplit: 1..32 lanes sharing a pixel's samples (butterfly-summed) render the frame of one thread per pixel, for
    ragged sizes, row bands, sample ranges that do not divide by k, both schedules and the image sink."""
    import torch
    import ray_tracer_v1_b200 as pkg
    from ray_tracer_v1_b200 import scenes
    spec = scenes.build_chandelier()
    fs = pkg.flatten_scene(spec.spheres, background_colour=spec.background)
    sc = nat.DeviceScene(fs)
    W, H, spp = 101, 45, 11
    base = sc.path_params(spec.camera, W, H, spp, 6, 0.0, seed=2, ksplit=0)
    _, ref, st_ref = sc.render_path_host(base, nat.F32)
    for k in (-1, 2, 4, 8, 16, 32):
        for schedule in (0, 1):
            p = sc.path_params(spec.camera, W, H, spp, 6, 0.0, seed=2, ksplit=k, schedule=schedule)
            _, out, st = sc.render_path_host(p, nat.F32)
            assert np.array_equal(out, ref), (k, schedule)
            assert np.array_equal(st[:5], st_ref[:5]), (k, schedule)      # slot 5 (executed sphere tests) depends on the tiling
            assert st[5] <= st[4] * len(fs.ids)
    # row band + sample range, accumulated in two launches with different k
    acc = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    for (s0, s1), k in (((0, 7), 4), ((7, 11), 8)):
        p = sc.path_params(spec.camera, W, H, spp, 6, 0.0, seed=2, rows=(9, 30), samples=(s0, s1), accumulate=s0 > 0, ksplit=k)
        sc.render_path(p, acc, nat.F32)
    got = acc.cpu().numpy()
    assert np.array_equal(got[9:30], ref[9:30]) and not got[:9].any() and not got[30:].any()
    # image sink, interleaved stripes, k = 8: CTA rows of 4 inside 8-row stripes
    img = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    for r in range(2):
        sink = nat.PathSink()
        sink.mode, sink.tile_first, sink.tile_step, sink.image = nat.SINK_IMAGE, r, 2, img.data_ptr()
        sc.render_path_sink(sc.path_params(spec.camera, W, H, spp, 6, 0.0, seed=2, ksplit=8), sink)
    want = np.minimum(1.0, np.floor(ref[..., :3].astype(np.float64) / spp) / 255.0).astype(np.float32)
    assert np.array_equal(img.cpu().numpy(), want)
    # automatic split: the last owned stripes of a launch are traced by FINER work units (more lanes per pixel) than the
    # first ones -- two tile grids in one launch; three ranks' interleaved stripes, ragged last stripe
    img.zero_()
    for r in range(3):
        sink = nat.PathSink()
        sink.mode, sink.tile_first, sink.tile_step, sink.image = nat.SINK_IMAGE, r, 3, img.data_ptr()
        sc.render_path_sink(sc.path_params(spec.camera, W, H, spp, 6, 0.0, seed=2), sink)
    assert np.array_equal(img.cpu().numpy(), want)
    # 2-D interleave (rt_path_sink::col_split): rank r renders column segment (r + s) mod 4 of every stripe s
    W2, H2 = 256, 45
    _, ref2, _ = sc.render_path_host(sc.path_params(spec.camera, W2, H2, spp, 6, 0.0, seed=2, ksplit=0), nat.F32)
    img2 = torch.zeros((H2, W2, 3), dtype=torch.float32, device="cuda")
    for k in (-1, 8):
        img2.zero_()
        for r in range(4):
            sink = nat.PathSink()
            sink.mode, sink.tile_first, sink.tile_step, sink.image, sink.col_split = nat.SINK_IMAGE, r, 4, img2.data_ptr(), 1
            sc.render_path_sink(sc.path_params(spec.camera, W2, H2, spp, 6, 0.0, seed=2, ksplit=k), sink)
            if r == 0:                    # one rank alone: exactly its diagonal of segments is written
                got = img2.cpu().numpy().any(axis=2)
                for srow in range(0, H2, 8):
                    seg = (srow // 8) % 4
                    assert got[srow:srow + 8, seg * 64:(seg + 1) * 64].any() and not np.delete(got[srow:srow + 8], np.s_[seg * 64:(seg + 1) * 64], axis=1).any()
        want2 = np.minimum(1.0, np.floor(ref2[..., :3].astype(np.float64) / spp) / 255.0).astype(np.float32)
        assert np.array_equal(img2.cpu().numpy(), want2), k
    sink = nat.PathSink()
    sink.mode, sink.tile_first, sink.tile_step, sink.image, sink.col_split = nat.SINK_IMAGE, 0, 3, img2.data_ptr(), 1
    with pytest.raises(Exception):        # 256 is not a multiple of 3
        sc.render_path_sink(sc.path_params(spec.camera, W2, H2, spp, 6, 0.0, seed=2), sink)
    # a segment too narrow for the work units of the launch (256 / 16 = 16 pixels, units 32 wide at k = 1): whole stripes
    img2.zero_()
    for r in range(16):
        sink = nat.PathSink()
        sink.mode, sink.tile_first, sink.tile_step, sink.image, sink.col_split = nat.SINK_IMAGE, r, 16, img2.data_ptr(), 1
        sc.render_path_sink(sc.path_params(spec.camera, W2, H2, spp, 6, 0.0, seed=2, ksplit=1), sink)
    assert np.array_equal(img2.cpu().numpy(), want2)
    big = sc.path_params(spec.camera, 640, 360, 16, 6, 0.0, seed=4)               # coarse k = 2, fine k = 8 over the last stripes
    _, auto, st_auto = sc.render_path_host(big, nat.F32)
    _, flat, st_flat = sc.render_path_host(sc.path_params(spec.camera, 640, 360, 16, 6, 0.0, seed=4, ksplit=2), nat.F32)
    assert np.array_equal(auto, flat) and np.array_equal(st_auto[:5], st_flat[:5])
    sc.close()


# ------------------------------------------------------------------ "Algorithm C": FB/output6.py traditional mode
@pytest.mark.parametrize("name", ["simple_balls_a_64x48", "simple_balls_b_40x30"])
def test_output6_frames_match_reference(nat, orc, name):
    z, fs = load_golden(name)
    W, H, seed, depth = int(z["W"]), int(z["H"]), int(z["seed"]), int(z["max_bounces"])
    sc = nat.DeviceScene(fs)
    p = sc.simple_params(W, H, max_bounces=depth, seed=seed)
    img64, rgb64, st64 = sc.render_simple_host(p, nat.F64)
    assert np.array_equal(rgb64[..., :3], z["rgb"].astype(np.int32))          # the reference's own render, exactly
    assert np.array_equal(img64, z["image"])
    assert [int(st64[0]), int(st64[1])] == list(z["stats"])
    img32, rgb32, st32 = sc.render_simple_host(p, nat.F32)
    d = np.abs(rgb32[..., :3].astype(np.int64) - z["rgb"].astype(np.int64)).max(axis=2)
    gate(f"{name} (output6) FP32 pixels beyond one level", (d > 1).mean(), 6e-3)          # same Philox stream: a few silhouette / int() flips
    assert abs(int(st32[0]) - int(st64[0])) <= 0.01 * int(st64[0])
    # a larger frame against the oracle (FP64 exact), and the explicit-ray entry against the frame
    q = sc.simple_params(200, 150, max_bounces=depth, seed=seed + 1)
    _, big64, stb = sc.render_simple_host(q, nat.F64)
    ref, sto = orc.render_simple(fs, 200, 150, seed=seed + 1, max_bounces=depth)
    assert np.array_equal(big64[..., :3], ref.astype(np.int32)) and int(stb[0]) == sto["total_rays"] and int(stb[1]) == sto["sun_hits"]
    sc.close()


def test_output6_lighting_helper_matches_reference(nat, rt, orc):
    """calculate_lighting_exact_original on its own (rt_simple_params.lighting_only): the reference's own (intersection
    -> Colour) pairs exactly in FP64, within one level in FP32 but for a few shadow / int() flips; then the class
    method."""
    z, fs = load_golden("simple_lighting_balls")
    sc = nat.DeviceScene(fs)
    p = sc.simple_params(1, 1)
    _, rgb64, st = sc.render_simple_host(p, nat.F64, hits=z["hits"])
    assert np.array_equal(rgb64[0, :, :3], z["rgb"].astype(np.int32)) and int(st[1]) == 0
    _, rgb32, _ = sc.render_simple_host(p, nat.F32, hits=z["hits"])
    d = np.abs(rgb32[0, :, :3].astype(np.int64) - z["rgb"].astype(np.int64)).max(axis=1)
    gate("output6 lighting helper FP32 rows beyond one level", (d > 1).mean(), 6e-3)
    sc.close()
    r = rt.SimplifiedFBRenderer(precision="f64", seed=5)
    sph = r.scene[2]
    n = rt.Vector(0.0, 0.6, 0.8)
    pt = sph.centre.addVector(n.scaleByLength(sph.radius))
    it = SimpleNamespace(object=sph, point=pt, normal=n)
    c = r.calculate_lighting_exact_original(it)
    want, _ = orc.simple_lighting(rt.flatten_scene(r.scene), [[pt.x, pt.y, pt.z, n.x, n.y, n.z, 2.0]])
    assert (c.r, c.g, c.b) == tuple(int(v) for v in want[0])
    sun = next(s for s in r.scene if s.id == 7)
    c = r.calculate_lighting_exact_original(SimpleNamespace(object=sun, point=sun.centre, normal=n))
    assert (c.r, c.g, c.b) == (255, 255, 204) and r.stats["sun_hits"] == 1


def test_output6_dropin_class(rt, orc):
    r = rt.SimplifiedFBRenderer(precision="f64", seed=5)
    image, path = r.render_original_style(96, 72, output_path="")
    fs = rt.flatten_scene(r.scene)
    ref, st = orc.render_simple(fs, 96, 72, seed=5, max_bounces=5)
    assert np.array_equal(image, np.minimum(1.0, ref / 255.0).astype(np.float32))
    assert r.stats["total_rays"] == st["total_rays"] and r.stats["sun_hits"] == st["sun_hits"]
    # scalar entry: a ray at the big blue sphere and one that bounces off the mirror towards the sun's side
    r.max_bounces = 8
    c = r.trace_ray_simple(rt.Ray(rt.Vector(0, 0, 1), rt.Vector(0.07, -0.07, -1)))
    want, _ = orc.render_simple(fs, 1, 1, seed=5, max_bounces=8, rays=np.array([[0, 0, 1, 0.07, -0.07, -1.0]]))
    assert (c.r, c.g, c.b) == tuple(int(v) for v in want[0, 0])
    r.fb_usage_prob = 0.5
    with pytest.raises(NotImplementedError):
        r.render_original_style(8, 8, output_path="")


# ------------------------------------------------------------------ edge cases
def test_edge_cases_match_the_oracle(nat, orc):
    """Empty scene, a scene without lights, one sphere, frames that are not tile multiples, depth 0 / 1, a sample
    range longer than the integer accumulators allow, and the error codes of bad arguments."""
    import ctypes as C
    import ray_tracer_v1_b200 as pkg
    from ray_tracer_v1_b200 import scenes
    V, Col, M, S = pkg.Vector, pkg.Colour, pkg.Material, pkg.Sphere
    cam = (0.0, 0.0, 5.0)

    def both(fs, W, H, spp, depth, thr=0.0, seed=1):
        sc = nat.DeviceScene(fs)
        p = sc.path_params(cam, W, H, spp, depth, thr, seed=seed)
        ref, st = orc.render_path(fs, cam, W, H, spp, depth, thr, seed=seed)
        for prec in (nat.F64, nat.F32):
            img, sums, stats = sc.render_path_host(p, prec)
            if prec == nat.F64:
                assert np.array_equal(sums[..., :3], ref), "FP64 path differs"
                assert int(stats[0]) == st["total_rays"] and int(stats[1]) == st["total_intersections"]
            else:
                assert (np.abs(sums[..., :3] - ref).max(axis=2) > spp).mean() < 0.03
            assert np.array_equal(sums[..., 3], np.full((H, W), spp))
        sc.close()
        return ref

    empty = pkg.flatten_scene([], background_colour=Col(2, 2, 5))
    r = both(empty, 33, 9, 3, 4)
    assert np.array_equal(r, np.broadcast_to(np.array([2.0, 2.0, 5.0]) * 3, r.shape))       # every ray misses
    matte = M(reflective=0, transparent=0, emitive=0, refractive_index=1)
    lamp = M(reflective=0, transparent=0, emitive=1, refractive_index=1)
    no_lights = pkg.flatten_scene([S(V(0, 0, 0), 1.0, matte, Col(200, 100, 50), id=1),
                                   S(V(0, -101, 0), 100.0, matte, Col(90, 90, 90), id=2)], background_colour=Col(2, 2, 5))
    both(no_lights, 37, 21, 2, 3)
    one = pkg.flatten_scene([S(V(0, 0, 0), 1.5, lamp, Col(255, 240, 200), id=9)], background_colour=Col(2, 2, 5))
    r = both(one, 31, 17, 1, 2)
    assert r.max() == 255 and r.min() == 2
    spec = scenes.build_chandelier()
    fs = pkg.flatten_scene(spec.spheres, background_colour=spec.background)
    both(fs, 45, 13, 2, 0)          # max_bounces 0: every call returns at the depth check
    both(fs, 45, 13, 2, 1)
    # 9 spheres: the padded selection loop reads a second, mostly empty group
    nine = pkg.flatten_scene(spec.spheres[:9], background_colour=spec.background)
    both(nine, 40, 24, 2, 4)
    # argument errors come back as status codes with a message, not as crashes
    sc = nat.DeviceScene(fs)
    p = sc.path_params(cam, 16, 16, 1, 40, 0.0)             # depth above RT_PATH_MAX_DEPTH
    with pytest.raises(nat.NativeLibraryError, match="max_bounces"):
        sc.render_path_host(p, nat.F32)
    p = sc.path_params(cam, 16, 16, 1, 4, 0.0, rows=(8, 40))
    with pytest.raises(nat.NativeLibraryError, match="row band"):
        sc.render_path_host(p, nat.F32)
    p = sc.path_params(cam, 16, 16, 1, 4, 0.0)
    with pytest.raises(nat.NativeLibraryError):
        sc.render_path_host(p, 7)                            # unknown precision
    sink = nat.PathSink()
    sink.mode = nat.SINK_IMAGE                               # no image pointer
    with pytest.raises(nat.NativeLibraryError, match="image sink"):
        sc.render_path_sink(p, sink)
    assert nat.lib().rt_render_path(None, nat.F32, C.byref(p), None, None, None) != 0
    sc.close()


# ------------------------------------------------------------------ wavefront Algorithm B with a policy (f-4)
def test_wavefront_fb_renderer(rt, nat, orc):
    """rt_wf_*: (1) without a policy the wavefront equals rt_render_path bit for bit (FP64 and FP32: same device
    functions, same Philox streams), also when the frame is cut into sample chunks; (2) with the stand-in policy the
    FP64 build reproduces the reference's own WorkingFBRenderer.render (golden) and the oracle; (3) the drop-in class."""
    import torch
    from test_oracle_golden import fb_test_policy
    from ray_tracer_v1_b200 import renderers
    z, fs = load_golden("path_fb_complex_40x24")
    W, H, spp, depth, thr = int(z["W"]), int(z["H"]), int(z["spp"]), int(z["max_bounces"]), float(z["mirror_threshold"])
    seed, prob = int(z["seed"]), float(z["fb_usage_prob"])
    sc = nat.DeviceScene(fs)
    p = sc.path_params(z["cam"], W, H, spp, depth, thr, seed=seed)
    for prec, name in ((nat.F64, "f64"), (nat.F32, "f32")):
        _, ref, st = sc.render_path_host(p, prec)
        for max_paths in (1 << 22, W * H):                     # one chunk / one sample per chunk
            _, sums, out = renderers.render_path_wavefront(fs, z["cam"], W, H, spp, depth, thr, seed=seed, precision=name,
                                                            max_paths=max_paths, scene=sc)
            assert np.array_equal(sums, ref), (name, max_paths)
            assert [out["total_rays"], out["total_intersections"], out["light_hits"], out["small_light_hits"]] == [int(v) for v in st[:4]]
            assert out["fb_used"] == 0

    def policy(obs):                                            # torch twin of fb_test_policy: same float32 bits
        a0 = (obs[:, 6] * 0.5 + obs[:, 7] * 0.25 - 0.125).clamp(-1, 1)
        a1 = (obs[:, 8] * 0.5 + obs[:, 3] * 0.25 + obs[:, 16] * 0.5).clamp(-1, 1)
        return torch.stack([a0, a1], dim=1)

    _, sums64, out = renderers.render_path_wavefront(fs, z["cam"], W, H, spp, depth, thr, policy, prob, seed=seed,
                                                     precision="f64", scene=sc)
    assert np.array_equal(sums64[..., :3], z["sums"])          # the reference's own FB render
    assert [out[k] for k in ("total_rays", "total_intersections", "light_hits", "small_light_hits", "fb_used")] == list(z["stats"])
    _, sums32, out32 = renderers.render_path_wavefront(fs, z["cam"], W, H, spp, depth, thr, policy, prob, seed=seed,
                                                       precision="f32", scene=sc)
    assert (np.abs(sums32[..., :3] - z["sums"]).max(axis=2) > spp).mean() < 0.03
    assert abs(out32["fb_used"] - out["fb_used"]) < 0.01 * out["fb_used"]
    # a bigger frame against the oracle, the policy through the reference's per-observation protocol
    ref, sto = orc.render_path_fb(fs, z["cam"], 64, 36, 2, depth, thr, fb_test_policy, 1.0, seed + 1)
    agent = type("Agent", (), {"choose_direction": staticmethod(fb_test_policy)})()
    _, s2, o2 = renderers.render_path_wavefront(fs, z["cam"], 64, 36, 2, depth, thr, renderers._batched_policy(agent), 1.0,
                                                seed=seed + 1, precision="f64", scene=sc)
    assert np.array_equal(s2[..., :3], ref) and o2["fb_used"] == sto["fb_used"] and o2["total_rays"] == sto["total_rays"]
    sc.close()
    # drop-in class
    from ray_tracer_v1_b200 import scenes
    spec = scenes.build_complex()
    r = rt.WorkingFBRenderer(camera_position=rt.Vector(*spec.camera), precision="f64", seed=seed)
    r.scene = spec.spheres
    r.light_sources = [s for s in spec.spheres if s.material.emitive]
    r.small_lights = [s for s in r.light_sources if s.radius < 0.5]
    r.fb_agent = type("Agent", (), {"choose_directions": staticmethod(policy)})()
    r.fb_loaded, r.fb_usage_prob = True, prob
    img = r.render(W, H, spp, depth)
    assert np.array_equal(img, z["image"]) and r.stats["fb_used"] == int(z["stats"][4]) and r.stats["fb_success"] == r.stats["fb_used"]
    # without an agent (the reference without a checkpoint) the class renders the traditional frame through the fused
    # path kernel: same image and counters as ComplexTraditionalRenderer with the same seed, in both precisions
    for name in ("f64", "f32"):
        plain = rt.WorkingFBRenderer(camera_position=rt.Vector(*spec.camera), precision=name, seed=seed)
        trad = rt.ComplexTraditionalRenderer(precision=name, seed=seed)
        for q in (plain, trad):
            q.scene, q.light_sources, q.small_lights = spec.spheres, r.light_sources, r.small_lights
        trad.camera_position = plain.camera_position
        a, b = plain.render(W, H, spp, depth), trad.render(W, H, spp, depth)
        assert np.array_equal(a, b), name
        assert all(plain.stats[k] == trad.stats[k] for k in ("total_rays", "total_intersections", "light_hits", "small_light_hits"))
        assert plain.stats["fb_used"] == 0 and plain.stats["total_rays"] > 0
        # an agent with fb_usage_prob = 0 asks nothing of it either
        plain.fb_agent, plain.fb_loaded, plain.fb_usage_prob = r.fb_agent, True, 0.0
        assert np.array_equal(plain.render(W, H, spp, depth), b), name


# ------------------------------------------------------------------ full-size properties of C2 and C4
def test_c2_full_size_properties(nat):
    """Marbles scene 1280x720, 16 spp, depth 4 (BASELINE config 2): row bands x sample ranges accumulate to the whole
    frame bit for bit, the deterministic centre-of-pixel frame is reproducible, and jitter only moves edge pixels."""
    from ray_tracer_v1_b200 import scenes, flatten_scene
    spec = scenes.build_marbles4()
    fs = flatten_scene(spec.spheres, spec.global_lights, spec.point_lights, spec.background)
    sc = nat.DeviceScene(fs)
    W, H, spp = 1280, 720, 16
    k = 640 * spec.ray_step
    X, Y = np.linspace(-k * 16 / 9, k * 16 / 9, W), np.linspace(k, -k, H)
    miss = [spec.miss.r, spec.miss.g, spec.miss.b]
    whole = nat.DeviceBuffer((H, W, 4), np.float32)
    st = nat.DeviceBuffer(8, np.uint64)
    sc.render_whitted(sc.whitted_params(spec.camera, X, Y, spp=spp, max_bounces=4, miss=miss, seed=3), whole, nat.F32, stats=st)
    parts = nat.DeviceBuffer((H, W, 4), np.float32)
    for rows in ((0, 250), (250, 720)):
        for i, smp in enumerate(((0, 5), (5, 16))):
            sc.render_whitted(sc.whitted_params(spec.camera, X, Y, spp=spp, max_bounces=4, miss=miss, seed=3, rows=rows,
                                                samples=smp, accumulate=i > 0), parts, nat.F32)
    a = whole.download()
    assert np.array_equal(a, parts.download())
    assert np.all(a[..., 3] == spp) and int(st.download()[0]) == W * H * spp
    one = nat.DeviceBuffer((H, W, 4), np.float32)
    sc.render_whitted(sc.whitted_params(spec.camera, X, Y, spp=1, max_bounces=4, miss=miss), one, nat.F32)
    d = np.abs(a[..., :3] / spp - one.download()[..., :3]).max(axis=2)
    assert (d > 2).mean() < 0.02              # supersampling changes silhouettes and shadow edges only
    sc.close()


def test_c4_full_size_properties(nat):
    """Chandelier 1920x1080 (BASELINE config 4): sample ranges summed == the whole frame, LBVH == brute force up to
    near-tie pixels, rays per sample in the band the reference measures (8.14 at depth 8)."""
    from ray_tracer_v1_b200 import scenes, flatten_scene
    spec = scenes.build_chandelier()
    fs = flatten_scene(spec.spheres, background_colour=spec.background)
    sc = nat.DeviceScene(fs)
    W, H, spp = 1920, 1080, 8
    whole = nat.DeviceBuffer((H, W, 4), np.float32)
    st = nat.DeviceBuffer(8, np.uint64)
    sc.render_path(sc.path_params(spec.camera, W, H, spp, 8, 0.0, seed=1), whole, nat.F32, stats=st)
    parts = nat.DeviceBuffer((H, W, 4), np.float32)
    for i, smp in enumerate(((0, 3), (3, 4), (4, 8))):
        sc.render_path(sc.path_params(spec.camera, W, H, spp, 8, 0.0, seed=1, samples=smp, accumulate=i > 0), parts, nat.F32)
    a = whole.download()
    assert np.array_equal(a, parts.download())
    rays_per_sample = st.download()[0] / (W * H * spp)
    assert 7.5 < rays_per_sample < 8.6, rays_per_sample
    sc.build_lbvh(50.0)
    bvh = nat.DeviceBuffer((H, W, 4), np.float32)
    sc.render_path(sc.path_params(spec.camera, W, H, spp, 8, 0.0, seed=1), bvh, nat.F32)
    assert (bvh.download() != a).any(axis=2).mean() < 0.01
    sc.close()


# ------------------------------------------------------------------ determinism under concurrency (race hunting)
def test_concurrent_launches_are_deterministic(nat):
    """compute-sanitizer's racecheck is not available on the pool these kernels run on, so the shared device state is
    stressed directly: 24 frames in flight on three streams of ONE scene handle (self-re-arming work counters, rotating
    counter slots), each into its own buffer, must all equal the frame rendered alone; twelve LBVH builds (atomic
    arrival counters + fences in the refit) must all render that frame too."""
    import torch
    import ray_tracer_v1_b200 as pkg
    from ray_tracer_v1_b200 import scenes
    spec = scenes.build_complex()
    fs = pkg.flatten_scene(spec.spheres, background_colour=spec.background)
    sc = nat.DeviceScene(fs)
    W, H, spp = 640, 360, 8
    p = sc.path_params(spec.camera, W, H, spp, 5, 0.9, seed=12)
    ref = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    sc.render_path(p, ref, nat.F32)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    bufs = [torch.zeros((H, W, 4), dtype=torch.float32, device="cuda") for _ in range(24)]
    torch.cuda.synchronize()
    for i, b in enumerate(bufs):
        sc.render_path(p, b, nat.F32, stream=streams[i % 3].cuda_stream)
    torch.cuda.synchronize()
    assert all(torch.equal(b, ref) for b in bufs)
    big = scenes.build_many_spheres_flat(4000, seed=3)
    sb = nat.DeviceScene(big)
    q = sb.path_params((0.0, 2.0, 0.0), 320, 180, 2, 8, 0.0, seed=1)
    first = None
    for _ in range(12):
        sb.build_lbvh(50.0)
        out = torch.zeros((180, 320, 4), dtype=torch.float32, device="cuda")
        sb.render_path(q, out, nat.F32)
        torch.cuda.synchronize()
        first = out if first is None else first
        assert torch.equal(out, first)
    sb.close(); sc.close()


def test_two_pass_whitted_frames_equal_single_pass(nat):
    """Algorithm-A frames of >= 4 samples take the two-pass schedule (pass 1 fills the sky tiles and lists the others,
    pass 2 spreads (tile, sample) units over the device and adds them with FP32 reductions).  With integer colours the
    sums are exact, so the frame equals the single-pass frame (rows x sample ranges accumulated: no list, every pixel
    traced by its own thread) bit for bit -- also with frames in flight on three streams at once, in accumulate mode,
    and with a non-integer miss colour (which must fall back to the single pass)."""
    import torch
    from ray_tracer_v1_b200 import scenes, flatten_scene
    for spec, W, H in ((scenes.build_planets2(), 640, 360), (scenes.build_marbles4(), 333, 187)):
        fs = flatten_scene(spec.spheres, spec.global_lights, spec.point_lights, spec.background)
        sc = nat.DeviceScene(fs)
        k = 640 * spec.ray_step
        X, Y = np.linspace(-k * 16 / 9, k * 16 / 9, W), np.linspace(k, -k, H)
        miss = [spec.miss.r, spec.miss.g, spec.miss.b]
        spp = 12
        # single pass: sample ranges of fewer than 4 samples never split
        single = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        for i, s0 in enumerate(range(0, spp, 3)):
            sc.render_whitted(sc.whitted_params(spec.camera, X, Y, spp=spp, max_bounces=4, miss=miss, seed=8, samples=(s0, s0 + 3),
                                                accumulate=i > 0), single, nat.F32)
        st = torch.zeros(8, dtype=torch.int64, device="cuda")
        st1 = torch.zeros(8, dtype=torch.int64, device="cuda")
        sc.render_whitted(sc.whitted_params(spec.camera, X, Y, spp=spp, max_bounces=4, miss=miss, seed=8, samples=(0, 3)), torch.zeros_like(single), nat.F32, stats=st1)
        p = sc.whitted_params(spec.camera, X, Y, spp=spp, max_bounces=4, miss=miss, seed=8)
        two = torch.zeros_like(single)
        sc.render_whitted(p, two, nat.F32, stats=st)
        torch.cuda.synchronize()
        assert torch.equal(two, single)
        assert int(st[0]) == W * H * spp and int(st1[0]) == W * H * 3
        # frames in flight on three streams share nothing but the scene
        streams = [torch.cuda.Stream() for _ in range(3)]
        bufs = [torch.zeros_like(single) for _ in range(9)]
        torch.cuda.synchronize()
        for i, b in enumerate(bufs):
            sc.render_whitted(p, b, nat.F32, stream=streams[i % 3].cuda_stream)
        torch.cuda.synchronize()
        assert all(torch.equal(b, single) for b in bufs)
        # accumulate mode: two halves of the samples, both through the two-pass schedule
        acc = torch.zeros_like(single)
        sc.render_whitted(sc.whitted_params(spec.camera, X, Y, spp=spp, max_bounces=4, miss=miss, seed=8, samples=(0, 6)), acc, nat.F32)
        sc.render_whitted(sc.whitted_params(spec.camera, X, Y, spp=spp, max_bounces=4, miss=miss, seed=8, samples=(6, 12), accumulate=True), acc, nat.F32)
        torch.cuda.synchronize()
        assert torch.equal(acc, single)
        # non-integer miss colour: single pass, same per-pixel addition order as the sample-range reference
        m2 = [miss[0] + 0.25, miss[1], miss[2]]
        a = torch.zeros_like(single); b = torch.zeros_like(single)
        sc.render_whitted(sc.whitted_params(spec.camera, X, Y, spp=spp, max_bounces=4, miss=m2, seed=8), a, nat.F32)
        for i, s0 in enumerate(range(0, spp, 3)):
            sc.render_whitted(sc.whitted_params(spec.camera, X, Y, spp=spp, max_bounces=4, miss=m2, seed=8, samples=(s0, s0 + 3),
                                                accumulate=i > 0), b, nat.F32)
        torch.cuda.synchronize()
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-3)
        sc.close()
