"""Pins the CPU oracle (oracle/rt_oracle.c) to the reference.

The golden vectors under tests/golden/ were produced by oracle/gen_golden.py
from the UNMODIFIED reference Python modules (and its one valid stored notebook
probe).  The oracle is IEEE-double and follows the reference's operation order,
so agreement is expected to ~1 ulp; the asserted tolerance is 1e-9 relative
(north_star's FP64 parity bound), and exact for integer-valued colours.
"""
import numpy as np
import pytest

from conftest import load_golden

RTOL = 1e-9


def test_philox_known_answers(orc):
    # Random123 kat_vectors, philox4x32-10
    assert [hex(x) for x in orc.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in orc.philox([0xffffffff] * 4, [0xffffffff] * 2)] == \
        ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    u = orc.rng_pair(7, 123, 4, 3)
    assert 0 <= u[0] < 1 and 0 <= u[1] < 1 and np.float32(u[0]) == u[0]


def test_notebook_cell7_kat(orc):
    """RL/Marbles 1.ipynb cell 7, values as STORED in the notebook output."""
    hit, t, p, n = orc.sphere_discriminant([0.1, 0, 5], [0, 0, -1], [0, 0, 0], 1.0)
    assert hit
    assert tuple(p) == (0.1, 0.0, 0.9949874371066194)
    r = orc.refract([0, 0, -1], n, 1, 1.5)
    assert tuple(r) == (-0.033445034506863806, 0.0, -0.9994405583459353)
    z, _ = load_golden("kat")
    assert np.array_equal(z["notebook7"], np.array([*p, *r]))


def test_sphere_discriminant_kat(orc):
    z, _ = load_golden("kat")
    d = z["disc"]
    assert d[:, 11].sum() > 100 and (d[:, 11] == 0).sum() > 50
    neg = 0
    for row in d:
        hit, t, p, n = orc.sphere_discriminant(row[0:3], row[3:6], row[6:9], row[9], int(row[10]))
        assert hit == bool(row[11])
        if hit:
            np.testing.assert_allclose([t, *p, *n], row[12:19], rtol=RTOL, atol=1e-12)
            neg += t < 0
    assert neg > 5      # origin-inside-sphere cases return a NEGATIVE distance (ray.py:93-96)


def test_reflect_refract_kat(orc):
    z, _ = load_golden("kat")
    for row in z["reflect"]:
        np.testing.assert_allclose(orc.reflect(row[0:3], row[3:6]), row[6:9], rtol=RTOL, atol=1e-15)
    tir = 0
    for row in z["refract"]:
        out = orc.refract(row[0:3], row[3:6], row[6], row[7])
        if row[8] == 0:
            assert out is False
            tir += 1
        else:
            np.testing.assert_allclose(out, row[9:12], rtol=RTOL, atol=1e-15)
    assert tir > 10


WHITTED = ["whitted_c1_balls_320x240", "whitted_balls_true_original_121", "whitted_marbles4_d4_121",
           "whitted_marbles4_d8_121", "whitted_planets2_d4_121", "whitted_planets2_d10_121"]


@pytest.mark.parametrize("name", WHITTED)
def test_whitted_frames(orc, name):
    z, fs = load_golden(name)
    rgb, hit, q = orc.render_whitted(fs, z["cam"], z["X"], z["Y"], spp=1, max_bounces=int(z["max_bounces"]),
                                     miss=z["miss"], prenorm=bool(z["prenorm"]))
    assert np.array_equal(hit, z["hit"].astype(np.int32))
    assert np.array_equal(rgb, z["rgb"].astype(np.float64))      # integer-valued colours: exact
    assert q >= rgb.shape[0] * rgb.shape[1]
    if "image" in z:
        assert np.array_equal(orc.resolve(rgb, 1), z["image"])


def test_whitted_jittered_spp4(orc):
    """render_custom_scene with spp>1 (RL/output5.py:1463-1505), Philox-fed jitter."""
    z, fs = load_golden("whitted_balls_spp4_80x60")
    rgb, _, _ = orc.render_whitted(fs, z["cam"], z["X"], z["Y"], spp=int(z["spp"]), max_bounces=int(z["max_bounces"]),
                                   miss=z["miss"], seed=int(z["seed"]), prenorm=True)
    assert np.array_equal(orc.resolve(rgb, int(z["spp"])), z["image"])


@pytest.mark.parametrize("name", ["path_chandelier_48x27", "path_complex_48x27"])
def test_path_frames(orc, name):
    z, fs = load_golden(name)
    W, H, spp = int(z["W"]), int(z["H"]), int(z["spp"])
    sums, st = orc.render_path(fs, z["cam"], W, H, spp, int(z["max_bounces"]), float(z["mirror_threshold"]),
                               seed=int(z["seed"]))
    assert [st[k] for k in ("total_rays", "total_intersections", "light_hits", "small_light_hits")] == list(z["stats"])
    assert np.array_equal(sums, z["sums"].astype(np.float64))
    assert np.array_equal(orc.resolve(sums, spp), z["image"])
    # shard composition: tiles x sample ranges sum to the whole frame (what the multi-GPU split relies on)
    parts = np.zeros_like(sums)
    for rows in ((0, H // 2), (H // 2, H)):
        for smp in ((0, 1), (1, spp)):
            parts += orc.render_path(fs, z["cam"], W, H, spp, int(z["max_bounces"]), float(z["mirror_threshold"]),
                                     seed=int(z["seed"]), rows=rows, samples=smp)[0]
    assert np.array_equal(parts, sums)


ENVS = ["env_rl_optimized", "env_rl_demo", "env_fb_demo", "env_fb_balls", "env_rl_balls_rotated", "env_rl_adaptive"]


@pytest.mark.parametrize("name", ENVS)
def test_env_rollouts(orc, name):
    z, fs = load_golden(name)
    B = z["pixels"].shape[0]
    env = orc.OracleEnv(fs, B, int(z["width"]), int(z["height"]), camera=z["cam"], camera_angle=z["cam_angle"],
                        fov=float(z["fov"]), max_bounces=int(z["max_bounces"]), flavour=str(z["flavour"]),
                        adaptive=name.endswith("adaptive"))      # RL/train_raytracer_optimized.py AdaptiveRewardRayTracerEnv
    # The reference env hands the agent's float32 action straight into numpy trig (RL/ray_tracer_env.py:155-163),
    # so under NumPy>=2 promotion rules part of ITS arithmetic runs in float32; the double oracle therefore agrees
    # with it to float32 rounding (amplified by the trace), not to 1e-9.  Observations are float32 anyway.
    OBS = dict(rtol=5e-5, atol=5e-6)
    obs0 = env.reset(z["pixels"])
    np.testing.assert_allclose(obs0, z["obs0"], rtol=1e-6, atol=1e-7)
    for t in range(z["actions"].shape[0]):
        obs, rew, term, trunc, reason = env.step(z["actions"][t])
        np.testing.assert_allclose(obs, z["obs"][t], err_msg=f"obs step {t}", **OBS)
        np.testing.assert_allclose(rew, z["reward"][t], rtol=1e-5, atol=1e-6, err_msg=f"reward step {t}")
        assert np.array_equal(term, z["terminated"][t].astype(bool)), f"terminated step {t}"
        assert np.array_equal(trunc, z["truncated"][t].astype(bool)), f"truncated step {t}"
        assert np.array_equal(reason, z["reason"][t]), f"reason step {t}"


@pytest.mark.parametrize("name", ["simple_balls_a_64x48", "simple_balls_b_40x30"])
def test_output6_frames(orc, name):
    """FB/output6.py render_original_style (traditional mode) rendered by the reference itself with the Philox stream
    patched into np.random.random: per-pixel colours, float32 image and both counters, exactly."""
    z, fs = load_golden(name)
    rgb, st = orc.render_simple(fs, int(z["W"]), int(z["H"]), seed=int(z["seed"]), max_bounces=int(z["max_bounces"]))
    assert np.array_equal(rgb, z["rgb"])
    assert np.array_equal(np.minimum(1.0, rgb / 255.0).astype(np.float32), z["image"])
    assert [st["total_rays"], st["sun_hits"]] == list(z["stats"])
    # explicit rays = the scalar trace_ray_simple entry: the camera grid's own rays give the frame back
    W, H = int(z["W"]), int(z["H"])
    x, y = np.meshgrid(np.arange(W), np.arange(H))
    t = np.tan(np.pi / 6)
    d = np.stack([(x / W - 0.5) * 2.0 * (W / H) * t, (y / H - 0.5) * -2.0 * t, -np.ones_like(x, float)], axis=-1).reshape(-1, 3)
    d = d / np.sqrt((d * d).sum(axis=1, keepdims=True))       # render_original_style normalises before Ray() does
    rays = np.concatenate([np.tile([0.0, 0.0, 1.0], (W * H, 1)), d], axis=1)
    rgb2, _ = orc.render_simple(fs, W, H, seed=int(z["seed"]), max_bounces=int(z["max_bounces"]), rays=rays)
    assert (rgb2.reshape(H, W, 3) != rgb).any(axis=2).mean() < 0.002      # vnorm in numpy vs C: last-ulp flips only


def test_output6_lighting_helper(orc):
    """FB/output6.py calculate_lighting_exact_original on its own: 2,920 (intersection -> Colour) pairs recorded from the
    reference's own calls while it rendered the two frames above."""
    z, fs = load_golden("simple_lighting_balls")
    rgb, sun_hits = orc.simple_lighting(fs, z["hits"])
    assert np.array_equal(rgb, z["rgb"]) and sun_hits == 0
    # on the sun sphere (id 7) the helper returns the sun's colour and counts the hit (output6.py:204-206)
    sun = int(np.nonzero(fs.ids == 7)[0][0])
    row = np.array([[0, 0, 0, 0, 0, 1, sun]], float)
    rgb, sun_hits = orc.simple_lighting(fs, row)
    assert rgb.tolist() == [[255, 255, 204]] and sun_hits == 1


def test_fb_trajectories(orc):
    """FB/train_complex_only.py generate_trajectory run by the reference itself (random.* patched to the Philox
    stream): 256 random walks on the complex scene, every transition bit for bit."""
    z, fs = load_golden("traj_complex_256")
    o = orc.generate_trajectories(fs, 256, int(z["max_steps"]), int(z["max_bounces"]), int(z["seed"]))
    assert int(o["length"].sum()) > 1000 and int(o["hit_light"].sum()) > 5
    for k in ("length", "hit_light", "obs", "action", "next_obs", "reward", "hit"):
        assert np.array_equal(o[k], z[k]), k


def fb_test_policy(obs):
    """Stand-in for fb_agent.choose_direction (the one oracle/gen_golden.py used): IEEE basic ops only, float32."""
    o = np.asarray(obs, np.float32)
    a0 = np.clip(o[6] * np.float32(0.5) + o[7] * np.float32(0.25) - np.float32(0.125), np.float32(-1), np.float32(1))
    a1 = np.clip(o[8] * np.float32(0.5) + o[3] * np.float32(0.25) + o[16] * np.float32(0.5), np.float32(-1), np.float32(1))
    return np.array([a0, a1], dtype=np.float64)


def test_fb_guided_path_frame(orc):
    """WorkingFBRenderer.render of the reference itself (stand-in agent, np.random.random patched to the Philox
    streams): sums and all five counters exactly; and with no policy the FB renderer IS the traditional one."""
    z, fs = load_golden("path_fb_complex_40x24")
    args = (z["cam"], int(z["W"]), int(z["H"]), int(z["spp"]), int(z["max_bounces"]), float(z["mirror_threshold"]))
    sums, st = orc.render_path_fb(fs, *args, fb_test_policy, float(z["fb_usage_prob"]), int(z["seed"]))
    assert np.array_equal(sums, z["sums"])
    assert [st[k] for k in ("total_rays", "total_intersections", "light_hits", "small_light_hits", "fb_used")] == list(z["stats"])
    assert st["fb_used"] > 1000
    plain, st0 = orc.render_path_fb(fs, *args, None, 0.0, int(z["seed"]))
    trad, st1 = orc.render_path(fs, *args, seed=int(z["seed"]))
    assert np.array_equal(plain, trad) and st0["total_rays"] == st1["total_rays"] and st0["fb_used"] == 0


def test_oracle_mean_image_agrees_with_the_reference_rng(orc):
    """Statistical pin of the stochastic path: the unmodified reference, drawing from numpy's own MT19937 (golden made by
    oracle/gen_golden.py path_native), and the oracle on the Philox stream give the same mean image within the noise
    two seeds of the oracle show against each other."""
    z, fs = load_golden("path_chandelier_native_rng_40x24")
    W, H, spp = int(z["W"]), int(z["H"]), int(z["spp"])
    ref = z["image"].astype(np.float64) * 255.0
    imgs, rays = [], []
    for seed in (1, 2):
        sums, st = orc.render_path(fs, z["cam"], W, H, spp, int(z["max_bounces"]), float(z["mirror_threshold"]), seed=seed)[:2]
        imgs.append(np.minimum(255.0, np.floor(np.asarray(sums)[..., :3] / spp)))
        rays.append(st["total_rays"])
    rm = lambda a, b: float(np.sqrt(np.mean((a - b) ** 2)))      # noqa: E731
    own = rm(imgs[0], imgs[1])
    for im in imgs:
        assert rm(ref, im) < 1.25 * own, (rm(ref, im), own)
        assert np.abs((ref - im).mean(axis=(0, 1))).max() < 0.25
    assert abs(np.mean(rays) - int(z["stats"][0])) < 0.01 * int(z["stats"][0])
