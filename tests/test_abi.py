"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/rt_b200.h declares (no compute calls -- there is no GPU here), and fails loudly without a device."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def native():
    from ray_tracer_v1_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        _native.build()
    return _native


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(native):
    names = declared_symbols()
    assert len(names) >= 30
    lib = native.load_symbols()
    for n in names:
        assert hasattr(lib, n), f"librt_b200.so does not export {n}"
    assert sorted(native.SIGNATURES) == names, "ctypes SIGNATURES and include/rt_b200.h disagree"
    assert lib.rt_version() >= 100


def test_struct_layouts_match_header(native):
    """ctypes mirrors of the parameter blocks have the C sizes (LP64)."""
    import ctypes as C
    assert C.sizeof(native.PathParams) == 96
    assert C.sizeof(native.WhittedParams) == 120
    assert C.sizeof(native.EnvDesc) == 104
    assert C.sizeof(native.SceneDesc) == 216
    assert C.sizeof(native.SimpleParams) == 120
    assert C.sizeof(native.PathSink) == 4 * 4 + 8 + 16 * 8 + 17 * 4 + 4 * 4 + 4 + 16 * 8 + 8 + 4 * 4      # 392: flags[] is 8-aligned


def test_no_cpu_fallback(native):
    """Without a CUDA device every product entry raises; nothing routes through the oracle."""
    if native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(native.NativeLibraryError):
        native.lib()
    import ray_tracer_v1_b200 as pkg
    from ray_tracer_v1_b200 import scenes
    spec = scenes.build_balls_in_space()
    fs = pkg.flatten_scene(spec.spheres, spec.global_lights, spec.point_lights, spec.background)
    with pytest.raises(native.NativeLibraryError):
        native.DeviceScene(fs)
    src = "".join(open(os.path.join(ROOT, "ray-tracer-v1_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "ray-tracer-v1_b200")) if f.endswith(".py"))
    assert "oracle" not in src.replace("oracle/", "").replace("the oracle", "").replace("CPU oracle", ""), \
        "the product package must not import the oracle"


def test_scene_descriptor_marshalling_is_faithful():
    """make_desc packs the float64 / int32 arrays of a FlatScene into two buffers and hands the C ABI offsets into them:
    every pointer must read back exactly the array it stands for, for scenes with and without lights, and the
    cached-constant-lights path of the output5 entries must flatten to the same scene as the plain path."""
    import ctypes as C
    import ray_tracer_v1_b200 as pkg
    from ray_tracer_v1_b200 import _native as native, scenes
    from ray_tracer_v1_b200.renderers import CustomSceneExperiment
    balls = scenes.build_balls_in_space(as_rendered=True)
    plain = pkg.flatten_scene(balls.spheres, balls.global_lights, balls.point_lights, balls.background)
    fast = CustomSceneExperiment._as_rendered(scenes.build_balls_in_space(as_rendered=False).spheres)
    for name in native._DESC_F + native._DESC_I + ("bg",):
        if not name.startswith("l_"):
            assert np.array_equal(getattr(fast, name), getattr(plain, name)), name
    cases = [plain, fast, pkg.flatten_scene(scenes.build_chandelier().spheres),
             pkg.flatten_scene([], background_colour=balls.background)]
    for fs in cases:
        d, keep = native.make_desc(fs)
        assert (d.n, d.nG, d.nP, d.nL) == (fs.radius.shape[0], fs.g_strength.shape[0], fs.p_strength.shape[0], fs.l_index.shape[0])
        for names, ct, dt in ((native._DESC_F, C.c_double, np.float64), (native._DESC_I, C.c_int32, np.int32)):
            for name in names:
                want = np.ascontiguousarray(getattr(fs, name), dt).reshape(-1)
                if want.size:
                    got = np.ctypeslib.as_array(C.cast(getattr(d, name), C.POINTER(ct)), shape=(want.size,))
                    assert np.array_equal(got, want), name
        assert [d.bg[0], d.bg[1], d.bg[2]] == [float(v) for v in np.asarray(fs.bg).reshape(3)]


def test_host_side_frame_logic():
    """Host logic that decides what crosses PCIe and in how many launches, no GPU needed: the content key that lets the
    sub-millisecond frame entries skip the upload of an unchanged scene, and the row bands of the pipelined read-back."""
    import ray_tracer_v1_b200 as pkg
    from ray_tracer_v1_b200 import _native as native, scenes
    from ray_tracer_v1_b200.frames import FrameContext
    spheres = scenes.build_balls_in_space(as_rendered=False).spheres
    a, b = pkg.flatten_scene(spheres), pkg.flatten_scene(spheres)
    assert native.scene_signature(a) == native.scene_signature(b)             # re-flattened, unchanged
    spheres[1].centre = pkg.Vector(0.4, 0.1, -2.5)
    assert native.scene_signature(pkg.flatten_scene(spheres)) != native.scene_signature(a)
    spheres[1].centre = a_centre = pkg.Vector(*a.centre[1])
    spheres[2].colour = pkg.Colour(1, 2, 3)
    assert native.scene_signature(pkg.flatten_scene(spheres)) != native.scene_signature(a)
    assert a_centre.x == a.centre[1][0]
    # bands: stripe-aligned, contiguous, covering [y0, y1); small or short frames stay one launch
    ctx = FrameContext.__new__(FrameContext)
    bands = ctx._bands(1920, 0, 1080, 64)
    assert len(bands) == FrameContext.BANDS and bands[0][0] == 0 and bands[-1][1] == 1080
    assert all(b0 % 8 == 0 for b0, _ in bands) and all(x[1] == y[0] for x, y in zip(bands[:-1], bands[1:]))
    sizes = [b1 - b0 for b0, b1 in bands]
    assert sizes == sorted(sizes, reverse=True) and sizes[-1] * 8 <= 1080 + 64      # tapered: the exposed last copy is the smallest
    # a band's copy (12 B/pixel at ~50 GB/s) hides behind the next band's render (~8 ns per pixel at 64 spp)
    assert all(a * 12 / 50e9 < b * 8e-9 for a, b in zip(sizes[:-1], sizes[1:]))
    assert ctx._bands(1920, 0, 1080, 4) is None and ctx._bands(320, 0, 240, 4096) is None
    ragged = ctx._bands(1920, 3, 1077, 64)
    assert ragged[0][0] == 3 and ragged[-1][1] == 1077 and all(x[1] == y[0] for x, y in zip(ragged[:-1], ragged[1:]))


def test_package_exports_resolve():
    """Every name the package advertises lazily imports without a GPU (the classes only touch the library when they trace),
    and the reference's entry-point names are among them."""
    import ray_tracer_v1_b200 as pkg
    for name in pkg._LAZY:
        assert getattr(pkg, name) is not None, name
    for name in ("Ray", "Intersection", "TraditionalRenderer", "ComplexTraditionalRenderer", "WorkingFBRenderer",
                 "CustomSceneExperiment", "SimplifiedFBRenderer", "RayTracerEnv", "FBRayTracerEnv",
                 "AdaptiveRewardRayTracerEnv", "RayTracerVecEnv", "BatchedRayTracerEnv"):
        assert name in pkg._LAZY, name
    with pytest.raises(AttributeError):
        pkg.no_such_name
