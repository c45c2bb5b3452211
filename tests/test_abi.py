"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/rt_b200.h declares (no compute calls -- there is no GPU here), and fails loudly without a device."""
import os
import re

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
