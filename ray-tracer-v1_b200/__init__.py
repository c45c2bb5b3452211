"""b200-raytrace: the traditional sphere ray-tracing inner loop of
JoaquinRodriguezph/ray-tracer-v1 as hand-written sm_100a CUDA behind the
reference's Python scene API.

Scene API (host, drop-in):  Vector, Angle, Colour, Material, Sphere, GlobalLight, PointLight
Hot path (CUDA via ctypes): Ray / Intersection, TraditionalRenderer, CustomSceneExperiment,
                            RayTracerEnv, BatchedRayTracerEnv
There is no CPU fallback: any tracing call raises ``NativeLibraryError`` if
``csrc/librt_b200.so`` is missing or no CUDA device is present.
"""
from .vector import Vector, Angle
from .colour import Colour
from .material import Material
from .object import Sphere
from .light import GlobalLight, PointLight, incidence
from .scene import FlatScene, flatten_scene

__version__ = "0.1.0"

_LAZY = {
    "Ray": "ray", "Intersection": "ray",
    "TraditionalRenderer": "renderers", "ComplexTraditionalRenderer": "renderers", "CustomSceneExperiment": "renderers",
    "SimplifiedFBRenderer": "renderers", "save_png": "renderers",
    "WorkingFBRenderer": "renderers", "render_path_wavefront": "renderers",
    "render_whitted": "renderers", "render_path": "renderers",
    "RayTracerEnv": "ray_tracer_env", "BatchedRayTracerEnv": "ray_tracer_env", "FBRayTracerEnv": "ray_tracer_env",
    "AdaptiveRewardRayTracerEnv": "ray_tracer_env", "RayTracerVecEnv": "ray_tracer_env",
    "generate_trajectories": "fb_trajectories", "generate_trajectory": "fb_trajectories", "TrajectoryBatch": "fb_trajectories",
    "NativeLibraryError": "_native", "DeviceScene": "_native",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(name)
