"""Batched FB training trajectories: the random-walk experience generator of FB/train_complex_only.py on the GPU.

Reference: ``RayTracedComplexTrainer.generate_trajectory`` (FB/train_complex_only.py:254-334) with its helpers
``random_point_on_sphere`` (:54-66), ``sample_cosine_weighted_direction`` (:69-96), ``direction_to_action`` (:99-127),
``create_observation`` (:130-150) and ``nearest_intersection`` (:153-166).  The reference walks ONE path per call in
Python and hands every transition to ``agent.record_success``; here ``n`` paths are walked by one kernel launch
(``rt_generate_trajectories``) and the transitions come back as CUDA tensors, ready for a replay buffer.

Randomness: the reference draws from Python's global ``random``; here every draw is Philox keyed (trajectory, draw
slot) with ``seed``, so a batch is reproducible and independent of how it is split across launches or GPUs.
"""
import numpy as np

from . import _native as nat
from .scene import flatten_scene

__all__ = ["TrajectoryBatch", "generate_trajectories", "generate_trajectory", "create_observation", "direction_to_action"]

OBS_DIM = 22


class TrajectoryBatch:
    """Padded transitions of ``n`` trajectories (torch CUDA tensors): obs / next_obs [n,S,22] f32, action [n,S,2] f32,
    reward [n,S] f32, hit [n,S] bool, length [n] int32 (valid transitions per trajectory), hit_light [n] bool."""

    def __init__(self, obs, action, next_obs, reward, hit, length, hit_light, queries):
        self.obs, self.action, self.next_obs, self.reward = obs, action, next_obs, reward
        self.hit, self.length, self.hit_light, self.queries = hit, length, hit_light, queries

    def valid_mask(self):
        import torch
        steps = torch.arange(self.obs.shape[1], device=self.obs.device)
        return steps[None, :] < self.length[:, None]

    def flat(self):
        """The valid transitions only, concatenated: (obs [m,22], action [m,2], next_obs [m,22], reward [m], hit [m])."""
        m = self.valid_mask()
        return self.obs[m], self.action[m], self.next_obs[m], self.reward[m], self.hit[m]

    def transitions(self, j):
        """Trajectory ``j`` as the reference returns it: list of (obs, action, next_obs, reward, hit_light) numpy tuples."""
        k = int(self.length[j])
        o, a, no = self.obs[j, :k].cpu().numpy(), self.action[j, :k].cpu().numpy(), self.next_obs[j, :k].cpu().numpy()
        r, h = self.reward[j, :k].cpu().numpy(), self.hit[j, :k].cpu().numpy()
        return [(o[t].copy(), a[t].copy(), no[t].copy(), float(r[t]), bool(h[t])) for t in range(k)]


def generate_trajectories(spheres, n, max_steps=8, max_bounces=None, seed=0, precision="f32", device=0, scene=None):
    """Walk ``n`` random paths over ``spheres`` (a list of ``Sphere`` or an already flattened scene).

    max_steps    transitions per trajectory at most (the reference passes ``self.max_bounces`` = 8)
    max_bounces  normaliser of the observation's bounce entry (``self.max_bounces``); defaults to max_steps
    scene        an existing ``DeviceScene`` to reuse (then ``spheres`` is ignored)"""
    import torch
    nat.lib()
    prec = nat.F64 if precision in ("f64", "fp64", "float64", "double", nat.F64) and precision != nat.F32 else nat.F32
    own = scene is None
    if own:
        fs = spheres if hasattr(spheres, "radius") and hasattr(spheres, "centre") else flatten_scene(spheres)
        scene = nat.DeviceScene(fs, device)
    try:
        dev = torch.device("cuda", scene.device)
        n, S = int(n), int(max_steps)
        mb = int(max_bounces if max_bounces is not None else max_steps)
        obs = torch.zeros((n, S, OBS_DIM), dtype=torch.float32, device=dev)
        nxt = torch.zeros((n, S, OBS_DIM), dtype=torch.float32, device=dev)
        act = torch.zeros((n, S, 2), dtype=torch.float32, device=dev)
        rew = torch.zeros((n, S), dtype=torch.float32, device=dev)
        hit = torch.zeros((n, S), dtype=torch.uint8, device=dev)
        length = torch.zeros(n, dtype=torch.int32, device=dev)
        lit = torch.zeros(n, dtype=torch.uint8, device=dev)
        stats = torch.zeros(8, dtype=torch.int64, device=dev)
        nat.check(nat.lib().rt_generate_trajectories(scene.handle, prec, n, S, mb, int(seed) & (2 ** 64 - 1), obs.data_ptr(),
                                                     act.data_ptr(), nxt.data_ptr(), rew.data_ptr(), hit.data_ptr(),
                                                     length.data_ptr(), lit.data_ptr(), stats.data_ptr(), None))
        torch.cuda.synchronize(dev)
    finally:
        if own:
            scene.close()
    return TrajectoryBatch(obs, act, nxt, rew, hit.bool(), length, lit.bool(), int(stats[4]))


def generate_trajectory(spheres, max_steps=8, max_bounces=None, seed=0, precision="f32", device=0):
    """One trajectory, with the reference method's return value: (transitions, hit_light)."""
    b = generate_trajectories(spheres, 1, max_steps, max_bounces, seed, precision, device)
    return b.transitions(0), bool(b.hit_light[0])


# ---- host-side formatting helpers of the reference (no tracing in them) --------------------------------------------
def create_observation(point, normal, incoming_dir, bounce_count, color, material, sphere_id, max_bounces):
    """22-dim observation (FB/train_complex_only.py:130-150)."""
    return np.array([
        point.x, point.y, point.z, incoming_dir.x, incoming_dir.y, incoming_dir.z, normal.x, normal.y, normal.z,
        float(getattr(material, 'reflective', False)), float(getattr(material, 'transparent', False)),
        float(getattr(material, 'emitive', False)), float(getattr(material, 'refractive_index', 1.0)),
        color.r / 255.0, color.g / 255.0, color.b / 255.0, float(bounce_count) / max_bounces, 0.0,
        float(sphere_id) / 100.0, 0.5, 0.5, 0.5], dtype=np.float32)


def direction_to_action(direction, normal):
    """World direction -> (theta, phi) action in [-1,1]^2 (FB/train_complex_only.py:99-127)."""
    import math
    from .vector import Vector
    if abs(normal.z) < 0.999:
        tangent = Vector(0, 0, 1).crossProduct(normal).normalise()
    else:
        tangent = Vector(1, 0, 0).crossProduct(normal).normalise()
    bitangent = normal.crossProduct(tangent).normalise()
    lx, ly, lz = direction.dotProduct(tangent), direction.dotProduct(bitangent), direction.dotProduct(normal)
    theta = min(math.acos(max(-1, min(1, lz))), math.pi / 2)
    phi = math.atan2(ly, lx)
    return np.array([(theta / (math.pi / 2)) * 2 - 1, phi / math.pi], dtype=np.float32)
