"""``RayTracerEnv`` -- the reference's Gymnasium environment (RL/ray_tracer_env.py:21-425 and the FB flavour
FB/ray_tracer_env.py:21-538) with ``reset`` / ``step`` running on the GPU, plus the batched form the B200 path is
built for.

* ``BatchedRayTracerEnv``  B episodes stepped by ONE kernel launch (``rt_env_step``): SoA state in HBM, actions in and
  observations / rewards / done flags out as torch CUDA tensors that alias the device buffers -- nothing crosses PCIe,
  the policy network consumes them in place.  ``BatchedRayTracerEnv.shard`` gives each rank of a torchrun job its
  contiguous slice of the environments (no collective: episodes are independent).
* ``RayTracerEnv`` / ``FBRayTracerEnv``  the scalar drop-ins (one episode, numpy observation, python reward, info dict
  with the reference's keys), implemented as a batch of one.

Semantics restated from the reference (all verified against recorded rollouts of the unmodified env, tests/golden/):
18-float observation (:184-222), pinhole first ray (:121-142), action -> direction in the TBN frame of the current
normal (:144-182; FB maps [-1,1]^2 to the hemisphere, FB :171-172), the ray leaves the exact hit point with the hit
sphere's id suppressed (:343-359), RL reward = brightness of terminalRGB at the PRE-step hit - 0.01 * bounces (:362),
FB reward = 10 on sphere id 7 / lighting reward with one shadow test / -0.1 on a miss (FB :241-336, :417-476).
"""
import ctypes as C

import numpy as np

from . import _native as nat
from .colour import Colour
from .scene import flatten_scene
from .vector import Angle, Vector

__all__ = ["AdaptiveRewardRayTracerEnv", "RayTracerVecEnv", "BatchedRayTracerEnv", "RayTracerEnv", "FBRayTracerEnv", "Box"]

OBS_DIM = 18


class Box:
    """Minimal stand-in for ``gymnasium.spaces.Box`` (gymnasium is optional; when importable the real one is used)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low, self.high, self.dtype = np.asarray(low, dtype), np.asarray(high, dtype), np.dtype(dtype)
        self.shape = self.low.shape
        self._rs = np.random.RandomState()

    def seed(self, seed=None):
        self._rs = np.random.RandomState(seed)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1e3)
        hi = np.where(np.isfinite(self.high), self.high, 1e3)
        return self._rs.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


def _spaces(max_bounces, flavour):
    try:
        from gymnasium.spaces import Box as B
    except Exception:
        B = Box
    inf = np.inf
    low = np.array([-inf] * 3 + [-1] * 6 + [0, 0, 0, 1] + [0, 0, 0] + [0, 0], np.float32)
    high = np.array([inf] * 3 + [1] * 6 + [1, 1, 1, 3] + [1, 1, 1] + [max_bounces, max_bounces], np.float32)
    obs = B(low=low, high=high, dtype=np.float32)
    if flavour == "fb":
        act = B(low=np.array([-1.0, -1.0], np.float32), high=np.array([1.0, 1.0], np.float32), dtype=np.float32)
    else:
        act = B(low=np.array([0.0, 0.0], np.float32), high=np.array([np.pi / 2, 2 * np.pi], np.float32), dtype=np.float32)
    return obs, act


def _xyz(v):
    return (float(v.x), float(v.y), float(v.z)) if hasattr(v, "x") else tuple(float(c) for c in v)


class BatchedRayTracerEnv:
    """B independent RayTracerEnv episodes on one GPU.

    spheres / lights          the Python scene graph (re-flattened and re-uploaded at every ``reset`` -- scenes are
                              mutable lists in the reference) or an already flattened ``FlatScene``
    flavour                   'rl' (RL/ray_tracer_env.py) or 'fb' (FB/ray_tracer_env.py, sun id 7)
    precision                 'f32' product path / 'f64' parity build
    """

    def __init__(self, spheres, n_envs, image_width=800, image_height=600, camera_position=Vector(0, 0, 0),
                 camera_angle=Angle(0, 0, 0), fov=90, max_bounces=5, background_colour=Colour(0, 0, 0),
                 global_light_sources=None, point_light_sources=None, flavour="rl", sun_id=7, precision="f32",
                 device=0, seed=0, reward_mode="default", light_ids=(99, 100)):
        import torch
        self.torch = torch
        self.spheres = spheres
        self.n_envs = int(n_envs)
        self.image_width, self.image_height = int(image_width), int(image_height)
        self.camera_position, self.camera_angle = camera_position, camera_angle
        self.fov, self.max_bounces = fov, int(max_bounces)
        self.background_colour = background_colour
        self.global_light_sources = global_light_sources if global_light_sources is not None else []
        self.point_light_sources = point_light_sources if point_light_sources is not None else []
        if flavour not in ("rl", "fb"):
            raise ValueError("flavour must be 'rl' or 'fb'")
        self.flavour, self.sun_id = flavour, int(sun_id)
        if reward_mode not in ("default", "adaptive"):
            raise ValueError("reward_mode must be 'default' or 'adaptive'")
        if reward_mode == "adaptive" and flavour != "rl":
            raise ValueError("the adaptive reward (RL/train_raytracer_optimized.py) shapes the RL flavour's reward")
        self.reward_mode, self.light_ids = reward_mode, (int(light_ids[0]), int(light_ids[1]))
        from .renderers import _precision
        self.precision = _precision(precision)        # unknown strings raise ValueError
        self.device = int(device)
        self.seed = int(seed)
        self.observation_space, self.action_space = _spaces(self.max_bounces, flavour)
        nat.lib()
        dev = torch.device("cuda", self.device)
        B = self.n_envs
        self.obs = torch.zeros((B, OBS_DIM), dtype=torch.float32, device=dev)
        self.reward = torch.zeros(B, dtype=torch.float64, device=dev)
        self.terminated = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.truncated = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.reason = torch.zeros(B, dtype=torch.int32, device=dev)
        self.info = torch.zeros((B, 4), dtype=torch.float64, device=dev)
        self.pixels = torch.zeros((B, 2), dtype=torch.int32, device=dev)
        self.stats = torch.zeros(8, dtype=torch.int64, device=dev)
        self._actions = torch.zeros((B, 2), dtype=torch.float32, device=dev)
        self.scene = None
        self.handle = None
        self._resets = 0
        self._step_args = None
        self._desc_key = None
        self.env_offset = 0

    # ---- sharding ------------------------------------------------------------------------------------------------
    @classmethod
    def shard(cls, spheres, n_envs_total, rank=None, world=None, **kw):
        """This rank's slice of ``n_envs_total`` environments (env-sharded rollouts): no collective."""
        from .distributed import env_slices
        if rank is None or world is None:
            import torch.distributed as dist
            rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        b0, b1 = env_slices(n_envs_total, world)[rank]
        env = cls(spheres, b1 - b0, **kw)
        env.env_offset = b0          # device-drawn start pixels are keyed by the GLOBAL env index: shard == rows of the whole
        return env

    # ---- scene / handle ------------------------------------------------------------------------------------------
    def _flat(self):
        if hasattr(self.spheres, "radius") and hasattr(self.spheres, "centre"):
            return self.spheres
        return flatten_scene(self.spheres, self.global_light_sources, self.point_light_sources, self.background_colour)

    def _ensure(self):
        fs = self._flat()
        if self.scene is None:
            self.scene = nat.DeviceScene(fs, self.device)
        else:
            self.scene.update(fs)
        # the reference re-reads camera / fov / max_bounces / image size on every reset (RL/ray_tracer_env.py:254-293):
        # a changed descriptor gets a fresh device env (the episode state is re-initialised by the reset anyway)
        desc_key = (self.n_envs, int(self.image_width), int(self.image_height), _xyz(self.camera_position),
                    _xyz(self.camera_angle), float(self.fov), int(self.max_bounces), self.flavour, int(self.sun_id),
                    self.reward_mode, tuple(int(v) for v in self.light_ids), int(self.env_offset))
        if self.handle is not None and desc_key != self._desc_key:
            nat.load_symbols().rt_env_destroy(self.handle)
            self.handle = None
            self._graph = None                         # a captured step launch holds the old handle's pointers
        if self.handle is None:
            d = nat.EnvDesc()
            d.B, d.W, d.H = self.n_envs, int(self.image_width), int(self.image_height)
            d.cam[:] = _xyz(self.camera_position)
            d.cam_angle[:] = _xyz(self.camera_angle)
            d.fov, d.max_bounces = float(self.fov), int(self.max_bounces)
            d.flavour, d.sun_id = (nat.ENV_FB if self.flavour == "fb" else nat.ENV_RL), int(self.sun_id)
            d.reward_mode = 1 if self.reward_mode == "adaptive" else 0
            d.light_ids[:] = [int(v) for v in self.light_ids]
            d.env_offset = int(self.env_offset)
            h = C.c_void_p()
            nat.check(nat.lib().rt_env_create(self.scene.handle, self.precision, C.byref(d), C.byref(h)))
            self.handle = h.value
            self._desc_key = desc_key

    def _dev_tensor(self, x, dtype, shape):
        torch = self.torch
        if isinstance(x, torch.Tensor):
            t = x.to(device=self.obs.device, dtype=dtype)
        else:
            t = torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device=self.obs.device)
        return t.reshape(shape).contiguous()

    # ---- Gymnasium-style API -------------------------------------------------------------------------------------
    def reset(self, seed=None, options=None, mask=None):
        """Start new episodes (all, or those with ``mask`` != 0).  options={'pixels': [B,2] (x, y)} fixes the start
        pixels, otherwise they are drawn on the device with Philox(seed).  -> (obs [B,18] CUDA tensor, info dict)."""
        if seed is not None:
            self.seed = int(seed)
        if mask is None or self.handle is None:
            self._ensure()            # full reset: the scene list may have been mutated since (re-flatten + upload)
            # a captured step_auto launch has the scene's device pointers, its census and the seed baked into its
            # parameter block: re-captured at the next step_auto(graph=True)
            self._graph = None
        torch = self.torch
        pix = None
        if options is not None and "pixels" in options:
            pix = self._dev_tensor(options["pixels"], torch.int32, (self.n_envs, 2))
        m = None if mask is None else self._dev_tensor(mask, torch.uint8, (self.n_envs,))
        key = (self.seed + 0x632BE59BD9B4E019 * self._resets) & (2 ** 64 - 1)
        self._resets += 1
        nat.check(nat.lib().rt_env_reset(self.handle, nat._ptr(pix), nat._ptr(m), key, self.obs.data_ptr(),
                                         self.pixels.data_ptr(), None))
        return self.obs, {"pixels": self.pixels}

    def step(self, actions):
        """actions [B,2] (torch CUDA tensor, consumed in place, or array-like) ->
        (obs [B,18] f32, reward [B] f64, terminated [B] bool, truncated [B] bool, info dict of tensors).
        The returned tensors alias the env's device buffers and are overwritten by the next call."""
        torch = self.torch
        if isinstance(actions, torch.Tensor) and actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous():
            a = actions.reshape(self.n_envs, 2)
        else:
            a = self._dev_tensor(actions, torch.float32, (self.n_envs, 2))
        self._last_actions = a
        if self._step_args is None:     # the output buffers never move: resolve their pointers and views once
            self._step_args = (self.obs.data_ptr(), self.reward.data_ptr(), self.terminated.data_ptr(),
                               self.truncated.data_ptr(), self.reason.data_ptr(), self.info.data_ptr(), self.stats.data_ptr())
            self._info = {"bounce_count": self.info[:, 0], "through_count": self.info[:, 1], "total_reward": self.info[:, 2],
                          "hit_sun": self.info[:, 3], "reason": self.reason}
            self._flags = (self.terminated.view(torch.bool), self.truncated.view(torch.bool))
            self._step_fn = nat.lib().rt_env_step
        rc = self._step_fn(self.handle, a.data_ptr(), *self._step_args, None)
        if rc:
            nat.check(rc)
        return self.obs, self.reward, self._flags[0], self._flags[1], dict(self._info)

    # ---- one launch per step: step + restart of finished episodes ------------------------------------------------
    def _auto_buffers(self):
        torch = self.torch
        if getattr(self, "_auto", None) is None:
            dev, B = self.obs.device, self.n_envs
            ft = torch.float64 if self.precision == nat.F64 else torch.float32
            self.actions = torch.zeros((B, 2), dtype=torch.float32, device=dev)       # static: a policy may write it in place
            self.reward_auto = torch.zeros(B, dtype=ft, device=dev)
            self.info_auto = torch.zeros((B, 4), dtype=ft, device=dev)
            self.final_obs = torch.zeros((B, OBS_DIM), dtype=torch.float32, device=dev)
            self._auto = True
            self._auto_info = {"bounce_count": self.info_auto[:, 0], "through_count": self.info_auto[:, 1],
                               "total_reward": self.info_auto[:, 2], "hit_sun": self.info_auto[:, 3], "reason": self.reason,
                               "terminal_observation": self.final_obs, "pixels": self.pixels}
            self._auto_flags = (self.terminated.view(torch.bool), self.truncated.view(torch.bool))
            self._graph = None

    def _launch_auto(self, stream=None):
        rc = nat.lib().rt_env_step_auto(self.handle, self.actions.data_ptr(), self.obs.data_ptr(), self.reward_auto.data_ptr(),
                                        self.terminated.data_ptr(), self.truncated.data_ptr(), self.reason.data_ptr(),
                                        self.info_auto.data_ptr(), self.final_obs.data_ptr(), self.pixels.data_ptr(),
                                        self.seed & (2 ** 64 - 1), self.stats.data_ptr(), stream)
        if rc:
            nat.check(rc)

    def step_auto(self, actions=None, graph=False):
        """One step of every episode AND the restart of the finished ones in ONE kernel launch (``rt_env_step_auto``: the
        VecEnv protocol of Stable-Baselines3).  -> (obs, reward, terminated, truncated, info): ``obs`` [B,18] holds the
        first observation of the new episode where an episode ended (its last one is in
        ``info['terminal_observation']``), reward is float32 for the FP32 env.  ``actions=None`` steps on
        ``self.actions`` as it stands (a policy network can write its output there in place).  ``graph=True`` replays
        the launch from a CUDA graph captured at the first call (nothing in the launch changes from step to step: new
        start pixels are keyed by the per-env episode counter on the device)."""
        torch = self.torch
        if self.handle is None:
            raise RuntimeError("reset() first")
        self._auto_buffers()
        if actions is not None:
            if not (isinstance(actions, torch.Tensor) and actions.data_ptr() == self.actions.data_ptr()):
                self.actions.copy_(self._dev_tensor(actions, torch.float32, (self.n_envs, 2)))
        if graph:
            if self._graph is None:
                side = torch.cuda.Stream(device=self.obs.device)
                side.wait_stream(torch.cuda.current_stream(self.obs.device))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    self._launch_auto(torch.cuda.current_stream(self.obs.device).cuda_stream)
                self._graph = g
            self._graph.replay()
        else:
            self._launch_auto(None)
        return self.obs, self.reward_auto, self._auto_flags[0], self._auto_flags[1], self._auto_info

    def close(self):
        self._graph = None
        if self.handle is not None:
            try:
                nat.load_symbols().rt_env_destroy(self.handle)
            finally:
                self.handle = None
        if self.scene is not None:
            self.scene.close()
            self.scene = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Hit:
    """What the scalar env exposes as ``current_intersection`` (read-only view of the device state)."""

    def __init__(self, obs, sphere):
        self.intersects = True
        self.point = Vector(float(obs[0]), float(obs[1]), float(obs[2]))
        self.normal = Vector(float(obs[6]), float(obs[7]), float(obs[8]))
        self.object = sphere


try:                                    # be a real gymnasium.Env when gymnasium is installed (it is optional)
    import gymnasium as _gym
    _EnvBase = _gym.Env
except Exception:                       # pragma: no cover - depends on the image
    _EnvBase = object


class RayTracerEnv(_EnvBase):
    """Scalar drop-in for ``RL/ray_tracer_env.RayTracerEnv`` (same constructor, ``reset`` / ``step`` / ``render``
    signatures and info keys); the tracing runs on the GPU as a batch of one."""

    metadata = {"render_modes": ["rgb_array"], "render_fps": 30}
    flavour = "rl"
    reward_mode = "default"

    def __init__(self, spheres=None, image_width=800, image_height=600, camera_position=Vector(0, 0, 0),
                 camera_angle=Angle(0, 0, 0), fov=90, max_bounces=5, background_colour=Colour(0, 0, 0),
                 global_light_sources=None, point_light_sources=None, render_mode=None, precision="f64", device=0):
        if _EnvBase is not object:
            super().__init__()
        self.spheres = spheres if spheres is not None else []
        self.image_width, self.image_height = image_width, image_height
        self.camera_position, self.camera_angle = camera_position, camera_angle
        self.fov, self.max_bounces = fov, max_bounces
        self.background_colour = background_colour
        self.global_light_sources = global_light_sources if global_light_sources is not None else []
        self.point_light_sources = point_light_sources if point_light_sources is not None else []
        self.render_mode = render_mode
        self.current_ray = None
        self.current_intersection = None
        self.current_pixel = None
        self.accumulated_color = Colour(0, 0, 0)
        self.bounce_count = 0
        self.through_count = 0
        self.total_reward = 0.0
        self.observation_space, self.action_space = _spaces(max_bounces, self.flavour)
        self._np_random = np.random.RandomState()
        self._batch = BatchedRayTracerEnv(self.spheres, 1, image_width, image_height, camera_position, camera_angle, fov,
                                          max_bounces, background_colour, self.global_light_sources,
                                          self.point_light_sources, flavour=self.flavour, precision=precision,
                                          device=device, reward_mode=self.reward_mode,
                                          light_ids=getattr(self, "light_ids", (99, 100)))

    # pinhole camera of the reference (RL/ray_tracer_env.py:121-142), host-side copy for info['initial_ray'] only
    def _get_initial_ray(self, pixel_x, pixel_y):
        aspect = self.image_width / self.image_height
        fov_rad = self.fov * np.pi / 180
        px = (2 * (pixel_x + 0.5) / self.image_width - 1) * aspect * np.tan(fov_rad / 2)
        py = (1 - 2 * (pixel_y + 0.5) / self.image_height) * np.tan(fov_rad / 2)
        d = Vector(px, py, -1).normalise()
        a = self.camera_angle
        if a.x != 0 or a.y != 0 or a.z != 0:
            d = d.rotate(a)
        return self.camera_position, d.normalise()

    def _sync_state(self, obs):
        b = self._batch
        self._batch.spheres = self.spheres
        hit = bool(np.any(obs != 0))
        if hit:
            # the sphere under the ray: identified by its material / position is ambiguous, so ask the device
            self.current_intersection = _Hit(obs, None)
        else:
            self.current_intersection = None
        self.accumulated_color = Colour(float(obs[13]) * 255.0, float(obs[14]) * 255.0, float(obs[15]) * 255.0)

    def reset(self, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.RandomState(seed)
        self.bounce_count = 0
        self.through_count = 0
        self.accumulated_color = Colour(0, 0, 0)
        self.total_reward = 0.0
        if options is not None and 'pixel' in options:
            self.current_pixel = options['pixel']
        else:
            self.current_pixel = (int(self._np_random.randint(0, self.image_width)),
                                  int(self._np_random.randint(0, self.image_height)))
        b = self._batch
        b.spheres, b.global_light_sources, b.point_light_sources = self.spheres, self.global_light_sources, self.point_light_sources
        b.background_colour = self.background_colour
        obs, _ = b.reset(options={"pixels": np.array([self.current_pixel], np.int32)})
        observation = obs[0].cpu().numpy().copy()
        origin, d = self._get_initial_ray(*self.current_pixel)
        self._sync_state(observation)
        info = {
            'pixel': self.current_pixel,
            'bounce_count': self.bounce_count,
            'through_count': self.through_count,
            'initial_ray': {'origin': (origin.x, origin.y, origin.z), 'direction': (d.x, d.y, d.z)},
        }
        return observation, info

    def step(self, action):
        b = self._batch
        a = np.asarray(action, np.float32).reshape(1, 2)
        obs, rew, term, trunc, binfo = b.step(a)
        observation = obs[0].cpu().numpy().copy()
        reward = float(rew[0].item())
        terminated, truncated = bool(term[0].item()), bool(trunc[0].item())
        row = b.info[0].cpu().numpy()
        reason = nat.REASONS[int(b.reason[0].item())]
        self.bounce_count = int(row[0])
        self.total_reward = float(row[2]) if reason != 'already_on_sun' else self.total_reward
        info = {'bounce_count': self.bounce_count, 'through_count': int(row[1])}
        if reason is not None:
            info['reason'] = reason
        if row[3] >= 0:
            info['hit_sun'] = bool(row[3])
        info['total_reward'] = float(row[2])
        self._sync_state(observation)
        return observation, reward, terminated, truncated, info

    def render(self):
        if self.render_mode == "rgb_array":
            img = np.zeros((self.image_height, self.image_width, 3), dtype=np.uint8)
            if self.current_pixel is not None:
                px, py = self.current_pixel
                c = self.accumulated_color
                img[py, px] = [min(255, max(0, c.r)), min(255, max(0, c.g)), min(255, max(0, c.b))]
            return img
        return None

    def close(self):
        self._batch.close()


class AdaptiveRewardRayTracerEnv(RayTracerEnv):
    """Drop-in for ``AdaptiveRewardRayTracerEnv`` (RL/train_raytracer_optimized.py:16-67): the RL reward plus a bonus for
    standing on a light (growing with consecutive hits), a bonus for mirrors and a short-path penalty."""
    reward_mode = "adaptive"

    def __init__(self, *args, **kwargs):
        self.light_ids = [99, 100]
        self.consecutive_light_hits = 0
        self.total_light_hits = 0
        super().__init__(*args, **kwargs)


class FBRayTracerEnv(RayTracerEnv):
    """Scalar drop-in for ``FB/ray_tracer_env.RayTracerEnv``: actions in [-1,1]^2, sun id 7, lighting reward."""
    flavour = "fb"


# ---- vectorised adapter -----------------------------------------------------------------------------
try:                                    # subclass SB3's VecEnv when stable_baselines3 is installed (it is optional)
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
except Exception:                       # pragma: no cover - depends on the image
    _VecEnvBase = object


class RayTracerVecEnv(_VecEnvBase):
    """``n_envs`` ray-tracer environments behind the Stable-Baselines3 ``VecEnv`` protocol, stepped by ONE kernel launch.

    The reference trains PPO/SAC through ``DummyVecEnv([lambda: RayTracerEnv(...)])`` (RL/train_raytracer.py:128-147),
    i.e. one Python env; this adapter gives the same agents 65,536 of them.  Protocol kept from SB3: ``reset()`` -> obs
    [n,18]; ``step(actions)`` (or ``step_async`` + ``step_wait``) -> ``(obs, rewards, dones, infos)`` where finished
    episodes are reset automatically -- the returned observation is the first one of the new episode and
    ``infos[i]['terminal_observation']`` holds the last one of the old, ``infos[i]['TimeLimit.truncated']`` its
    truncation flag.  ``as_torch=True`` returns CUDA tensors and a dict of tensors instead of numpy arrays and a list
    of dicts (no device->host copy on the rollout path)."""

    def __init__(self, spheres, n_envs, as_torch=False, **env_kwargs):
        import time
        self.env = BatchedRayTracerEnv(spheres, n_envs, **env_kwargs)
        if _VecEnvBase is not object:
            # SB3's own constructor: num_envs, the spaces, reset_infos / _seeds / _options (VecEnv.__init__)
            super().__init__(int(n_envs), self.env.observation_space, self.env.action_space)
        self.num_envs = int(n_envs)
        self.observation_space, self.action_space = self.env.observation_space, self.env.action_space
        self.as_torch = bool(as_torch)
        self.render_mode = None
        self._actions = None
        self._seed = env_kwargs.get("seed", 0)
        self.episode_returns = None
        torch = self.env.torch
        self._ep_len = torch.zeros(self.num_envs, dtype=torch.int32, device=self.env.obs.device)   # steps of the running episode
        self._t0 = time.time()

    def seed(self, seed=None):
        self._seed = 0 if seed is None else int(seed)
        return [self._seed + i for i in range(self.num_envs)]

    def reset(self):
        obs, _ = self.env.reset(seed=self._seed)
        self._seed += 1
        self._ep_len.zero_()
        if hasattr(self, "reset_infos"):
            self.reset_infos = [{} for _ in range(self.num_envs)]
        return obs.clone() if self.as_torch else obs.cpu().numpy().copy()

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        import time
        torch = self.env.torch
        # ONE launch: the step and the restart of the finished episodes (their new first observation comes back in obs,
        # the last one of the old episode in info['terminal_observation'])
        obs, rew, term, trunc, binfo = self.env.step_auto(self._actions)
        dones = term | trunc
        # SB3's DummyVecEnv sets TimeLimit.truncated = truncated and not terminated; the reference env returns
        # terminated = truncated = True at max_bounces (RL/ray_tracer_env.py:384-392), so the flag is False there and
        # SB3 does not bootstrap the value at those ends -- exactly as when it trains on the reference env
        time_limit = trunc & ~term
        self._ep_len += 1
        ep_len = self._ep_len.clone()
        self._ep_len.masked_fill_(dones, 0)
        if self.as_torch:
            infos = {"terminal_observation": binfo["terminal_observation"].clone(), "TimeLimit.truncated": time_limit,
                     "reason": binfo["reason"].clone(), "total_reward": binfo["total_reward"].clone(), "done": dones.clone(),
                     "episode_length": ep_len}
            return obs.clone(), rew.to(torch.float32).clone(), dones.clone(), infos
        d = dones.cpu().numpy()
        infos = [{} for _ in range(self.num_envs)]
        if d.any():
            lo, tn, rs, tt, ln = (binfo["terminal_observation"].cpu().numpy(), time_limit.cpu().numpy(),
                                  binfo["reason"].cpu().numpy(), binfo["total_reward"].cpu().numpy(), ep_len.cpu().numpy())
            elapsed = round(time.time() - self._t0, 6)
            for i in np.nonzero(d)[0]:
                # 'episode' carries what SB3's Monitor writes and its logger reads: return, length, elapsed seconds
                infos[i] = {"terminal_observation": lo[i].copy(), "TimeLimit.truncated": bool(tn[i]),
                            "reason": nat.REASONS[int(rs[i])], "episode": {"r": float(tt[i]), "l": int(ln[i]), "t": elapsed}}
        return obs.cpu().numpy().copy(), rew.cpu().numpy().astype(np.float32), d, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    # the rest of the VecEnv protocol, for wrappers that probe it
    def get_attr(self, attr_name, indices=None):
        return [getattr(self.env, attr_name)] * self.num_envs

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.env, attr_name, value)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        return [getattr(self.env, method_name)(*method_args, **method_kwargs)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    def get_images(self):
        return []
