"""Batched, device-resident replacement for the reference's Ackermann RL environment.

``BatchedAckermannEnv`` is the vectorised face (what SB3's DummyVecEnv + Monitor give the reference,
src/rl/train.py:70-76); ``AckermannRobotEnv`` is a num_envs=1 adapter with the exact signature of the
reference class (src/rl/envs/ackermann_env.py:32-325).  Both call the C ABI of libackb.so; PyTorch is
used only to own device memory and streams.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib
from .compiler.constants import build_consts
from .models import load_model, model_kind
from .spaces import GymEnv, action_space, observation_space

_DTYPES = {"float32": 0, "f32": 0, torch.float32: 0, "float64": 1, "f64": 1, torch.float64: 1}


class BatchedAckermannEnv:
    """num_envs independent Ackermann robots stepped by one fused CUDA kernel per ``step``.

    Keyword names and defaults of the reference constructor are kept (ackermann_env.py:51-60).
    ``frame_skip`` is new (the reference always does one mj_step per step) and defaults to 1.
    """

    def __init__(self, num_envs: int, device="cuda:0", frame_skip: int = 1, dtype="float32", seed: int = 0, model: str = "v2",
                 auto_reset: bool = True, lanes_per_env: int = 0, lidar_index_map: str = "reference",
                 max_episode_steps: int = 1000, goal_distance_threshold: float = 0.5, collision_threshold: float = 0.15,
                 max_linear_velocity: float = 1.0, max_angular_velocity: float = 1.0, render_mode=None, map_spawner=None,
                 solver_tolerance: Optional[float] = None, spawn_yaw_range: float = 0.0, spawn_xy_jitter: float = 0.0,
                 model_table: Optional[dict] = None, settle_steps: Optional[int] = None, env_id_base: int = 0):
        if render_mode is not None:
            raise NotImplementedError("rendering is outside the hot path (SURVEY.md section 2)")
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedAckermannEnv needs a CUDA device: there is no CPU fallback")
        self.L = _lib.load()
        self.device = torch.device(device)
        self.num_envs = int(num_envs)
        self.frame_skip = int(frame_skip)
        self.auto_reset = bool(auto_reset)
        self.model_name = model
        self.table = model_table if model_table is not None else load_model(model)
        self.consts = build_consts(self.table, model_kind=model_kind(model), max_episode_steps=max_episode_steps,
                                   goal_distance_threshold=goal_distance_threshold, collision_threshold=collision_threshold,
                                   max_linear_velocity=max_linear_velocity, max_angular_velocity=max_angular_velocity,
                                   lidar_index_map=lidar_index_map, tolerance=solver_tolerance,
                                   spawn_yaw_range=spawn_yaw_range, spawn_xy_jitter=spawn_xy_jitter)
        if settle_steps is not None:     # maze scenes: override the number of settle steps after a reset (tests)
            from .compiler.constants import consts_layout
            self.consts[consts_layout()["settle_steps"][0]] = float(settle_steps)
        assert len(self.consts) == self.L.ackb_consts_len(), "constants layout mismatch between Python and libackb.so"
        self.h = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(self.L.ackb_create(self.consts.ctypes.data_as(ctypes.c_void_p), len(self.consts), self.num_envs, dev_index,
                                      _DTYPES[dtype], int(seed), int(lanes_per_env), ctypes.byref(self.h)))
        # global id of environment 0: random streams are keyed by (seed, GLOBAL env id), so a batch sharded over several
        # handles / ranks with the same seed equals the single-handle batch
        self.env_id_base = int(env_id_base)
        if self.env_id_base:
            _lib.check(self.L.ackb_set_env_id_base(self.h, self.env_id_base), self.h)
        self.obs_dim = self.L.ackb_obs_dim(self.h)
        # spaces of ONE environment, as the reference declares them (ackermann_env.py:95-108)
        self.observation_space, self.action_space = observation_space(self.obs_dim), action_space()
        self.single_observation_space, self.single_action_space = self.observation_space, self.action_space
        self.max_episode_steps = int(max_episode_steps)
        n, d = self.num_envs, self.device
        self.obs = torch.empty((n, self.obs_dim), dtype=torch.float32, device=d)
        self.reward = torch.empty((n,), dtype=torch.float32, device=d)
        self.terminated = torch.empty((n,), dtype=torch.uint8, device=d)
        self.truncated = torch.empty((n,), dtype=torch.uint8, device=d)
        self.terminal_obs = torch.zeros((n, self.obs_dim), dtype=torch.float32, device=d)
        self.ncon = torch.zeros((n,), dtype=torch.int32, device=d)
        self._pitch = self.obs_dim        # row pitch (floats) the handle currently writes observations with

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.ackb_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_obs_out(self, obs_out: torch.Tensor) -> int:
        """Validates a caller-provided observation array [N, pitch >= obs_dim] and returns its row pitch in floats."""
        if (obs_out.device != self.device or obs_out.dtype != torch.float32 or obs_out.dim() != 2 or obs_out.shape[0] != self.num_envs
                or obs_out.shape[1] < self.obs_dim or not obs_out.is_contiguous()):
            raise ValueError(f"obs_out must be a contiguous float32 [{self.num_envs}, >= {self.obs_dim}] tensor on {self.device}")
        return int(obs_out.shape[1])

    def _set_pitch(self, pitch: int) -> None:
        """Row pitch of the observation arrays handed to the next call (ackb_set_obs_pitch); terminal observations follow it."""
        if pitch != self._pitch:
            _lib.check(self.L.ackb_set_obs_pitch(self.h, int(pitch)), self.h)
            self._pitch = int(pitch)
        if self.terminal_obs.shape[1] != pitch:
            self.terminal_obs = torch.zeros((self.num_envs, pitch), dtype=torch.float32, device=self.device)

    def reset(self, mask: Optional[torch.Tensor] = None, obs_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reset all environments (or those where mask != 0); returns the observation tensor [N, obs_dim] (or obs_out, see step)."""
        obs = self.obs if obs_out is None else obs_out
        self._set_pitch(self.obs_dim if obs_out is None else self._check_obs_out(obs_out))
        mp = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mp = ctypes.c_void_p(mask.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(self.L.ackb_reset(self.h, mp, ctypes.c_void_p(obs.data_ptr()), self._stream()), self.h)
        return obs

    def step(self, actions: Optional[torch.Tensor], obs_out: Optional[torch.Tensor] = None):
        """actions: [N, 2] float32 CUDA tensor in [-1, 1] (clipped like the reference), or None for synthetic
        device-generated U(-1,1) actions.  Returns (obs, reward, terminated, truncated, info) as device tensors.
        obs_out: optional contiguous [N, pitch >= obs_dim] float32 tensor on the env's device that receives the observations instead
        of the environment's own buffer (e.g. the next slot of a rollout buffer: saves the copy).  With pitch > obs_dim (80: rows on
        16-byte boundaries for the learner) the columns beyond obs_dim are left untouched and info["terminal_observation"] has the
        same pitch."""
        obs = self.obs
        if obs_out is not None:
            self._set_pitch(self._check_obs_out(obs_out))
            obs = obs_out
        else:
            self._set_pitch(self.obs_dim)
        ap = None
        if actions is not None:
            if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous():
                actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(actions.shape) != (self.num_envs, 2):
                raise ValueError(f"actions must have shape ({self.num_envs}, 2)")
            ap = ctypes.c_void_p(actions.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(self.L.ackb_step(self.h, ap, self.frame_skip, int(self.auto_reset), ctypes.c_void_p(obs.data_ptr()),
                                        ctypes.c_void_p(self.reward.data_ptr()), ctypes.c_void_p(self.terminated.data_ptr()),
                                        ctypes.c_void_p(self.truncated.data_ptr()), ctypes.c_void_p(self.terminal_obs.data_ptr()),
                                        ctypes.c_void_p(self.ncon.data_ptr()), self._stream()), self.h)
        info = {"terminal_observation": self.terminal_obs, "ncon": self.ncon}
        return obs, self.reward, self.terminated, self.truncated, info

    def step_host(self, actions: torch.Tensor, obs: torch.Tensor, reward: torch.Tensor, terminated: torch.Tensor, truncated: torch.Tensor):
        """End-to-end step with HOST (pinned) tensors: H2D actions, step, D2H results; synchronous."""
        _lib.check(self.L.ackb_step_host(self.h, ctypes.c_void_p(actions.data_ptr()), self.frame_skip, int(self.auto_reset),
                                         ctypes.c_void_p(obs.data_ptr()), ctypes.c_void_p(reward.data_ptr()),
                                         ctypes.c_void_p(terminated.data_ptr()), ctypes.c_void_p(truncated.data_ptr())), self.h)

    # ---- state access for tests ----------------------------------------------------------------
    def get_state(self):
        n = self.num_envs
        qpos, qvel, warm = np.zeros((n, 13)), np.zeros((n, 12)), np.zeros((n, 12))
        _lib.check(self.L.ackb_get_state(self.h, qpos.ctypes.data_as(ctypes.c_void_p), qvel.ctypes.data_as(ctypes.c_void_p),
                                         warm.ctypes.data_as(ctypes.c_void_p)), self.h)
        return qpos, qvel, warm

    def set_state(self, qpos=None, qvel=None, warm=None):
        def p(a, w):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            assert a.shape == (self.num_envs, w)
            keep.append(a)
            return a.ctypes.data_as(ctypes.c_void_p)
        keep = []
        _lib.check(self.L.ackb_set_state(self.h, p(qpos, 13), p(qvel, 12), p(warm, 12)), self.h)

    def get_episode(self):
        n = self.num_envs
        goal, ref, sc = np.zeros((n, 2)), np.zeros((n, 2)), np.zeros(n, np.int32)
        _lib.check(self.L.ackb_get_episode(self.h, goal.ctypes.data_as(ctypes.c_void_p), ref.ctypes.data_as(ctypes.c_void_p),
                                           sc.ctypes.data_as(ctypes.c_void_p)), self.h)
        return goal, ref, sc

    def set_episode(self, goal=None, ref=None, step_count=None):
        keep = []

        def p(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(self.L.ackb_set_episode(self.h, p(goal, np.float64), p(ref, np.float64), p(step_count, np.int32)), self.h)

    def stats(self) -> dict:
        s = _lib.AckbStats()
        _lib.check(self.L.ackb_stats(self.h, ctypes.byref(s)), self.h)
        return {k: getattr(s, k) for k, _ in s._fields_}

    def stats_reset(self):
        _lib.check(self.L.ackb_stats_reset(self.h), self.h)

    @property
    def launch_count(self) -> int:
        return int(self.L.ackb_launch_count(self.h))


class AckermannRobotEnv(GymEnv):
    """Single-environment adapter with the reference's gym.Env signature (ackermann_env.py:51-60,143,187).

    ``reset(seed, options) -> (obs[79] float32 ndarray, info)``;
    ``step(action) -> (obs, reward, terminated, truncated, info)`` with the reference's info keys
    (goal_distance, collision, min_lidar, step, linear_velocity, angular_velocity).  No auto-reset.
    """

    metadata = {"render_modes": [], "render_fps": 50}

    def __init__(self, map_spawner=None, max_episode_steps=1000, goal_distance_threshold=0.5, collision_threshold=0.15,
                 max_linear_velocity=1.0, max_angular_velocity=1.0, render_mode=None, device="cuda:0", dtype="float64",
                 frame_skip=1, seed=0, model="v2"):
        self.max_episode_steps = max_episode_steps
        self.goal_distance_threshold = goal_distance_threshold
        self.collision_threshold = collision_threshold
        self.max_linear_velocity = max_linear_velocity
        self.max_angular_velocity = max_angular_velocity
        self._kw = dict(device=device, dtype=dtype, frame_skip=frame_skip, max_episode_steps=max_episode_steps,
                        goal_distance_threshold=goal_distance_threshold, collision_threshold=collision_threshold,
                        max_linear_velocity=max_linear_velocity, max_angular_velocity=max_angular_velocity,
                        render_mode=render_mode, auto_reset=False, model=model)
        self._map_name = "simple_floor" if model == "v2" else model
        self._seed = seed
        self._env = BatchedAckermannEnv(1, seed=seed, **self._kw)
        self.observation_shape, self.action_shape = (self._env.obs_dim,), (2,)
        self.observation_space, self.action_space = self._env.observation_space, self._env.action_space   # ackermann_env.py:95-108
        self.render_mode = render_mode
        self.step_count = 0
        self.goal_position = None
        self.robot_start_position = np.zeros(2)

    def reset(self, seed=None, options=None):
        if seed is not None:     # gym contract: an explicit seed restarts the random streams, also when it is the same seed again
            self._seed = seed
            self._env.close()
            self._env = BatchedAckermannEnv(1, seed=seed, **self._kw)
        obs = self._env.reset().cpu().numpy()[0].copy()
        goal, ref, _ = self._env.get_episode()
        self.step_count = 0
        self.goal_position = goal[0].copy()
        if self._map_name != "simple_floor":
            self.robot_start_position = ref[0].copy()
        info = {"map_name": self._map_name, "goal_position": self.goal_position.tolist(),
                "start_position": self.robot_start_position.tolist()}
        return obs, info

    def step(self, action):
        a = np.clip(np.asarray(action, dtype=np.float32), -1.0, 1.0)
        act = torch.from_numpy(a.reshape(1, 2)).to(self._env.device)
        obs, rew, term, trunc, _ = self._env.step(act)
        obs = obs.cpu().numpy()[0].copy()
        self.step_count += 1
        nb = self._env.obs_dim - 7
        lidar_min = float(obs[:nb].min())
        info = {"goal_distance": float(obs[nb + 5]), "collision": bool(lidar_min < self.collision_threshold), "min_lidar": lidar_min,
                "step": self.step_count, "linear_velocity": np.float32(a[0] * self.max_linear_velocity),
                "angular_velocity": np.float32(a[1] * self.max_angular_velocity)}
        return obs, float(rew.item()), bool(term.item()), bool(trunc.item()), info

    def render(self):
        return None

    def close(self):
        self._env.close()


class AckermannGymnasiumMazeEnv(AckermannRobotEnv):
    """Single-environment adapter with the signature of the reference's maze environment
    (src/rl/envs/ackermann_gymnasium_maze_env.py:50-61): ``maze_env_id`` selects the PointMaze layout (compiler/maze.py)."""

    def __init__(self, maze_env_id="PointMaze_UMaze-v3", max_episode_steps=1000, goal_distance_threshold=0.5, collision_threshold=0.15,
                 max_linear_velocity=1.0, max_angular_velocity=1.0, render_mode=None, **kw):
        from .compiler.maze import MAZE_ENV_IDS
        if maze_env_id not in MAZE_ENV_IDS:
            raise ValueError(f"Failed to load maze environment '{maze_env_id}': unknown id (known: {sorted(MAZE_ENV_IDS)})")
        self.maze_env_id = maze_env_id
        super().__init__(max_episode_steps=max_episode_steps, goal_distance_threshold=goal_distance_threshold,
                         collision_threshold=collision_threshold, max_linear_velocity=max_linear_velocity,
                         max_angular_velocity=max_angular_velocity, render_mode=render_mode, model="maze:" + MAZE_ENV_IDS[maze_env_id], **kw)

    def reset(self, seed=None, options=None):
        obs, info = super().reset(seed=seed, options=options)
        return obs, {"maze_type": self.maze_env_id, "goal_position": info["goal_position"], "start_position": info["start_position"]}
