"""`AckermannB200VecEnv`: the Stable-Baselines3 VecEnv face of the batched simulator.

The reference trains through ``DummyVecEnv([make_env])`` around ``Monitor(env)`` (src/rl/train.py:70-76).  This class gives SB3 the
same contract for N device-resident environments: ``reset() -> obs[N, obs_dim]``, ``step(actions[N, 2]) -> (obs, rewards, dones,
infos)`` as NumPy arrays with auto-reset, ``infos[i]["terminal_observation"]`` and ``infos[i]["TimeLimit.truncated"]`` for finished
environments and Monitor's ``infos[i]["episode"] = {"r", "l", "t"}``.  It subclasses ``stable_baselines3.common.vec_env.VecEnv`` when
SB3 is importable.  Data path: one ``ackb_step`` on the device, results copied into pinned host buffers, terminal observations
fetched for the finished rows only.
"""
from __future__ import annotations

import time
from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from .env import BatchedAckermannEnv
from .spaces import action_space, observation_space

try:  # pragma: no cover - SB3 is absent from the build image
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
    HAVE_SB3 = True
except Exception:
    HAVE_SB3 = False

    class _VecEnvBase:
        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space

        def step(self, actions):
            self.step_async(actions)
            return self.step_wait()


class AckermannB200VecEnv(_VecEnvBase):
    """N Ackermann environments behind the VecEnv API.  Keyword arguments are those of BatchedAckermannEnv / the reference env."""

    def __init__(self, num_envs: int, device="cuda:0", **env_kwargs):
        env_kwargs.setdefault("auto_reset", True)
        self.env = BatchedAckermannEnv(num_envs, device=device, **env_kwargs)
        super().__init__(num_envs, observation_space(self.env.obs_dim), action_space())
        n, d = num_envs, self.env.obs_dim
        self._act = torch.zeros((n, 2), dtype=torch.float32).pin_memory()
        self._obs = torch.zeros((n, d), dtype=torch.float32).pin_memory()
        self._rew = torch.zeros((n,), dtype=torch.float32).pin_memory()
        self._term = torch.zeros((n,), dtype=torch.uint8).pin_memory()
        self._trunc = torch.zeros((n,), dtype=torch.uint8).pin_memory()
        self._tobs_dev = torch.zeros((n, d), dtype=torch.float32, device=self.env.device)
        self._ep_ret = np.zeros(n, np.float64)
        self._ep_len = np.zeros(n, np.int64)
        self._t0 = time.time()
        self.render_mode = None

    # ---- VecEnv API ---------------------------------------------------------------------------------------------------------
    def reset(self) -> np.ndarray:
        obs = self.env.reset()
        self._ep_ret[:] = 0; self._ep_len[:] = 0
        return obs.cpu().numpy()

    def step_async(self, actions) -> None:
        a = np.clip(np.asarray(actions, dtype=np.float32).reshape(self.num_envs, 2), -1.0, 1.0)    # SB3 clips to the Box bounds
        self._act.numpy()[:] = a

    def step_wait(self):
        env = self.env
        # device step (keeps terminal observations on the device), then one D2H of the results; finished rows only for terminal obs
        obs, rew, term, trunc, info = env.step(self._act.to(env.device, non_blocking=True))
        self._obs.copy_(obs, non_blocking=True); self._rew.copy_(rew, non_blocking=True)
        self._term.copy_(term, non_blocking=True); self._trunc.copy_(trunc, non_blocking=True)
        torch.cuda.current_stream(env.device).synchronize()
        rew_np, term_np, trunc_np = self._rew.numpy().copy(), self._term.numpy().astype(bool), self._trunc.numpy().astype(bool)
        dones = term_np | trunc_np
        self._ep_ret += rew_np; self._ep_len += 1
        infos: List[dict] = [{} for _ in range(self.num_envs)]
        idx = np.nonzero(dones)[0]
        if idx.size:
            tobs = info["terminal_observation"][torch.from_numpy(idx).to(env.device)].cpu().numpy()
            for j, i in enumerate(idx):
                infos[i] = {"terminal_observation": tobs[j], "TimeLimit.truncated": bool(trunc_np[i] and not term_np[i]),
                            "episode": {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i]), "t": round(time.time() - self._t0, 6)},
                            "is_success": bool(term_np[i])}
            self._ep_ret[idx] = 0; self._ep_len[idx] = 0
        return self._obs.numpy().copy(), rew_np, dones, infos

    def close(self) -> None:
        self.env.close()

    def seed(self, seed: Optional[int] = None) -> Sequence[Optional[int]]:
        return [seed] * self.num_envs

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        n = len(self._indices(indices))
        return [getattr(self.env, attr_name)] * n

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        setattr(self.env, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> List[Any]:
        return [getattr(self.env, method_name)(*args, **kwargs)]

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False] * len(self._indices(indices))

    def get_images(self):
        raise NotImplementedError("rendering is outside the hot path")

    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        return [indices] if isinstance(indices, int) else list(indices)
