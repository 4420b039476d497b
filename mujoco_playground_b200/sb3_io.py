"""Stable-Baselines3 checkpoint interop for the PPO policy.

The reference saves ``<algo>_model_<N>_steps.zip`` through SB3's CheckpointCallback (src/rl/train.py:140-144) and
``<algo>_final.zip`` through ``model.save`` (:182-183).  Inside the zip, ``policy.pth`` is a plain torch state dict whose
keys match ``ppo.ActorCritic``; ``data`` is a JSON document with the constructor arguments and the rollout state.

load_sb3_policy  reads policy.pth of such a zip into an ActorCritic (e.g. the three checkpoints under rl_logs/ppo/).
save_sb3_policy  writes policy.pth + policy.optimizer.pth + a ``data`` JSON with the hyper-parameters this trainer knows.
                 It is a weights container in SB3's layout, not a full ``PPO.load``-able archive (SB3 also pickles
                 spaces, schedules and the policy class into ``data``; those need SB3 itself).
"""
from __future__ import annotations

import io
import json
import zipfile
from typing import Optional, Tuple

import torch

from .ppo import ActorCritic


def load_sb3_policy(zip_path: str, device="cpu") -> Tuple[ActorCritic, dict]:
    """Returns (policy, data) where data is the parsed ``data`` JSON of the archive (may be empty)."""
    with zipfile.ZipFile(zip_path) as z:
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location=device, weights_only=True)
        data = json.loads(z.read("data")) if "data" in z.namelist() else {}
    obs_dim = sd["mlp_extractor.policy_net.0.weight"].shape[1]
    hidden = sd["mlp_extractor.policy_net.0.weight"].shape[0]
    act_dim = sd["action_net.weight"].shape[0]
    pol = ActorCritic(obs_dim, act_dim, hidden).to(device)
    pol.load_state_dict(sd)
    return pol, data


def save_sb3_policy(policy: ActorCritic, zip_path: str, optimizer: Optional[torch.optim.Optimizer] = None, num_timesteps: int = 0,
                    hyper: Optional[dict] = None) -> None:
    with zipfile.ZipFile(zip_path, "w", zipfile.ZIP_DEFLATED) as z:
        buf = io.BytesIO()
        torch.save({k: v.detach().cpu() for k, v in policy.state_dict().items()}, buf)
        z.writestr("policy.pth", buf.getvalue())
        if optimizer is not None:
            buf = io.BytesIO()
            torch.save(optimizer.state_dict(), buf)
            z.writestr("policy.optimizer.pth", buf.getvalue())
        data = {"policy_class": "ActorCriticPolicy (MlpPolicy)", "num_timesteps": int(num_timesteps), "n_envs": None}
        data.update(hyper or {})
        z.writestr("data", json.dumps(data))
        z.writestr("_stable_baselines3_version", "layout-compatible weights container (written by mujoco_playground_b200)")


@torch.no_grad()
def evaluate_agent(env, policy: ActorCritic, n_steps: int = 1000, deterministic: bool = True) -> dict:
    """Batched counterpart of src/rl/utils.py:20-50 (evaluate_agent): run the policy for n_steps on every environment of
    a BatchedAckermannEnv and report episode statistics from the device-side counters."""
    obs = env.reset()
    env.stats_reset()
    for _ in range(n_steps):
        mean, log_std = policy.dist_params(obs)
        act = mean if deterministic else mean + torch.exp(log_std) * torch.randn_like(mean)
        obs, _, _, _, _ = env.step(torch.clamp(act, -1.0, 1.0))
    st = env.stats()
    ep = max(1, st["episodes"])
    return {"episodes": st["episodes"], "success_rate": st["successes"] / ep, "mean_reward": st["return_sum"] / ep,
            "mean_length": st["length_sum"] / ep, "collision_step_fraction": st["collisions"] / max(1, st["env_steps"])}
