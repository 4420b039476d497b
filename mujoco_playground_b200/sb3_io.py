"""Stable-Baselines3 checkpoint interop for the PPO policy.

The reference saves ``<algo>_model_<N>_steps.zip`` through SB3's CheckpointCallback (src/rl/train.py:140-144) and
``<algo>_final.zip`` through ``model.save`` (:182-183).  Inside the zip, ``policy.pth`` is a plain torch state dict whose
keys match ``ppo.ActorCritic``; ``data`` is a JSON document with the constructor arguments and the rollout state.

load_sb3_policy  reads policy.pth of such a zip into an ActorCritic (e.g. the three checkpoints under rl_logs/ppo/).
save_sb3_policy  writes the full archive layout of SB3 2.7.0 (the version of the reference's checkpoints): ``data`` with the spaces,
                 policy class, rollout-buffer class and schedules as the same by-reference pickles SB3 writes, ``policy.pth`` in
                 ActorCriticPolicy order, a per-parameter ``policy.optimizer.pth``, ``pytorch_variables.pth`` and the version file --
                 what ``PPO.load`` opens.  SB3 is not needed to write it (class references are written through stand-in modules).
"""
from __future__ import annotations

import contextlib
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


SB3_PARAM_ORDER = ("log_std", "mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias", "mlp_extractor.policy_net.2.weight",
                   "mlp_extractor.policy_net.2.bias", "mlp_extractor.value_net.0.weight", "mlp_extractor.value_net.0.bias",
                   "mlp_extractor.value_net.2.weight", "mlp_extractor.value_net.2.bias", "action_net.weight", "action_net.bias",
                   "value_net.weight", "value_net.bias")     # ActorCriticPolicy.parameters() order = key order of the reference's policy.pth


class _Stub:
    """Instance of a third-party class that is pickled BY REFERENCE (module + qualified name) with a plain __dict__ state --
    exactly what `cloudpickle.dumps` writes for an importable class (NEWOBJ + BUILD, see the reference's own archives)."""


@contextlib.contextmanager
def sb3_stub_modules():
    """SB3 / gymnasium are not installable in the build image.  Their classes are only NAMED inside an archive (pickle stores
    `module.QualName` + the instance __dict__), so stand-in classes registered under the real module paths are enough to write --
    and, in the tests, to read back -- every pickled field of `data`.  Real modules are never shadowed: a stub is installed only
    for a module that cannot be imported, and removed afterwards."""
    import importlib
    import sys
    import types
    wanted = {"gymnasium.spaces.box": ("Box",), "stable_baselines3.common.utils": ("FloatSchedule", "ConstantSchedule"),
              "stable_baselines3.common.policies": ("ActorCriticPolicy",), "stable_baselines3.common.buffers": ("RolloutBuffer",)}
    installed = []
    try:
        for mod, names in wanted.items():
            try:
                importlib.import_module(mod)
                continue
            except Exception:
                pass
            parts = mod.split(".")
            for i in range(1, len(parts) + 1):
                name = ".".join(parts[:i])
                if name not in sys.modules:
                    sys.modules[name] = types.ModuleType(name)
                    installed.append(name)
            m = sys.modules[mod]
            for n in names:
                if not hasattr(m, n):
                    setattr(m, n, type(n, (_Stub,), {"__module__": mod, "__qualname__": n}))
        yield {mod: sys.modules[mod] for mod in wanted}
    finally:
        for name in installed:
            sys.modules.pop(name, None)


def _entry(obj, **shown) -> dict:
    """One non-JSON field of SB3's `data` document (stable_baselines3/common/save_util.py data_to_json)."""
    import base64
    import pickle
    tname = "<class 'abc.ABCMeta'>" if isinstance(obj, type) else str(type(obj))     # SB3's policy / buffer classes are ABCs
    return {":type:": tname, ":serialized:": base64.b64encode(pickle.dumps(obj, protocol=5)).decode(), **{k: str(v) for k, v in shown.items()}}


def _box(mods, low, high, shape):
    import numpy as np
    b = mods["gymnasium.spaces.box"].Box.__new__(mods["gymnasium.spaces.box"].Box)
    lo, hi = np.full(shape, low, np.float32), np.full(shape, high, np.float32)
    b.__dict__.update(dtype=np.dtype(np.float32), _shape=tuple(shape), low=lo, bounded_below=np.isfinite(lo), high=hi,
                      bounded_above=np.isfinite(hi), low_repr=str(float(low)), high_repr=str(float(high)), _np_random=None)
    return b


def _schedule(mods, val: float):
    U = mods["stable_baselines3.common.utils"]
    c = U.ConstantSchedule.__new__(U.ConstantSchedule)
    c.__dict__.update(val=float(val))
    f = U.FloatSchedule.__new__(U.FloatSchedule)
    f.__dict__.update(value_schedule=c)
    return f


def sb3_optimizer_state(policy: ActorCritic, optimizer: torch.optim.Optimizer, flat_order=None) -> dict:
    """The optimiser state as SB3 stores it: one Adam state per policy parameter in ActorCriticPolicy.parameters() order.
    The fused learner keeps ONE flat parameter (layout `flat_order` = FusedMinibatchStep.order): its state is split back."""
    named = dict(policy.named_parameters())
    groups = optimizer.state_dict()["param_groups"]
    flat_state = None
    plist = [p for g in optimizer.param_groups for p in g["params"]]
    if len(plist) == 1 and flat_order is not None and plist[0].numel() == sum(p.numel() for p in flat_order):
        flat_state = optimizer.state.get(plist[0], {})
    state = {}
    for i, name in enumerate(SB3_PARAM_ORDER):
        p = named[name]
        if flat_state is not None:
            off = 0
            for q in flat_order:
                if q is p:
                    break
                off += q.numel()
            if flat_state:
                state[i] = {"step": flat_state["step"].detach().cpu().clone().reshape(()),
                            "exp_avg": flat_state["exp_avg"][off:off + p.numel()].detach().cpu().clone().view_as(p),
                            "exp_avg_sq": flat_state["exp_avg_sq"][off:off + p.numel()].detach().cpu().clone().view_as(p)}
        else:
            st = optimizer.state.get(p, {})
            if st:
                state[i] = {k: (v.detach().cpu().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
    g = dict(groups[0])
    g["params"] = list(range(len(SB3_PARAM_ORDER)))
    g["capturable"] = False
    return {"state": state, "param_groups": [g]}


def save_sb3_policy(policy: ActorCritic, zip_path: str, optimizer: Optional[torch.optim.Optimizer] = None, num_timesteps: int = 0,
                    hyper: Optional[dict] = None, flat_order=None, last_obs=None, total_timesteps: int = 0, n_envs: int = 1) -> None:
    """Write an archive with the layout of the reference's `rl_logs/ppo/*.zip` (SB3 2.7.0): `data` (JSON; spaces, policy class,
    schedules, rollout-buffer class and buffers as base64 pickles that name gymnasium / SB3 classes), `policy.pth` (state dict in
    ActorCriticPolicy order), `policy.optimizer.pth` (per-parameter Adam state), `pytorch_variables.pth`,
    `_stable_baselines3_version`, `system_info.txt` -- what `PPO.load(path)` opens (model.save, src/rl/train.py:140-144,182-183).
    Neither SB3 nor gymnasium is needed to write it; `tests/test_capi_and_compiler.py` reads every pickled field back."""
    import collections
    import numpy as np
    h = dict(learning_rate=3e-4, n_steps=16, batch_size=64, n_epochs=10, gamma=0.99, gae_lambda=0.95, clip_range=0.2, ent_coef=0.01, vf_coef=0.5,
             max_grad_norm=0.5)
    h.update(hyper or {})
    sd_all = {k: v.detach().cpu().clone() for k, v in policy.state_dict().items()}
    sd = {k: sd_all[k] for k in SB3_PARAM_ORDER}
    obs_dim = sd["mlp_extractor.policy_net.0.weight"].shape[1]
    with sb3_stub_modules() as mods:
        lo = np.zeros((n_envs, obs_dim), np.float32) if last_obs is None else np.asarray(last_obs, np.float32).reshape(-1, obs_dim)
        data = {
            "policy_class": _entry(mods["stable_baselines3.common.policies"].ActorCriticPolicy, __module__="stable_baselines3.common.policies"),
            "verbose": 1, "policy_kwargs": {}, "num_timesteps": int(num_timesteps), "_total_timesteps": int(total_timesteps or num_timesteps),
            "_num_timesteps_at_start": 0, "seed": None, "action_noise": None, "start_time": 0, "learning_rate": float(h["learning_rate"]),
            "tensorboard_log": None, "_last_obs": _entry(lo), "_last_episode_starts": _entry(np.zeros((lo.shape[0],), bool)),
            "_last_original_obs": None, "_episode_num": 0, "use_sde": False, "sde_sample_freq": -1,
            "_current_progress_remaining": float(1.0 - num_timesteps / max(1, total_timesteps or num_timesteps)), "_stats_window_size": 100,
            "ep_info_buffer": _entry(collections.deque(maxlen=100)), "ep_success_buffer": _entry(collections.deque(maxlen=100)),
            "_n_updates": 0,
            "observation_space": _entry(_box(mods, -np.inf, np.inf, (obs_dim,)), dtype="float32", _shape=[obs_dim]),
            "action_space": _entry(_box(mods, -1.0, 1.0, (2,)), dtype="float32", _shape=[2]),
            "n_envs": int(lo.shape[0]), "n_steps": int(h["n_steps"]), "gamma": float(h["gamma"]), "gae_lambda": float(h["gae_lambda"]),
            "ent_coef": float(h["ent_coef"]), "vf_coef": float(h["vf_coef"]), "max_grad_norm": float(h["max_grad_norm"]),
            "rollout_buffer_class": _entry(mods["stable_baselines3.common.buffers"].RolloutBuffer, __module__="stable_baselines3.common.buffers"),
            "rollout_buffer_kwargs": {}, "batch_size": int(h["batch_size"]), "n_epochs": int(h["n_epochs"]),
            "clip_range": _entry(_schedule(mods, h["clip_range"]), value_schedule=f"ConstantSchedule(val={h['clip_range']})"),
            "clip_range_vf": None, "normalize_advantage": True, "target_kl": None,
            "lr_schedule": _entry(_schedule(mods, h["learning_rate"]), value_schedule=f"ConstantSchedule(val={h['learning_rate']})"),
        }
    with zipfile.ZipFile(zip_path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("data", json.dumps(data, indent=4))
        buf = io.BytesIO(); torch.save({}, buf); z.writestr("pytorch_variables.pth", buf.getvalue())
        buf = io.BytesIO(); torch.save(sd, buf); z.writestr("policy.pth", buf.getvalue())
        if optimizer is not None:
            buf = io.BytesIO(); torch.save(sb3_optimizer_state(policy, optimizer, flat_order), buf); z.writestr("policy.optimizer.pth", buf.getvalue())
        z.writestr("_stable_baselines3_version", "2.7.0")      # the layout written above is SB3 2.7.0's (version of the reference's checkpoints)
        z.writestr("system_info.txt", f"- written by mujoco_playground_b200.sb3_io (no Stable-Baselines3 at write time)\n- PyTorch: {torch.__version__}\n")


@torch.no_grad()
def evaluate_agent(env, policy, n_steps: int = 1000, deterministic: bool = True, n_episodes: int = 0) -> dict:
    """Batched counterpart of src/rl/utils.py:20-50 (evaluate_agent): run the policy on every environment of a BatchedAckermannEnv
    for n_steps (or, with n_episodes > 0, until that many episodes have finished over all ranks -- the reference's `n_episodes`
    argument) and report episode statistics from the device-side counters, summed over the ranks of a multi-GPU run.
    `policy`: an ActorCritic, or any callable obs[N, obs_dim] -> actions[N, 2] (scripted policies)."""
    from .shard import reduce_stats
    obs = env.reset()
    env.stats_reset()
    steps = 0
    while True:
        if hasattr(policy, "dist_params"):
            mean, log_std = policy.dist_params(obs)
            act = mean if deterministic else mean + torch.exp(log_std) * torch.randn_like(mean)
        else:
            act = policy(obs)
        obs, _, _, _, _ = env.step(torch.clamp(act, -1.0, 1.0))
        steps += 1
        if n_episodes > 0:
            if steps % 50 == 0 and reduce_stats(env.stats(), device=env.device)["episodes"] >= n_episodes:
                break
            if steps >= 100 * max(1, n_steps):
                break
        elif steps >= n_steps:
            break
    st = reduce_stats(env.stats(), device=env.device)
    ep = max(1, st["episodes"])
    return {"episodes": st["episodes"], "success_rate": st["successes"] / ep, "mean_reward": st["return_sum"] / ep,
            "mean_length": st["length_sum"] / ep, "collision_step_fraction": st["collisions"] / max(1, st["env_steps"])}
