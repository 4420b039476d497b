"""PPO driver for the batched Ackermann environment: the caller side of the hot path.

Mirrors what the reference gets from Stable-Baselines3 in ``train_with_stable_baselines3`` (src/rl/train.py:44-186):
PPO('MlpPolicy', lr=3e-4, n_epochs=10, gamma=0.99, gae_lambda=0.95, clip_range=0.2, ent_coef=0.01) with SB3's defaults
vf_coef=0.5, max_grad_norm=0.5, Adam(eps=1e-5), advantage normalisation, separate tanh MLPs 79->64->64 for policy and
value, state-independent log_std, action clipping to the Box bounds, and value bootstrapping on time-limit truncation.
Deviation: with 10^4..10^5 environments per GPU, n_steps=2048 / batch_size=64 are replaced by a short
rollout (n_steps, default 16) and large minibatches.

Multi-GPU: one process per GPU, each owns an independent environment shard; the ONLY collectives are one flattened
gradient all-reduce per optimiser step (18 757 fp32 = 75 KB) and one statistics all-reduce per rollout.
"""
from __future__ import annotations

import math
import os
import time
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .shard import reduce_stats


class ActorCritic(nn.Module):
    """SB3 MlpPolicy layout (parameter names match policy.pth inside rl_logs/ppo/*.zip): 18 757 parameters for obs 79."""

    def __init__(self, obs_dim: int = 79, act_dim: int = 2, hidden: int = 64, log_std_init: float = 0.0):
        super().__init__()
        self.mlp_extractor = nn.ModuleDict({
            "policy_net": nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh()),
            "value_net": nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh()),
        })
        self.action_net = nn.Linear(hidden, act_dim)
        self.value_net = nn.Linear(hidden, 1)
        self.log_std = nn.Parameter(torch.full((act_dim,), float(log_std_init)))
        for seq, gain in ((self.mlp_extractor["policy_net"], math.sqrt(2)), (self.mlp_extractor["value_net"], math.sqrt(2))):
            for m in seq:
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight, gain)
                    nn.init.zeros_(m.bias)
        nn.init.orthogonal_(self.action_net.weight, 0.01)
        nn.init.zeros_(self.action_net.bias)
        nn.init.orthogonal_(self.value_net.weight, 1.0)
        nn.init.zeros_(self.value_net.bias)

    def value(self, obs):
        return self.value_net(self.mlp_extractor["value_net"](obs)).squeeze(-1)

    def dist_params(self, obs):
        return self.action_net(self.mlp_extractor["policy_net"](obs)), self.log_std

    @staticmethod
    def log_prob(mean, log_std, act):
        var = torch.exp(2 * log_std)
        return (-((act - mean) ** 2) / (2 * var) - log_std - 0.5 * math.log(2 * math.pi)).sum(-1)

    def act(self, obs):
        mean, log_std = self.dist_params(obs)
        act = mean + torch.exp(log_std) * torch.randn_like(mean)
        return act, self.log_prob(mean, log_std, act), self.value(obs)

    def evaluate(self, obs, act):
        mean, log_std = self.dist_params(obs)
        entropy = (0.5 + 0.5 * math.log(2 * math.pi) + log_std).sum(-1)
        return self.log_prob(mean, log_std, act), entropy.expand(obs.shape[0]), self.value(obs)


def sanitize_obs(obs: torch.Tensor) -> torch.Tensor:
    """Observations are finite by construction (lidar -1..12, metres, radians); kept as a hook for input scaling."""
    return obs


@dataclass
class PPOConfig:
    n_steps: int = 16
    n_epochs: int = 10
    minibatches: int = 4
    learning_rate: float = 3e-4
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_range: float = 0.2
    ent_coef: float = 0.01
    vf_coef: float = 0.5
    max_grad_norm: float = 0.5
    adam_eps: float = 1e-5


def allreduce_gradients_(params, world: int, group=None) -> int:
    """Average gradients over ranks with ONE flattened all-reduce; returns the number of bytes reduced."""
    grads = [p.grad for p in params if p.grad is not None]
    if world <= 1 or not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * flat.element_size()


def compute_gae(rewards, values, dones, last_value, gamma: float, lam: float):
    """rewards/values/dones: [T, N]; dones[t] = episode ended after step t.  Returns (advantages, returns)."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(last_value)
    next_value = last_value
    for t in range(T - 1, -1, -1):
        nonterminal = 1.0 - dones[t]
        delta = rewards[t] + gamma * next_value * nonterminal - values[t]
        last = delta + gamma * lam * nonterminal * last
        adv[t] = last
        next_value = values[t]
    return adv, adv + values


def ppo_update(policy: ActorCritic, opt: torch.optim.Optimizer, batch: Dict[str, torch.Tensor], cfg: PPOConfig, world: int = 1,
               generator: Optional[torch.Generator] = None, graphed: Optional["GraphedMinibatchStep"] = None,
               perm_seed: int = 0) -> Dict[str, float]:
    """n_epochs passes over the flattened rollout in `minibatches` shuffled minibatches, gradient all-reduce per step."""
    n = batch["obs"].shape[0]
    mb = max(1, n // cfg.minibatches)
    stats = dict(pg_loss=0.0, v_loss=0.0, entropy=0.0, approx_kl=0.0, clip_frac=0.0, steps=0, allreduce_bytes=0)
    acc = torch.zeros(5, device=batch["obs"].device)      # diagnostics accumulate on the device: one host read per update
    in_kernel = isinstance(graphed, FusedMinibatchStep) and graphed.accumulate_diag      # ... inside the gradient kernel itself
    if in_kernel:
        graphed.diag.zero_()
    params = [p for p in policy.parameters() if p.requires_grad]
    # native per-epoch permutation (no sort) for the fused learner in index mode; an explicit generator keeps torch.randperm
    native_perm = (isinstance(graphed, FusedMinibatchStep) and graphed.index_mode and generator is None
                   and os.environ.get("ACKB_PPO_NATIVE_PERM", "1") != "0")
    for _ in range(cfg.n_epochs):
        if native_perm:
            graphed.epoch_counter = getattr(graphed, "epoch_counter", 0) + 1
            shuffled, perm = graphed.new_epoch(batch, perm_seed, graphed.epoch_counter), None
        else:
            perm = torch.randperm(n, device=batch["obs"].device, generator=generator)
            shuffled = graphed.shuffle_epoch(batch, perm) if isinstance(graphed, FusedMinibatchStep) else None
        if shuffled is not None and in_kernel:      # a whole epoch of optimiser steps: one graph launch when it was captured
            slots = [(i * mb, min(mb, n - i * mb)) for i in range(cfg.minibatches)]
            stats["allreduce_bytes"] += graphed.run_epoch(shuffled, slots, world)
            stats["steps"] += len(slots)
            continue
        for i in range(cfg.minibatches):
            if shuffled is not None:
                stats["allreduce_bytes"] += graphed.run(shuffled, (i * mb, min(mb, n - i * mb)), world)
                if not in_kernel:
                    acc += graphed.diag
                stats["steps"] += 1
                continue
            idx = perm[i * mb:(i + 1) * mb]
            if graphed is not None and graphed.mb in (-1, idx.numel()):
                stats["allreduce_bytes"] += graphed.run(batch, idx, world)
                if not in_kernel:
                    acc += graphed.diag
                stats["steps"] += 1
                continue
            obs, act, old_logp, adv, ret = (batch[k][idx] for k in ("obs", "act", "logp", "adv", "ret"))
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)
            logp, ent, val = policy.evaluate(obs, act)
            ratio = torch.exp(logp - old_logp)
            pg = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
            vl = torch.nn.functional.mse_loss(val, ret)
            loss = pg + cfg.vf_coef * vl - cfg.ent_coef * ent.mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            stats["allreduce_bytes"] += allreduce_gradients_(params, world)
            torch.nn.utils.clip_grad_norm_(params, cfg.max_grad_norm)
            opt.step()
            with torch.no_grad():
                acc += torch.stack([pg.detach(), vl.detach(), ent.mean(), ((ratio - 1) - (logp - old_logp)).mean(),
                                    ((ratio - 1).abs() > cfg.clip_range).float().mean()])
                stats["steps"] += 1
    if in_kernel:
        acc = graphed.diag
    for k, v in zip(("pg_loss", "v_loss", "entropy", "approx_kl", "clip_frac"), (acc / max(1, stats["steps"])).tolist()):
        stats[k] = v
    return stats


class GraphedMinibatchStep:
    """One PPO optimiser step on a fixed-size minibatch, captured in two CUDA graphs (the step is launch bound in eager mode:
    ~200 small kernels for an 18 757-parameter network):
      graph 1: advantage normalisation, forward, loss, backward (gradients land in static .grad tensors) + diagnostics
      [eager : the flattened gradient all-reduce when world > 1]
      graph 2: gradient clipping + Adam step (the optimiser must be built with capturable=True).
    Same arithmetic as the eager path of ppo_update (tests/test_ppo.py compares them)."""

    def __init__(self, policy: ActorCritic, opt: torch.optim.Optimizer, cfg: PPOConfig, mb: int, obs_dim: int, device):
        self.policy, self.opt, self.cfg, self.mb = policy, opt, cfg, mb
        f = dict(device=device, dtype=torch.float32)
        self.obs, self.act = torch.zeros((mb, obs_dim), **f), torch.zeros((mb, 2), **f)
        self.logp, self.adv, self.ret = torch.zeros(mb, **f), torch.randn(mb, **f), torch.zeros(mb, **f)
        self.diag = torch.zeros(5, **f)
        self.params = [p for p in policy.parameters() if p.requires_grad]
        # warm-up on a side stream (allocator / autograd / optimiser state), then restore weights and optimiser state in place
        saved_p = [p.detach().clone() for p in self.params]
        saved_s = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in opt.state.get(p, {}).items()} for p in self.params}
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._fwd_bwd()
                self._clip_step()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.g1, self.g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g1):
            self._fwd_bwd()
        with torch.cuda.graph(self.g2):
            self._clip_step()
        with torch.no_grad():
            for p, sp in zip(self.params, saved_p):
                p.copy_(sp)
                st = opt.state.get(p, {})
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if k in saved_s[id(p)]:
                            v.copy_(saved_s[id(p)][k])
                        else:
                            v.zero_()
        torch.cuda.synchronize(device)

    def _fwd_bwd(self):
        cfg = self.cfg
        adv = (self.adv - self.adv.mean()) / (self.adv.std() + 1e-8)
        logp, ent, val = self.policy.evaluate(self.obs, self.act)
        ratio = torch.exp(logp - self.logp)
        pg = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
        vl = torch.nn.functional.mse_loss(val, self.ret)
        loss = pg + cfg.vf_coef * vl - cfg.ent_coef * ent.mean()
        self.opt.zero_grad(set_to_none=False)
        loss.backward()
        with torch.no_grad():
            self.diag.copy_(torch.stack([pg.detach(), vl.detach(), ent.mean(), ((ratio - 1) - (logp - self.logp)).mean(),
                                         ((ratio - 1).abs() > cfg.clip_range).float().mean()]))

    def _clip_step(self):
        torch.nn.utils.clip_grad_norm_(self.params, self.cfg.max_grad_norm)
        self.opt.step()

    def run(self, batch: Dict[str, torch.Tensor], idx: torch.Tensor, world: int) -> int:
        torch.index_select(batch["obs"], 0, idx, out=self.obs)
        torch.index_select(batch["act"], 0, idx, out=self.act)
        torch.index_select(batch["logp"], 0, idx, out=self.logp)
        torch.index_select(batch["adv"], 0, idx, out=self.adv)
        torch.index_select(batch["ret"], 0, idx, out=self.ret)
        self.g1.replay()
        nbytes = allreduce_gradients_(self.params, world)
        self.g2.replay()
        return nbytes


_ROLLOUT_KEYS = ("obs", "act", "logp", "adv", "ret")


class FusedMinibatchStep:
    """One PPO optimiser step with the hand-written fused gradient kernel (csrc/ackb_ppo.cu, include/ackb_ppo.h): forward,
    PPO loss, backward and weight-gradient accumulation of both 64-wide MLPs in one launch with weights and activations in shared
    memory.  Parameters and gradients live in two flat device buffers in the kernel's layout (the nn.Parameters become views
    into them), so there is no packing, the gradient all-reduce is one call on the flat buffer, and clipping + Adam run on the
    flat vector: one kernel (ackb_ppo_clip_adam) on torch's own optimiser-state tensors, or torch ops if the optimiser is not a
    capturable plain Adam.  Same arithmetic as the eager loop of ppo_update up to fp32 summation order (tests/test_ppo.py)."""

    MODES = {"default": -1, "fp32": 0, "tf32": 1, "tcgen05": 2}
    DIAG_ACCUMULATE = 0x100      # ACKB_PPO_DIAG_ACCUMULATE (include/ackb_ppo.h): flag OR-ed onto the per-call mode

    def __init__(self, policy: ActorCritic, opt: torch.optim.Optimizer, cfg: PPOConfig, obs_dim: int, device, mode: str = "tcgen05"):
        import ctypes
        from . import _lib
        self.L, self.ct = _lib.load(), ctypes
        # arithmetic of the gradient kernel, chosen per learner and passed with every call (include/ackb_ppo.h: ACKB_PPO_MODE_*)
        # "tcgen05" (default): TF32 on the Blackwell tensor-core path (TMEM accumulators); "tf32": mma.sync fragments; "fp32": CUDA cores
        self.mode = self.MODES[os.environ.get("ACKB_PPO_MODE", mode)]
        if self.mode == 2 and obs_dim >= 80:
            self.mode = 1      # the tcgen05 kernel carries the layer-1 bias gradient in column 79 of the observation tile
        self.policy, self.opt, self.cfg, self.obs_dim, self.device = policy, opt, cfg, obs_dim, device
        pe, ve = policy.mlp_extractor["policy_net"], policy.mlp_extractor["value_net"]
        self.order = [pe[0].weight, pe[0].bias, pe[2].weight, pe[2].bias, ve[0].weight, ve[0].bias, ve[2].weight, ve[2].bias,
                      policy.action_net.weight, policy.action_net.bias, policy.value_net.weight, policy.value_net.bias, policy.log_std]
        total = sum(p.numel() for p in self.order)
        assert total == self.L.ackb_ppo_num_params(obs_dim), "ActorCritic does not match the kernel's parameter layout"
        self.flat_p = torch.empty(total, device=device, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=device, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p in self.order:
                n = p.numel()
                self.flat_p[off:off + n].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + n].view_as(p)       # parameters become views of the flat buffer
                p.grad = self.flat_g[off:off + n].view_as(p)       # ... and so do their gradients
                off += n
        self.params = list(self.order)
        # the optimiser works on ONE flat parameter (Adam and the global-norm clip are elementwise / global, so this is the same
        # arithmetic as 13 per-tensor updates with a fraction of the kernels); only for a fresh optimiser (no state to remap)
        if len(opt.param_groups) == 1 and not opt.state:
            self.flat_param = torch.nn.Parameter(self.flat_p, requires_grad=True)
            self.flat_param.grad = self.flat_g
            opt.param_groups[0]["params"] = [self.flat_param]
            self.params = [self.flat_param]
        self.diag = torch.zeros(5, device=device, dtype=torch.float32)
        # True: the gradient call ADDS the minibatch's diagnostics to `diag` (ACKB_PPO_DIAG_ACCUMULATE) instead of overwriting it;
        # ppo_update then zeroes it once per update and launches nothing per step.  Set before capture() (the flag is baked into
        # the captured launches).
        self.accumulate_diag = False
        self.adv_stats = torch.zeros(2, device=device, dtype=torch.float32)
        self.adv_ws = torch.zeros(3, device=device, dtype=torch.float64)     # this learner's own accumulators (ackb_ppo_adv_stats_ws)
        # optimiser step as one kernel (ackb_ppo_clip_adam) on torch's own Adam state tensors, so that opt.state_dict() stays the
        # checkpoint format; needs the flat parameter and a device-side step counter (capturable=True), plain Adam only
        g = opt.param_groups[0]
        self.native_opt = (getattr(self, "flat_param", None) is not None and type(opt) is torch.optim.Adam and bool(g.get("capturable", False))
                           and not g.get("amsgrad") and not g.get("maximize") and g.get("weight_decay", 0) == 0
                           and not torch.is_tensor(g["lr"]) and os.environ.get("ACKB_PPO_NATIVE_OPT", "1") != "0")
        if self.native_opt:
            st = opt.state[self.flat_param]
            if not st:
                st["step"] = torch.zeros((), dtype=torch.float32, device=device)
                st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(self.flat_p), torch.zeros_like(self.flat_p)
        self.mb = -1      # any minibatch size
        self.index_mode = os.environ.get("ACKB_PPO_SHUFFLE", "index") != "copy"
        self.peer = None  # peer-memory gradient all-reduce inside the optimiser-step kernel (enable_peer_allreduce)
        self._mbc = 0     # optimiser steps taken: its parity selects the gradient buffer of the peer path

    def enable_peer_allreduce(self, world: int, rank: int) -> bool:
        """Data-parallel learners on one node: move the 75 KB gradient all-reduce INTO the optimiser-step kernel
        (ackb_ppo_clip_adam_allreduce: flags + peer loads over NVLink instead of an NCCL call per optimiser step).  Collective: every
        rank must call it at the same point.  The gradient kernels then write into one of two buffers of a symmetric-memory
        allocation (torch.distributed._symmetric_memory), alternating per optimiser step.  Returns False (and keeps the NCCL path)
        if the optimiser step is not the native kernel, symmetric memory is unavailable, or ACKB_PPO_PEER_ALLREDUCE=0."""
        if world <= 1 or not self.native_opt or os.environ.get("ACKB_PPO_PEER_ALLREDUCE", "1") == "0":
            return False
        dev, n = self.device, self.flat_p.numel()
        stride = (n + 31) // 32 * 32
        ok, st = 1, None
        try:
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty(2 * stride + 64, dtype=torch.float32, device=dev)       # two gradient buffers + 64 flag words
            hdl = symm.rendezvous(buf, dist.group.WORLD)
            buf.zero_()
            torch.cuda.synchronize(dev)
            ptrs = [int(x) for x in hdl.buffer_ptrs]
            st = dict(buf=buf, hdl=hdl, stride=stride, world=world, rank=rank, g=[buf[0:n], buf[stride:stride + n]],
                      gptrs=torch.tensor(ptrs, dtype=torch.int64, device=dev),
                      fptrs=torch.tensor([x + 2 * stride * 4 for x in ptrs], dtype=torch.int64, device=dev),
                      cur=torch.zeros(1, dtype=torch.int32, device=dev), epoch=torch.zeros(1, dtype=torch.int32, device=dev),
                      err=torch.zeros(1, dtype=torch.int32, device=dev), gsum=torch.zeros(stride, dtype=torch.float32, device=dev))
        except Exception as ex:      # noqa: BLE001 -- any failure means "use NCCL"
            ok = 0
            print(f"[ppo] peer-memory all-reduce unavailable on rank {rank}: {type(ex).__name__}: {ex}", flush=True)
        agree = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)       # all ranks or none; also orders the zeroing before the first flag write
        torch.cuda.synchronize(dev)
        if int(agree.item()) == 0:
            return False
        self.peer = st
        return True

    def check_peer_error(self) -> None:
        if self.peer is not None and int(self.peer["err"].item()) != 0:
            raise RuntimeError("ackb_ppo_clip_adam_allreduce: a peer rank did not arrive within the time limit")

    def act(self, obs: torch.Tensor, action: torch.Tensor, logp: torch.Tensor, value: torch.Tensor, seed: int, step: int) -> None:
        """Fused rollout forward (csrc/ackb_ppo.cu: ppo_act_kernel): fills action (unclipped sample), logp and value in place."""
        c = self.ct
        ptr = lambda t: c.c_void_p(t.data_ptr())
        rc = self.L.ackb_ppo_act_pitched(ptr(obs), int(obs.stride(0)), int(obs.shape[0]), self.obs_dim, ptr(self.flat_p), None, ptr(value),
                                         ptr(action), ptr(logp), int(seed) & 0xFFFFFFFFFFFFFFFF, int(step) & 0xFFFFFFFF, 0,
                                         c.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"ackb_ppo_act failed with code {rc}")

    def value(self, obs: torch.Tensor, value: torch.Tensor) -> None:
        """V(obs) only (bootstrap values of terminal observations): the policy net is skipped."""
        c = self.ct
        ptr = lambda t: c.c_void_p(t.data_ptr())
        rc = self.L.ackb_ppo_act_pitched(ptr(obs), int(obs.stride(0)), int(obs.shape[0]), self.obs_dim, ptr(self.flat_p), None, ptr(value), None,
                                         None, 0, 0, 1, c.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"ackb_ppo_act failed with code {rc}")

    def bootstrap(self, terminal_obs: torch.Tensor, term: torch.Tensor, trunc: torch.Tensor, rew: torch.Tensor, rew_out: torch.Tensor,
                  done_out: torch.Tensor) -> None:
        """rew_out = rew + gamma V(terminal_obs) on time-limit truncation (else rew), done_out = term | trunc, in one launch
        (ackb_ppo_bootstrap); term / trunc are uint8."""
        c = self.ct
        ptr = lambda t: c.c_void_p(t.data_ptr())
        assert term.dtype == torch.uint8 and trunc.dtype == torch.uint8 and rew_out.is_contiguous() and done_out.is_contiguous()
        rc = self.L.ackb_ppo_bootstrap_pitched(ptr(terminal_obs), int(terminal_obs.stride(0)), ptr(term), ptr(trunc), ptr(rew), int(rew.shape[0]),
                                               self.obs_dim, ptr(self.flat_p), self.cfg.gamma, ptr(rew_out), ptr(done_out),
                                               c.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"ackb_ppo_bootstrap failed with code {rc}")

    def gae(self, rew: torch.Tensor, val: torch.Tensor, done: torch.Tensor, last_val: torch.Tensor, adv: torch.Tensor,
            ret: torch.Tensor) -> None:
        """compute_gae as one kernel (ackb_ppo_gae): [T, N] contiguous float32 arrays, results written into adv / ret."""
        c = self.ct
        ptr = lambda t: c.c_void_p(t.data_ptr())
        assert all(t.is_contiguous() and t.dtype == torch.float32 for t in (rew, val, done, last_val, adv, ret))
        rc = self.L.ackb_ppo_gae(ptr(rew), ptr(val), ptr(done), ptr(last_val), int(rew.shape[0]), int(rew.shape[1]), self.cfg.gamma,
                                 self.cfg.gae_lambda, ptr(adv), ptr(ret), c.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"ackb_ppo_gae failed with code {rc}")

    def new_epoch(self, batch: Dict[str, torch.Tensor], seed: int, epoch: int) -> Dict[str, torch.Tensor]:
        """Index mode without torch.randperm: the epoch's permutation is written straight into the persistent index buffer by
        ackb_ppo_permutation (keyed Feistel bijection, no sort; 0.22 -> 0.01 ms per million samples)."""
        n = int(batch["obs"].shape[0])
        if getattr(self, "_perm", None) is None or self._perm.shape[0] != n:
            self._perm = torch.empty(n, dtype=torch.int64, device=self.device)
        c = self.ct
        rc = self.L.ackb_ppo_permutation(c.c_void_p(self._perm.data_ptr()), n, int(seed) & 0xFFFFFFFFFFFFFFFF, int(epoch) & 0xFFFFFFFF,
                                         c.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"ackb_ppo_permutation failed with code {rc}")
        return batch

    def shuffle_epoch(self, batch: Dict[str, torch.Tensor], perm: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Prepare one epoch's minibatches as (start, count) ranges.  Index mode (default): the permutation is kept in a persistent
        index buffer and the gradient kernel gathers the rollout rows through it (its cp.async prefetch hides the gather), so
        nothing is copied.  Copy mode (ACKB_PPO_SHUFFLE=copy): a permuted copy of the rollout, minibatches are contiguous rows."""
        if self.index_mode:
            if getattr(self, "_perm", None) is None or self._perm.shape != perm.shape:
                self._perm = torch.empty_like(perm)
            self._perm.copy_(perm)
            return batch
        if getattr(self, "_shuf", None) is None or self._shuf["obs"].shape != batch["obs"].shape:
            self._shuf = {k: torch.empty_like(batch[k]) for k in _ROLLOUT_KEYS}
        for k, dst in self._shuf.items():
            torch.index_select(batch[k], 0, perm, out=dst)
        return self._shuf

    def _resolve(self, batch: Dict[str, torch.Tensor], idx):
        """(arrays, count, index tensor or None) for idx = int64 rows or a (start, count) range."""
        if isinstance(idx, tuple):
            lo, n = idx
            if self.index_mode:
                return batch, n, self._perm[lo:lo + n]
            return {k: batch[k][lo:lo + n] for k in _ROLLOUT_KEYS}, n, None
        idx = idx.contiguous()
        return batch, int(idx.numel()), idx

    def capture(self, shuffled: Dict[str, torch.Tensor], slots, epoch_graph: bool = False) -> None:
        """CUDA graphs for the minibatch slots (start, count) of the shuffled rollout: per slot {advantage statistics + gradient
        kernel}, and one graph for {gradient clipping + Adam}.  Needs an optimiser built with capturable=True.  The warm-up steps
        are undone (weights and optimiser state restored in place).  epoch_graph: additionally ONE graph holding all optimiser
        steps of an epoch back to back (run_epoch; needs accumulate_diag, and no NCCL call between gradient and optimiser step,
        i.e. a single rank or the peer-memory all-reduce)."""
        dev = self.device
        self._cap_ptrs = tuple(shuffled[k].data_ptr() for k in _ROLLOUT_KEYS)
        saved_p = self.flat_p.clone()
        saved_s = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in self.opt.state.get(p, {}).items()} for p in self.params}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                for sl in slots:
                    self._grad(shuffled, sl, self._mbc & 1)
                    self._clip_step()
                    self._mbc += 1
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.g1, self.g2 = {}, torch.cuda.CUDAGraph()
        self._slot_parity = {sl: i & 1 for i, sl in enumerate(slots)}      # gradient buffer baked into the slot's graph (peer path)
        for sl in slots:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._grad(shuffled, sl, self._slot_parity[sl])
            self.g1[sl] = g
        with torch.cuda.graph(self.g2):
            self._clip_step()
        self.g_epoch = None
        if epoch_graph and self.accumulate_diag and (len(slots) % 2 == 0 or self.peer is None):
            self.g_epoch, self._epoch_slots = torch.cuda.CUDAGraph(), tuple(slots)
            with torch.cuda.graph(self.g_epoch):
                for i, sl in enumerate(slots):
                    self._grad(shuffled, sl, i & 1)
                    self._clip_step()
        with torch.no_grad():
            self.flat_p.copy_(saved_p)
            for p in self.params:
                for k, v in self.opt.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        if k in saved_s[id(p)]:
                            v.copy_(saved_s[id(p)][k])
                        else:
                            v.zero_()
        torch.cuda.synchronize(dev)

    def _clip_step(self):
        if not self.native_opt:
            torch.nn.utils.clip_grad_norm_(self.params, self.cfg.max_grad_norm)
            self.opt.step()
            return
        c, g, st = self.ct, self.opt.param_groups[0], self.opt.state[self.flat_param]
        ptr = lambda t: c.c_void_p(t.data_ptr())
        if self.peer is not None:      # all-reduce over peer memory + clip + Adam in one kernel
            pr = self.peer
            rc = self.L.ackb_ppo_clip_adam_allreduce(ptr(self.flat_p), ptr(pr["gptrs"]), ptr(pr["fptrs"]), ptr(pr["cur"]), int(pr["stride"]),
                                                     int(pr["world"]), int(pr["rank"]), ptr(pr["gsum"]), ptr(pr["epoch"]), ptr(pr["err"]),
                                                     ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), ptr(st["step"]), int(self.flat_p.numel()),
                                                     float(self.cfg.max_grad_norm), float(g["lr"]), float(g["betas"][0]),
                                                     float(g["betas"][1]), float(g["eps"]),
                                                     c.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            if rc != 0:
                raise RuntimeError(f"ackb_ppo_clip_adam_allreduce failed with code {rc}")
            return
        rc = self.L.ackb_ppo_clip_adam(ptr(self.flat_p), ptr(self.flat_g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), ptr(st["step"]),
                                       int(self.flat_p.numel()), float(self.cfg.max_grad_norm), float(g["lr"]), float(g["betas"][0]),
                                       float(g["betas"][1]), float(g["eps"]), c.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"ackb_ppo_clip_adam failed with code {rc}")

    def _grad(self, batch: Dict[str, torch.Tensor], idx, parity: int = 0):
        """Advantage statistics of the minibatch + the gradient kernel (fills flat_g -- or, on the peer all-reduce path, gradient
        buffer `parity` of the symmetric allocation -- and diag)."""
        c, cfg = self.ct, self.cfg
        ptr = lambda t: c.c_void_p(t.data_ptr())
        gbuf = self.flat_g
        if self.peer is not None:
            gbuf = self.peer["g"][parity & 1]
            self.peer["cur"].fill_(parity & 1)
        view, n, rows = self._resolve(batch, idx)
        stream = c.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        # advantage statistics of the minibatch + gradient (one call; on the tcgen05 path the statistics ride in the prologue launch)
        rc = self.L.ackb_ppo_minibatch_grad_stats(ptr(view["obs"]), int(view["obs"].stride(0)), ptr(view["act"]), ptr(view["logp"]),
                                                  ptr(view["adv"]), ptr(view["ret"]), ptr(rows) if rows is not None else None, n,
                                                  self.obs_dim, ptr(self.adv_stats), ptr(self.adv_ws), ptr(self.flat_p), ptr(gbuf), ptr(self.diag),
                                                  cfg.clip_range, cfg.vf_coef, cfg.ent_coef,
                                                  self._call_mode(), stream)
        if rc != 0:
            raise RuntimeError(f"ackb_ppo_minibatch_grad failed with code {rc}")

    def _call_mode(self) -> int:
        """Per-call arithmetic mode (+ the diagnostics flag).  The flag needs a concrete mode: "default" (-1: the library's
        process-wide setting, ACKB_PPO_TC) is resolved here."""
        m = self.mode
        if self.accumulate_diag:
            if m < 0:
                m = 0 if os.environ.get("ACKB_PPO_TC") == "0" else 1
            m |= self.DIAG_ACCUMULATE
        return m

    def run(self, batch: Dict[str, torch.Tensor], idx, world: int) -> int:
        """idx: int64 index tensor (rows of `batch`), or a (start, count) range of the epoch prepared by shuffle_epoch."""
        g1 = getattr(self, "g1", None)
        parity = self._mbc & 1
        graphs = g1 is not None and isinstance(idx, tuple) and idx in g1 and self._same_arrays(batch)
        if graphs and self.peer is not None and self._slot_parity.get(idx) != parity:
            graphs = False      # the buffers must alternate from step to step: this slot's graph has the other one baked in
        if graphs:
            g1[idx].replay()
        else:
            self._grad(batch, idx, parity)
        self._mbc += 1
        nbytes = 0
        if world > 1 and self.peer is not None:
            nbytes = self.flat_g.numel() * 4      # moved by the optimiser-step kernel over peer memory
        elif world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
            self.flat_g.div_(world)
            nbytes = self.flat_g.numel() * 4
        if graphs:
            self.g2.replay()
        else:
            self._clip_step()
        return nbytes

    def run_epoch(self, batch: Dict[str, torch.Tensor], slots, world: int) -> int:
        """All optimiser steps of one epoch (slots = the (start, count) ranges of its minibatches): one graph launch when capture()
        recorded an epoch graph for exactly these slots and arrays, else step by step.  Returns the all-reduced bytes."""
        ge = getattr(self, "g_epoch", None)
        if (ge is not None and tuple(slots) == self._epoch_slots and self._same_arrays(batch) and (world == 1 or self.peer is not None)
                and (self.peer is None or (self._mbc & 1) == 0)):
            ge.replay()
            self._mbc += len(slots)
            return len(slots) * self.flat_g.numel() * 4 if world > 1 else 0
        return sum(self.run(batch, sl, world) for sl in slots)

    def _same_arrays(self, batch: Dict[str, torch.Tensor]) -> bool:
        """True if `batch` is made of the very arrays the graphs were captured on."""
        return tuple(batch[k].data_ptr() for k in _ROLLOUT_KEYS) == getattr(self, "_cap_ptrs", None)


class PPOTrainer:
    """Rollout collection on the batched CUDA environment + PPO updates; one instance per rank."""

    def __init__(self, env, cfg: PPOConfig = PPOConfig(), seed: int = 0, use_cuda_graphs: bool = True, learner: str = "fused",
                 learner_mode: str = "tcgen05"):
        self.env, self.cfg = env, cfg
        self.device = env.device
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        torch.manual_seed(seed)                       # identical initial weights on every rank
        self.policy = ActorCritic(env.obs_dim).to(self.device)
        torch.manual_seed(seed * 1000003 + self.rank)  # distinct exploration noise per rank
        # learner: "fused" = hand-written gradient kernel (default), "graph" = torch ops captured in CUDA graphs, "eager" = torch ops
        self.learner = learner if self.device.type == "cuda" else "eager"
        self.use_graphs = use_cuda_graphs and self.learner == "graph"
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=cfg.learning_rate, eps=cfg.adam_eps,
                                    capturable=self.use_graphs or (use_cuda_graphs and self.learner == "fused"))
        self.fused_graphs = use_cuda_graphs and self.learner == "fused"
        self.graphed: Optional[GraphedMinibatchStep] = None
        T, N, D = cfg.n_steps, env.num_envs, env.obs_dim
        f = dict(device=self.device, dtype=torch.float32)
        import inspect
        self._env_takes_obs_out = "obs_out" in inspect.signature(env.step).parameters
        # fused learner: observation rows padded to 80 floats (16-byte aligned rows: vector gathers over whole 32-byte sectors in the
        # gradient kernel); the environment writes straight into the padded rollout slots (env.step(obs_out=...))
        self.pitch = 80 if (self.learner == "fused" and D < 80 and self._env_takes_obs_out and "obs_out" in inspect.signature(env.reset).parameters) else D
        self.buf = dict(obs=torch.zeros((T, N, self.pitch), **f), act=torch.empty((T, N, 2), **f), logp=torch.empty((T, N), **f),
                        val=torch.empty((T, N), **f), rew=torch.empty((T, N), **f), done=torch.empty((T, N), **f))
        if self.pitch != D:
            self.obs = torch.zeros((N, self.pitch), **f)
            env.reset(obs_out=self.obs)
        else:
            self.obs = env.reset().clone()
        self.num_timesteps = 0
        self._noise_seed = (seed * 1000003 + self.rank) * 2654435761 + 12345
        self._act_step = 0
        self._tv = torch.empty(N, **f)
        self._adv, self._ret = torch.empty((T, N), **f), torch.empty((T, N), **f)    # persistent: CUDA graphs are captured on them
        if self.learner == "fused":     # flat parameter buffers exist from the start: the rollout forward uses them too
            self.graphed = FusedMinibatchStep(self.policy, self.opt, cfg, D, self.device, mode=learner_mode)
            self.graphed.accumulate_diag = True      # no per-step accumulation launch in update()
            self.peer_allreduce = self.graphed.enable_peer_allreduce(self.world, self.rank) if (self.world > 1 and self.device.type == "cuda") else False

    def grad_kernel_seconds(self, reps: int = 5) -> Optional[float]:
        """Average device time of ONE launch of the minibatch-gradient kernel on the current rollout (CUDA events on the launching
        stream; weights and gradients are left untouched apart from flat_g / diag, which the next update overwrites)."""
        f = self.graphed
        if not isinstance(f, FusedMinibatchStep) or getattr(f, "_perm", None) is None:
            return None
        b = self.buf
        flat = dict(obs=b["obs"].flatten(0, 1), act=b["act"].flatten(0, 1), logp=b["logp"].flatten(), adv=self._adv.flatten(), ret=self._ret.flatten())
        n = flat["obs"].shape[0]
        mb = max(1, n // self.cfg.minibatches)
        c = f.ct
        ptr = lambda t: c.c_void_p(t.data_ptr())
        rows = f._perm[:mb]
        stream = c.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        def launch():
            rc = f.L.ackb_ppo_minibatch_grad_pitched(ptr(flat["obs"]), int(flat["obs"].stride(0)), ptr(flat["act"]), ptr(flat["logp"]), ptr(flat["adv"]),
                                                     ptr(flat["ret"]), ptr(rows), mb, f.obs_dim, ptr(f.adv_stats), ptr(f.flat_p), ptr(f.flat_g), ptr(f.diag),
                                                     self.cfg.clip_range, self.cfg.vf_coef, self.cfg.ent_coef, f.mode, stream)
            assert rc == 0
        launch()
        torch.cuda.synchronize(self.device)
        e0.record()
        for _ in range(reps):
            launch()
        e1.record()
        torch.cuda.synchronize(self.device)
        return e0.elapsed_time(e1) * 1e-3 / reps

    @torch.no_grad()
    def collect(self) -> float:
        cfg, env, b = self.cfg, self.env, self.buf
        t0 = time.perf_counter()
        torch.cuda.nvtx.range_push("ppo.collect")        # NVTX phases: rollout / update
        for t in range(cfg.n_steps):
            obs = sanitize_obs(self.obs)
            fused = self.graphed if isinstance(self.graphed, FusedMinibatchStep) else None
            direct = fused is not None and self._env_takes_obs_out
            if fused is not None:     # one launch: both MLPs, Gaussian sample, log-probability, value -> straight into the buffers
                if not direct or t == 0:
                    b["obs"][t].copy_(obs)
                fused.act(b["obs"][t], b["act"][t], b["logp"][t], b["val"][t], self._noise_seed, self._act_step)
                self._act_step += 1
                act = b["act"][t]
            else:
                act, logp, val = self.policy.act(obs)
                b["obs"][t], b["act"][t], b["logp"][t], b["val"][t] = obs, act, logp, val
            if direct:      # the step kernel writes the next observation straight into the next rollout slot (or self.obs at the end)
                # (the step kernel clips the action to the Box bounds itself, like the reference env: no clamp launch here)
                nobs, rew, term, trunc, info = env.step(act, obs_out=b["obs"][t + 1] if t + 1 < cfg.n_steps else self.obs)
            else:
                nobs, rew, term, trunc, info = env.step(torch.clamp(act, -1.0, 1.0))     # SB3 clips to the Box bounds
            # bootstrap with V(terminal observation) on time-limit truncation; evaluated for every environment and masked, so
            # that the rollout loop has no device->host synchronisation (rows of environments that did not finish are stale, unused)
            if fused is not None and term.dtype == torch.uint8:
                fused.bootstrap(info["terminal_observation"], term, trunc, rew, b["rew"][t], b["done"][t])
            else:
                only_trunc = ((trunc != 0) & (term == 0)).float()
                tv = self.policy.value(sanitize_obs(info["terminal_observation"]))
                b["rew"][t] = rew + cfg.gamma * tv * only_trunc
                b["done"][t] = ((term != 0) | (trunc != 0)).float()
            if not direct:
                self.obs[:, :nobs.shape[1]].copy_(nobs)
        self.num_timesteps += cfg.n_steps * env.num_envs * self.world
        torch.cuda.synchronize(self.device)
        torch.cuda.nvtx.range_pop()
        return time.perf_counter() - t0

    def update(self) -> Dict[str, float]:
        cfg, b = self.cfg, self.buf
        t0 = time.perf_counter()
        torch.cuda.nvtx.range_push("ppo.update")
        with torch.no_grad():
            if isinstance(self.graphed, FusedMinibatchStep):      # value forward + one GAE kernel into the persistent arrays
                self.graphed.value(sanitize_obs(self.obs), self._tv)
                self.graphed.gae(b["rew"], b["val"], b["done"], self._tv, self._adv, self._ret)
                adv, ret = self._adv, self._ret
            else:
                last_val = self.policy.value(sanitize_obs(self.obs))
                adv, ret = compute_gae(b["rew"], b["val"], b["done"], last_val, cfg.gamma, cfg.gae_lambda)
        flat = dict(obs=b["obs"].flatten(0, 1), act=b["act"].flatten(0, 1), logp=b["logp"].flatten(), adv=adv.flatten(), ret=ret.flatten())
        if self.learner == "fused" and self.fused_graphs and getattr(self.graphed, "g1", None) is None:
            n = flat["obs"].shape[0]
            mb = max(1, n // cfg.minibatches)
            sh = self.graphed.shuffle_epoch(flat, torch.arange(n, device=self.device))
            self.graphed.capture(sh, [(i * mb, mb) for i in range(cfg.minibatches)],
                                 epoch_graph=(self.world == 1 or self.graphed.peer is not None) and os.environ.get("ACKB_PPO_EPOCH_GRAPH", "1") != "0")
        if self.use_graphs and self.graphed is None:
            n = flat["obs"].shape[0]
            self.graphed = GraphedMinibatchStep(self.policy, self.opt, cfg, max(1, n // cfg.minibatches), flat["obs"].shape[1], self.device)
        st = ppo_update(self.policy, self.opt, flat, cfg, self.world, graphed=self.graphed, perm_seed=self._noise_seed)
        torch.cuda.synchronize(self.device)
        if isinstance(self.graphed, FusedMinibatchStep):
            self.graphed.check_peer_error()
        torch.cuda.nvtx.range_pop()
        st["update_s"] = time.perf_counter() - t0
        return st

    def train(self, total_timesteps: int, log=print, log_all_ranks: bool = False) -> Dict[str, float]:
        out = {}
        it = 0
        while self.num_timesteps < total_timesteps:
            self.env.stats_reset()
            tr = self.collect()
            st = self.update()
            es = reduce_stats(self.env.stats(), device=self.device)
            n_roll = self.cfg.n_steps * self.env.num_envs * self.world
            out = dict(iteration=it, timesteps=self.num_timesteps, rollout_s=tr, update_s=st["update_s"],
                       rollout_env_steps_per_s=n_roll / tr, end_to_end_env_steps_per_s=n_roll / (tr + st["update_s"]),
                       episodes=es["episodes"], successes=es["successes"],
                       ep_rew_mean=(es["return_sum"] / es["episodes"]) if es["episodes"] else float("nan"),
                       ep_len_mean=(es["length_sum"] / es["episodes"]) if es["episodes"] else float("nan"),
                       **{k: st[k] for k in ("pg_loss", "v_loss", "entropy", "approx_kl", "clip_frac", "allreduce_bytes")})
            if (self.rank == 0 or log_all_ranks) and log:      # log_all_ranks: the callback runs collectives (evaluation)
                log(out)
            it += 1
        return out

    def save(self, path: str):
        """Checkpoint: policy + optimiser state (replaces CheckpointCallback / model.save, train.py:140-144,182-183)."""
        if self.rank == 0:
            torch.save({"policy": self.policy.state_dict(), "optimizer": self.opt.state_dict(), "num_timesteps": self.num_timesteps}, path)
