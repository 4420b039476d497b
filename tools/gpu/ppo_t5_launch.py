#!/usr/bin/env python
"""A few launches of the tcgen05 PPO gradient kernel on a 262144-sample minibatch gathered by index from a 1 M-row rollout with
80-float pitched rows (what PPOTrainer feeds it); target of the ncu capture (tools/gpu/prof_r02.sh)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from mujoco_playground_b200.ppo import ActorCritic, FusedMinibatchStep, PPOConfig  # noqa: E402

dev = torch.device("cuda:0")
N, D, mb = 1048576, 79, 262144
mode = sys.argv[1] if len(sys.argv) > 1 else "tcgen05"
obs = torch.zeros(N, 80, device=dev)
obs[:, :D] = torch.randn(N, D, device=dev)
batch = dict(obs=obs, act=torch.randn(N, 2, device=dev).clamp(-1, 1), logp=torch.randn(N, device=dev) * 0.3 - 2.0, adv=torch.randn(N, device=dev),
             ret=torch.randn(N, device=dev))
pol = ActorCritic(D).to(dev)
f = FusedMinibatchStep(pol, torch.optim.SGD(pol.parameters(), lr=0.0), PPOConfig(), D, dev, mode=mode)
idx = torch.randperm(N, device=dev)[:mb].contiguous()
for _ in range(3):
    f._grad(batch, idx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    f._grad(batch, idx)
e1.record()
torch.cuda.synchronize()
print(f"{mode}: {e0.elapsed_time(e1) / 5:.4f} ms per launch pair (adv stats + gradient kernel)")
