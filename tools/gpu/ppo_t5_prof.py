#!/usr/bin/env python
"""Phase breakdown (clock64, CTA 0) of the tcgen05 PPO gradient kernel; needs the -DACKB_T5_PROFILE build:
    ACKB_LIB=tools/gpu/libackb_prof.so python tools/gpu/ppo_t5_prof.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from mujoco_playground_b200 import _lib  # noqa: E402
from mujoco_playground_b200.ppo import ActorCritic, FusedMinibatchStep, PPOConfig  # noqa: E402

dev = torch.device("cuda:0")
D, big = 79, 262144
pol = ActorCritic(D).to(dev)
f = FusedMinibatchStep(pol, torch.optim.SGD(pol.parameters(), lr=0.0), PPOConfig(), D, dev, mode="tcgen05")
b2 = dict(obs=torch.randn(big, D, device=dev), act=torch.randn(big, 2, device=dev).clamp(-1, 1), logp=torch.randn(big, device=dev) * 0.3 - 2.0,
          adv=torch.randn(big, device=dev), ret=torch.randn(big, device=dev))
ix = torch.randperm(big, device=dev)
L = _lib.load()
out = (ctypes.c_longlong * 16)()
f._grad(b2, ix)
L.ackb_ppo_t5_profile(out)
for _ in range(5):
    f._grad(b2, ix)
L.ackb_ppo_t5_profile(out)
names = ["gather+round", "X->TMEM + mma1", "E1 (H1)", "mma2", "E2a heads", "loss", "E2b dZ2", "mma3+4", "E3 dZ1", "mma5"]
print(f"per launch: setup + weights of a net {out[10] / 5 / 2:.0f} cycles per net, accumulator flush {out[11] / 5 / 2:.0f} cycles per net")
tot = sum(out[i] for i in range(10))
for i, n in enumerate(names):
    print(f"{n:18s} {out[i] / 5 / 28:9.0f} cycles per tile-pass  {100.0 * out[i] / tot:5.1f} %")
print(f"total {tot / 5 / 28:.0f} cycles per tile-pass ({tot / 5:.0f} per launch)")
