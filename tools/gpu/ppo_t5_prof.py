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
N, D, big = 1048576, 79, 262144
pol = ActorCritic(D).to(dev)
f = FusedMinibatchStep(pol, torch.optim.SGD(pol.parameters(), lr=0.0), PPOConfig(), D, dev, mode="tcgen05")
obs = torch.zeros(N, 80, device=dev)        # 80-float pitched rows, as PPOTrainer stores the rollout
obs[:, :D] = torch.randn(N, D, device=dev)
b2 = dict(obs=obs, act=torch.randn(N, 2, device=dev).clamp(-1, 1), logp=torch.randn(N, device=dev) * 0.3 - 2.0, adv=torch.randn(N, device=dev),
          ret=torch.randn(N, device=dev))
ix = torch.randperm(N, device=dev)[:big].contiguous()
L = _lib.load()
out = (ctypes.c_longlong * 32)()
f._grad(b2, ix)
L.ackb_ppo_t5_profile(out)
for _ in range(5):
    f._grad(b2, ix)
L.ackb_ppo_t5_profile(out)
names = ["X: next loads+bar", "mma1", "E1 (H1)", "mma2", "E2a heads", "loss", "E2b dZ2", "mma3+4+6", "E3 dZ1", "mma5 (value net)"]
tiles = 14                                   # tiles of CTA 0: ceil(2048 / 148)
print(f"per launch: setup + first gather {out[10] / 5:.0f} cycles, accumulator flush {out[11] / 5:.0f} cycles")
tot = sum(out[i] for i in range(10))
for i, n in enumerate(names):
    per = out[i] / 5 / tiles / (1 if i == 0 else 2)
    print(f"{n:18s} {per:9.0f} cycles per tile{'' if i == 0 else ' and net'}  {100.0 * out[i] / tot:5.1f} %")
print(f"inside the 4 MMA steps per tile and net: own tcgen05.wait::st {out[15] / 5 / tiles / 2:.0f}, fences {out[12] / 5 / tiles / 2:.0f}, CTA barrier {out[13] / 5 / tiles / 2:.0f}, "
      f"issue by thread 0: (1) {out[16] / 5 / tiles / 2:.0f}, (2) {out[17] / 5 / tiles / 2:.0f}, (3)(4)(6) {out[18] / 5 / tiles / 2:.0f}, (5) {out[19] / 5 / tiles / 2:.0f} cycles (the mma rows above are what remains: execution + completion wait)")
print(f"completion wait after the issue (mbarrier + fence), per tile and net: (1) {out[24] / 5 / tiles / 2:.0f}, (2) {out[25] / 5 / tiles / 2:.0f}, (3)(4)(6) {out[26] / 5 / tiles / 2:.0f}, "
      f"(5) {out[27] / 5 / tiles / 2:.0f} cycles (now excluded from the mma rows above, which keep what follows the wait: weight reload issue, loads)")
tot += out[12] + out[13] + out[15] + sum(out[16:20]) + sum(out[24:28])
print(f"X step per tile: convert + store to smem {out[21] / 5 / tiles:.0f}, issue next loads + barrier {out[0] / 5 / tiles:.0f}, smem -> TMEM {out[22] / 5 / tiles:.0f}")
tot += out[21] + out[22]
print(f"total {tot / 5 / tiles:.0f} cycles per tile, both nets ({tot / 5:.0f} per launch)")
