#!/usr/bin/env python
"""Gradient of the tcgen05 PPO kernel against the fp32 CUDA-core kernel and the mma.sync TF32 kernel, per parameter block."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from mujoco_playground_b200 import _lib  # noqa: E402
from mujoco_playground_b200.ppo import ActorCritic, FusedMinibatchStep, PPOConfig  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 79
    torch.manual_seed(1)
    batch = dict(obs=torch.randn(n, D, device=dev), act=torch.randn(n, 2, device=dev).clamp(-1, 1), logp=torch.randn(n, device=dev) * 0.3 - 2.0,
                 adv=torch.randn(n, device=dev), ret=torch.randn(n, device=dev))
    cfg = PPOConfig(max_grad_norm=1e30)
    pol = ActorCritic(D).to(dev)
    with torch.no_grad():
        for p in pol.parameters():
            p.add_(0.05 * torch.randn_like(p))
    opt = torch.optim.SGD(pol.parameters(), lr=0.0)
    f = FusedMinibatchStep(pol, opt, cfg, D, dev)
    idx = torch.randperm(n, device=dev)[: n - 7]
    res = {}
    for name, mode in (("fp32", 0), ("tf32", 1), ("tcgen05", 2)):
        f.mode = mode
        f.run(batch, idx, world=1)
        torch.cuda.synchronize()
        res[name] = (f.flat_g.clone(), f.diag.clone())
    names = ["W1p", "b1p", "W2p", "b2p", "W1v", "b1v", "W2v", "b2v", "Wa", "ba", "Wv", "bv", "ls"]
    sizes = [64 * D, 64, 4096, 64, 64 * D, 64, 4096, 64, 128, 2, 64, 1, 2]
    off = 0
    ok = True
    for nm, sz in zip(names, sizes):
        ref = res["fp32"][0][off:off + sz]
        sc = max(1e-9, ref.abs().max().item())
        e1 = (res["tf32"][0][off:off + sz] - ref).abs().max().item() / sc
        e2 = (res["tcgen05"][0][off:off + sz] - ref).abs().max().item() / sc
        print(f"{nm:4s} scale {sc:.3e}  rel err tf32 {e1:.2e}  tcgen05 {e2:.2e}")
        ok = ok and e2 < 2e-2
        off += sz
    print("diag fp32   ", res["fp32"][1].tolist())
    print("diag tcgen05", res["tcgen05"][1].tolist())
    # timing
    for name, mode in (("tf32", 1), ("tcgen05", 2)):
        f.mode = mode
        big = 262144
        b2 = dict(obs=torch.randn(big, D, device=dev), act=torch.randn(big, 2, device=dev).clamp(-1, 1), logp=torch.randn(big, device=dev) * 0.3 - 2.0,
                  adv=torch.randn(big, device=dev), ret=torch.randn(big, device=dev))
        ix = torch.randperm(big, device=dev)
        f.run(b2, ix, world=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f._grad(b2, ix)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 10:.4f} ms per 262144-sample minibatch (incl. adv stats + memsets)")
    print("T5 CHECK", "OK" if ok else "FAILED")


if __name__ == "__main__":
    main()
