#!/usr/bin/env python
"""Peer-memory gradient all-reduce inside the optimiser-step kernel (ackb_ppo_clip_adam_allreduce) against the NCCL path:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu/ppo_peer_check.py
Every rank trains two PPOTrainers from the same seed for a few iterations, one per path; prints whether the peer path was
taken, the largest parameter difference between the paths, whether all ranks hold bit-identical parameters, and the update time."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mujoco_playground_b200 import BatchedAckermannEnv  # noqa: E402
from mujoco_playground_b200.ppo import PPOConfig, PPOTrainer  # noqa: E402


def run(peer: bool, n_envs: int, iters: int, rank: int, local: int):
    os.environ["ACKB_PPO_PEER_ALLREDUCE"] = "1" if peer else "0"
    env = BatchedAckermannEnv(n_envs, device=f"cuda:{local}", seed=5, env_id_base=rank * n_envs)
    tr = PPOTrainer(env, PPOConfig(n_steps=16), seed=3)
    env.reset()
    ts = []
    for _ in range(iters):
        tr.collect()
        t0 = time.perf_counter()
        tr.update()
        ts.append(time.perf_counter() - t0)
    flat = tr.graphed.flat_p.detach().clone()
    used = bool(getattr(tr, "peer_allreduce", False))
    env.close()
    return flat, used, min(ts[1:]) * 1e3


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    p_nccl, used0, t_nccl = run(False, n_envs, 4, rank, local)
    p_peer, used1, t_peer = run(True, n_envs, 4, rank, local)
    diff = (p_peer - p_nccl).abs().max().item()
    scale = p_nccl.abs().max().item()
    ref = p_peer.clone()
    dist.broadcast(ref, 0)
    same = torch.equal(ref, p_peer)
    flags = torch.tensor([int(same), int(used1), int(not used0)], device=p_peer.device)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world {world}: peer path taken {bool(flags[1])}, NCCL run used NCCL {bool(flags[2])}, ranks bit-identical {bool(flags[0])}, "
              f"max |param(peer) - param(nccl)| = {diff:.3e} (scale {scale:.2f}), update {t_nccl:.3f} ms (NCCL) vs {t_peer:.3f} ms (peer)")
        ok = bool(flags[0]) and bool(flags[1]) and diff < 2e-3 * max(1.0, scale)
        print("PEER CHECK", "OK" if ok else "FAILED")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
