"""Training driver: the batched counterpart of the reference's src/rl/train.py (same flag names where they exist).

    python -m mujoco_playground_b200.train --algo ppo --timesteps 20000000 --num-envs 65536
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m mujoco_playground_b200.train --algo ppo ...

--algo random reproduces train_with_custom_algo (train.py:189-227): a random-policy rollout;
--algo ppo    reproduces train_with_stable_baselines3 for PPO (train.py:44-186).  SAC / TD3 are out of scope.
"""
from __future__ import annotations

import argparse
import json
import os
import time

import torch
import torch.distributed as dist

from .env import BatchedAckermannEnv
from .ppo import PPOConfig, PPOTrainer
from .shard import rank_seed, reduce_stats


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", default="random", choices=["random", "ppo"])          # train.py:232
    ap.add_argument("--timesteps", type=int, default=100000)                          # train.py:236
    ap.add_argument("--max-velocity", type=float, default=1.0)
    ap.add_argument("--goal-threshold", type=float, default=0.5)
    ap.add_argument("--learning-rate", type=float, default=3e-4)
    ap.add_argument("--save-path", default="")
    ap.add_argument("--num-envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--frame-skip", type=int, default=1)
    ap.add_argument("--n-steps", type=int, default=16)
    ap.add_argument("--n-epochs", type=int, default=10)
    ap.add_argument("--minibatches", type=int, default=4)
    ap.add_argument("--model", default="v2", choices=["v2", "scene"])
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    torch.backends.cuda.matmul.allow_tf32 = True      # learner GEMMs (18 757-parameter MLP) on TF32 tensor cores; the simulator is unaffected
    torch.backends.cudnn.allow_tf32 = True
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    env = BatchedAckermannEnv(a.num_envs, device=dev, frame_skip=a.frame_skip, seed=rank_seed(a.seed, rank), model=a.model,
                              max_linear_velocity=a.max_velocity, goal_distance_threshold=a.goal_threshold)
    if a.algo == "random":
        env.reset()
        steps = max(1, a.timesteps // (a.num_envs * world))
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(steps):
            env.step(None)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        st = reduce_stats(env.stats(), device=dev)
        if rank == 0:
            print(json.dumps({"algo": "random", "env_steps": st["env_steps"], "env_steps_per_s": st["env_steps"] / dt, "episodes": st["episodes"],
                              "ep_rew_mean": st["return_sum"] / max(1, st["episodes"])}))
    else:
        cfg = PPOConfig(n_steps=a.n_steps, n_epochs=a.n_epochs, minibatches=a.minibatches, learning_rate=a.learning_rate)
        tr = PPOTrainer(env, cfg, seed=a.seed)
        tr.train(a.timesteps, log=lambda d: print(json.dumps(d), flush=True))
        if a.save_path:
            tr.save(a.save_path)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
