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
from .shard import reduce_stats


def main(argv=None):
    ap = argparse.ArgumentParser(description="Train Ackermann Robot RL Agent (batched B200 environment)")
    # ---- the reference's flags (src/rl/train.py:231-259), same names, defaults and meaning ---------------------------------
    ap.add_argument("--algo", default="random", choices=["random", "ppo", "sac", "td3"], help="RL algorithm to use")
    ap.add_argument("--episodes", type=int, default=1000, help="Number of episodes (for random)")
    ap.add_argument("--timesteps", type=int, default=100000, help="Number of timesteps (for PPO)")
    ap.add_argument("--render", action="store_true", help="Render environment during training (not available: no viewer on the GPU path)")
    ap.add_argument("--max-velocity", type=float, default=1.0, help="Maximum linear velocity (m/s)")
    ap.add_argument("--goal-threshold", type=float, default=0.5, help="Goal distance threshold (m)")
    ap.add_argument("--maze", default=None, choices=[None, "umaze", "open", "medium", "large"],
                    help="Use a PointMaze scene (layout chosen by --maze-id, as in the reference)")
    ap.add_argument("--maze-id", default="PointMaze_UMaze-v3", help="Gymnasium Robotics maze environment ID")
    ap.add_argument("--learning-rate", type=float, default=None, help="Learning rate for the algorithm (PPO default 3e-4)")
    ap.add_argument("--save-freq", type=int, default=10000, help="Frequency (timesteps) to save model checkpoints (needs --save-path)")
    ap.add_argument("--eval-freq", type=int, default=10000, help="Frequency (timesteps) to evaluate the model")
    ap.add_argument("--eval-episodes", type=int, default=10, help="Number of episodes for the final evaluation (0 = skip)")
    # ---- batched-environment additions --------------------------------------------------------------------------------------
    ap.add_argument("--save-path", default="", help="prefix for <prefix>_<N>_steps.zip / <prefix>_final.zip (SB3 layout, see sb3_io)")
    ap.add_argument("--num-envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--frame-skip", type=int, default=1)
    ap.add_argument("--n-steps", type=int, default=16)
    ap.add_argument("--n-epochs", type=int, default=10)
    ap.add_argument("--minibatches", type=int, default=4)
    ap.add_argument("--model", default="v2", help="v2 | scene | maze:umaze|open|medium|large")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-eval-steps", type=int, default=1000, help="upper bound of env steps per evaluation (one full episode)")
    a = ap.parse_args(argv)

    if a.algo in ("sac", "td3"):
        raise SystemExit(f"--algo {a.algo}: only the PPO path of the reference is built (SAC / TD3 are out of scope, DESIGN.md section 8)")
    if a.render:
        raise SystemExit("--render: the viewer is outside the hot path; there is no rendering on the GPU path")
    if a.maze is not None:                                  # train.py:262-270: --maze switches to the maze env selected by --maze-id
        from .compiler.maze import MAZE_ENV_IDS
        if a.maze_id not in MAZE_ENV_IDS:
            raise SystemExit(f"unknown --maze-id {a.maze_id}; known: {sorted(MAZE_ENV_IDS)}")
        a.model = "maze:" + MAZE_ENV_IDS[a.maze_id]
        print(f"Using Gymnasium Robotics maze: {a.maze_id}")
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    torch.backends.cuda.matmul.allow_tf32 = True      # learner GEMMs (18 757-parameter MLP) on TF32 tensor cores; the simulator is unaffected
    torch.backends.cudnn.allow_tf32 = True
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # one seed for the whole job, streams keyed by the global environment id: the result does not depend on the number of GPUs
    env = BatchedAckermannEnv(a.num_envs, device=dev, frame_skip=a.frame_skip, seed=a.seed, env_id_base=rank * a.num_envs, model=a.model,
                              max_linear_velocity=a.max_velocity, goal_distance_threshold=a.goal_threshold)
    if a.algo == "random":
        env.reset()
        # the reference runs `--episodes` episodes of one environment (train.py:189-227); here the batch is stepped with random
        # actions until that many episodes have finished over all ranks (checked every 100 steps: no per-step host sync)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        done_eps, steps = 0, 0
        while done_eps < a.episodes:
            for _ in range(100):
                env.step(None)
            steps += 100
            done_eps = reduce_stats(env.stats(), device=dev)["episodes"]
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        st = reduce_stats(env.stats(), device=dev)
        if rank == 0:
            print(json.dumps({"algo": "random", "env_steps": st["env_steps"], "env_steps_per_s": st["env_steps"] / dt, "episodes": st["episodes"],
                              "ep_rew_mean": st["return_sum"] / max(1, st["episodes"])}))
    else:
        from .sb3_io import evaluate_agent, save_sb3_policy
        cfg = PPOConfig(n_steps=a.n_steps, n_epochs=a.n_epochs, minibatches=a.minibatches,
                        learning_rate=3e-4 if a.learning_rate is None else a.learning_rate)
        tr = PPOTrainer(env, cfg, seed=a.seed)
        hyper = dict(learning_rate=cfg.learning_rate, n_steps=cfg.n_steps, n_epochs=cfg.n_epochs, gamma=cfg.gamma, gae_lambda=cfg.gae_lambda,
                     clip_range=cfg.clip_range, ent_coef=cfg.ent_coef, vf_coef=cfg.vf_coef, max_grad_norm=cfg.max_grad_norm,
                     batch_size=cfg.n_steps * a.num_envs // cfg.minibatches)
        next_save = [a.save_freq]
        next_eval = [a.eval_freq]

        def log(d):
            if rank == 0:
                print(json.dumps(d), flush=True)
            if a.eval_freq > 0 and a.eval_episodes > 0 and d["timesteps"] >= next_eval[0]:      # EvalCallback (train.py:149-158)
                ev = evaluate_agent(eval_env, tr.policy, n_steps=a.max_eval_steps, n_episodes=a.eval_episodes)   # collective: every rank
                if rank == 0:
                    print(json.dumps({**ev, "evaluation": True, "timesteps": d["timesteps"]}), flush=True)
                while next_eval[0] <= d["timesteps"]:
                    next_eval[0] += max(1, a.eval_freq)
            if a.save_path and rank == 0 and d["timesteps"] >= next_save[0]:     # CheckpointCallback (train.py:140-144)
                save_sb3_policy(tr.policy, f"{a.save_path}_{d['timesteps']}_steps.zip", tr.opt, d["timesteps"], hyper=hyper,
                                flat_order=getattr(tr.graphed, "order", None), total_timesteps=a.timesteps)
                while next_save[0] <= d["timesteps"]:
                    next_save[0] += max(1, a.save_freq)

        # a separate, smaller evaluation batch (the reference evaluates on the training env; with auto-reset batches a second handle
        # keeps the training episodes undisturbed -- handles on one device are independent)
        eval_env = BatchedAckermannEnv(min(a.num_envs, 1024), device=dev, frame_skip=a.frame_skip, seed=a.seed + 1, model=a.model,
                                       max_linear_velocity=a.max_velocity, goal_distance_threshold=a.goal_threshold,
                                       env_id_base=rank * min(a.num_envs, 1024))
        tr.train(a.timesteps, log=log, log_all_ranks=True)
        if a.save_path and rank == 0:                                               # model.save (train.py:182-183)
            save_sb3_policy(tr.policy, f"{a.save_path}_final.zip", tr.opt, tr.num_timesteps, hyper=hyper,
                            flat_order=getattr(tr.graphed, "order", None), total_timesteps=a.timesteps)
        if a.eval_episodes > 0:                                                     # evaluate_agent (train.py:303-309)
            ev = evaluate_agent(eval_env, tr.policy, n_steps=a.max_eval_steps, n_episodes=a.eval_episodes)
            ev = {**ev, "evaluation": True, "final": True}
            if rank == 0:
                print(json.dumps(ev), flush=True)
        eval_env.close()
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
