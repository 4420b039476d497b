#!/usr/bin/env python
"""Device time of the pieces of one PPO rollout (GPU box): the full collect() against 16 x {policy forward}, 16 x {env step},
16 x {time-limit bootstrap} issued alone (CUDA events, back to back)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from mujoco_playground_b200 import BatchedAckermannEnv  # noqa: E402
from mujoco_playground_b200.ppo import PPOConfig, PPOTrainer  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = BatchedAckermannEnv(n, seed=1)
tr = PPOTrainer(env, PPOConfig(n_steps=16), seed=0)
for _ in range(3):
    tr.collect(); tr.update()
b, f, T = tr.buf, tr.graphed, tr.cfg.n_steps


def timed(fn, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return round(best, 4)


def acts():
    for t in range(T):
        f.act(b["obs"][t], b["act"][t], b["logp"][t], b["val"][t], 1, t)


def steps():
    for t in range(T):
        env.step(b["act"][t], obs_out=b["obs"][t + 1] if t + 1 < T else tr.obs)


def boots():
    info = {"terminal_observation": env.terminal_obs}
    for t in range(T):
        f.bootstrap(info["terminal_observation"], env.terminated, env.truncated, env.reward, b["rew"][t], b["done"][t])


res = {"envs": n, "collect_ms": timed(tr.collect), "act_x16_ms": timed(acts), "step_x16_ms": timed(steps), "bootstrap_x16_ms": timed(boots)}
res["sum_ms"] = round(res["act_x16_ms"] + res["step_x16_ms"] + res["bootstrap_x16_ms"], 4)
print(json.dumps(res))
