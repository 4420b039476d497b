#!/usr/bin/env python
"""Tuning helper (GPU box): times back-to-back env.step(None) launches of one library variant.
    ACKB_LIB=build/variants/x.so python tools/gpu/time_step.py --envs 131072 --lanes 1 --fs 4 [--set ls_fast_iters=2 ...]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import torch

from mujoco_playground_b200 import BatchedAckermannEnv
from mujoco_playground_b200.compiler.constants import consts_layout
from mujoco_playground_b200 import _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=131072)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--fs", type=int, default=4)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--model", default="v2")
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--set", nargs="*", default=[])
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    kw = {}
    if a.model == "scene":
        kw = dict(spawn_yaw_range=3.14159, spawn_xy_jitter=0.3)
    env = BatchedAckermannEnv(a.envs, frame_skip=a.fs, dtype=a.dtype, seed=99, auto_reset=True, lanes_per_env=a.lanes, model=a.model, **kw)
    if a.set:
        lay = consts_layout()
        for kv in a.set:
            k, v = kv.split("=")
            env.consts[lay[k][0]] = float(v)
        # re-create the handle with the patched constants
        import ctypes
        env.close()
        env.h = ctypes.c_void_p()
        _lib.check(env.L.ackb_create(env.consts.ctypes.data_as(ctypes.c_void_p), len(env.consts), env.num_envs, 0,
                                     0 if a.dtype == "float32" else 1, 99, a.lanes, ctypes.byref(env.h)))
    env.reset()
    env.set_episode(step_count=np.random.default_rng(7).integers(0, 1000, a.envs).astype(np.int32))
    for _ in range((400 + a.fs - 1) // a.fs):
        env.step(None)
    env.stats_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            env.step(None)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.iters)
    st = env.stats()
    print(json.dumps({"tag": a.tag, "lib": os.path.basename(_lib.LIB_PATH), "envs": a.envs, "lanes": a.lanes, "fs": a.fs, "set": a.set,
                      "ms": round(best, 4), "Msteps_s": round(a.envs / best / 1e3, 2),
                      "iters_per_env_step": round(st["solver_iters"] / max(1, st["env_steps"]), 3), "unsupported": st["unsupported"],
                      "mean_ncon": round(st["contacts_sum"] / max(1, st["env_steps"]), 3),
                      "obstacle_frac": round(st["obstacle_steps"] / max(1, st["env_steps"]), 5)}), flush=True)
    env.close()


if __name__ == "__main__":
    main()
