#!/usr/bin/env python
"""Benchmark of the Ackermann env-step hot path (see BASELINE.json / SURVEY.md 8d).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the reference loop

A "step" is one env.step() of every environment of the batch = frame_skip physics substeps of 2 ms.
Workloads (BASELINE.json configs):
  N = 1 : configs[1] "ackermann flat-floor, 4096 batched envs on 1 B200, random actions, frame_skip=4"
  N > 1 : configs[3] "ackermann flat-floor, 131072 envs per GPU sharded across 2/4/8 B200" (no data-path collective)
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_F32 = 650.0   # algorithmic HBM bytes per env-step, fp32 state (SURVEY.md 8d / DESIGN.md)
ALGO_BYTES_F64 = 962.0


def measured_traffic(dtype, n_envs, fs):
    """(DRAM bytes, note, warp instructions) per launch of the step kernel from the committed ncu capture of the same
    configuration, or Nones."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        d = json.load(open(p)).get(f"{'f32' if dtype == 'float32' else 'f64'}_{n_envs}_fs{fs}")
        return (d["dram_bytes_read"] + d["dram_bytes_write"], d["note"], d.get("warp_instructions")) if d else (None, None, None)
    except Exception:
        return None, None, None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks and throttle reasons while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def _run_nvml(self):
        """Fast path: NVML through nvidia_ml_py (about 1 ms per sample instead of ~40 ms per nvidia-smi process)."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = int(get_reasons(h))
            self.rows.append([str(sm), str(mx)] + ["Active" if r & bits[k] else "Not Active" for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            self.stop_flag.wait(0.005)
        nv.nvmlShutdown()

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the reference's loop (controller + mj_step) restated in C (oracle/), one process per core
# ----------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init():
    """Pool initialiser: one oracle simulator per worker process, warmed up once."""
    from mujoco_playground_b200.models import load_model
    from oracle.oracle import OracleSim
    M = load_model("v2")
    sim = OracleSim(M)
    sq = M["qpos0"].copy()
    sq[0:3] = [0, 0, 0.1]
    sim.rollout(200, 4, 1000, 12345, sq)       # warm-up
    _W["sim"], _W["sq"] = sim, sq


def _cpu_worker(args):
    seed, n_steps, frame_skip = args
    if "sim" not in _W:
        _cpu_init()
    t0 = time.perf_counter()
    _W["sim"].rollout(n_steps, frame_skip, 1000, seed, _W["sq"])
    return time.perf_counter() - t0


class CpuArm:
    """The reference's loop (controller + mj_step, restated in C under oracle/) on every host core: one persistent worker
    process per core, each stepping its own environment."""

    def __init__(self, procs: int):
        import multiprocessing as mp
        from oracle import oracle as _o
        _o.build()
        self.procs = procs
        self.pool = mp.get_context("spawn").Pool(procs, initializer=_cpu_init)
        self.calls = 0

    def rollout(self, n_steps_per_proc: int, frame_skip: int):
        """Returns (env-steps/s aggregated over the processes, seconds of the slowest worker)."""
        self.calls += 1
        times = self.pool.map(_cpu_worker, [(self.calls * 1000 + s, n_steps_per_proc, frame_skip) for s in range(self.procs)], chunksize=1)
        tmax = max(times)
        # aggregate rate = sum of the workers' own rates (the cores run independent environments); with short samples the
        # slowest-worker convention would under-state the CPU arm
        return sum(n_steps_per_proc / t for t in times), tmax

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_rollout(n_steps_per_proc: int, frame_skip: int, procs: int):
    arm = CpuArm(procs)
    try:
        return arm.rollout(n_steps_per_proc, frame_skip)
    finally:
        arm.close()


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    fs = args.frame_skip
    # bounded sample: the whole --steps K --warmup W run spends about args.cpu_steps physics substeps per process in total
    # (~60 s at ~55 k substeps/s per core), whatever K is; at least 100 env-steps per step
    per = max(200, args.cpu_steps // fs // max(1, args.steps + args.warmup))
    arm = CpuArm(cores)
    vals = []
    for _ in range(args.warmup):
        arm.rollout(per, fs)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, _t = arm.rollout(per, fs)
        vals.append(v)
    wall = time.perf_counter() - t_all
    arm.close()
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "frame_skip": fs, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{per} env-steps x {cores} processes per step, {args.steps} steps (C restatement of controller + mj_step; "
                                   "real mujoco is not installable in this image)"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(args, world):
    if getattr(args, "workload", "flat") == "scene":
        return (f"configs[2]: ackermann obstacle scene (ackermann_maze_flat.xml), {args.envs or 65536} envs per GPU, AckermannController, "
                f"spawn yaw U(-pi,pi) + xy jitter 0.12 m, random actions, frame_skip={args.frame_skip}")
    if args.envs:
        return f"ackermann flat-floor, {args.envs} envs per GPU, random actions, frame_skip={args.frame_skip}"
    if world == 1:
        return "configs[1]: ackermann flat-floor, 4096 batched envs on 1 B200, random actions, frame_skip=4"
    return f"configs[3]: ackermann flat-floor, 131072 envs per GPU sharded across {world} B200, frame_skip=4"


# ----------------------------------------------------------------------------------------------------------
def run_cuda(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from mujoco_playground_b200 import BatchedAckermannEnv

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    scene = args.workload == "scene"
    n_envs = args.envs or (65536 if scene else (4096 if world == 1 else 131072))
    fs = args.frame_skip
    kw = dict(model="scene", spawn_yaw_range=3.141592653589793, spawn_xy_jitter=0.12) if scene else {}
    env = BatchedAckermannEnv(n_envs, device=dev, frame_skip=fs, dtype=args.dtype, seed=1234 + rank, auto_reset=True,
                              lanes_per_env=args.lanes, **kw)
    env.reset()
    # steady state of a long rollout: episode phases staggered uniformly (resets spread over time instead of all
    # environments resetting in the same step), and the robots already landed on their wheels
    import numpy as np
    env.set_episode(step_count=np.random.default_rng(rank).integers(0, 1000, n_envs).astype(np.int32))
    for _ in range((400 + fs - 1) // fs):
        env.step(None)
    env.stats_reset()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- kernel-resident throughput: synthetic actions generated in the kernel, state resident in HBM --------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        env.step(None)
    barrier()
    launches0 = env.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for s0, s1 in evs:
        flush.zero_()                     # evict state from L2 between timed iterations (not timed)
        s0.record()
        env.step(None)
        s1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = env.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * n_envs * args.steps / (total_ms * 1e-3)
    stats = env.stats()

    # ---- end to end through the C ABI with HOST buffers (H2D actions, D2H obs/reward/flags inside the timed region) ---
    h_act = torch.empty((n_envs, 2), dtype=torch.float32).uniform_(-1, 1).pin_memory()
    h_obs = torch.empty((n_envs, env.obs_dim), dtype=torch.float32).pin_memory()
    h_rew = torch.empty((n_envs,), dtype=torch.float32).pin_memory()
    h_term = torch.empty((n_envs,), dtype=torch.uint8).pin_memory()
    h_trunc = torch.empty((n_envs,), dtype=torch.uint8).pin_memory()
    for _ in range(args.warmup):
        env.step_host(h_act, h_obs, h_rew, h_term, h_trunc)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        env.step_host(h_act, h_obs, h_rew, h_term, h_trunc)
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * n_envs * args.steps / float(e2e_s.item())
    h2d = n_envs * 2 * 4
    d2h = n_envs * (env.obs_dim * 4 + 4 + 1 + 1)
    clocks = sampler.summary() if sampler else None

    if rank == 0:
        peak, how = measured_peak_gbs()
        algo = ALGO_BYTES_F32 if args.dtype == "float32" else ALGO_BYTES_F64
        if scene:   # 36-beam observation (43 floats): 506 B per env-step in fp32
            algo -= (79 - env.obs_dim) * 4
        avg_launch_s = (sum(step_ms) / len(step_ms)) * 1e-3
        achieved = algo * n_envs / avg_launch_s / 1e9
        traffic, traffic_note, warp_inst = measured_traffic(args.dtype, n_envs, fs)
        # the bound that actually binds: warp-instruction issue.  Peak = SMs x 4 schedulers x SM clock.
        issue = None
        if warp_inst:
            sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
            peak_issue = 148 * 4 * sm_clock * 1e6
            issue = {"bound": "issue", "unit": "warp-instr/s", "achieved": warp_inst / avg_launch_s, "peak": peak_issue,
                     "frac": warp_inst / avg_launch_s / peak_issue, "warp_instructions_per_launch": warp_inst,
                     "note": "instruction count from the committed ncu capture of this configuration, time measured live"}
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "envs_per_gpu": n_envs, "frame_skip": fs, "lanes_per_env": args.lanes,
                       "l2": "flushed between timed iterations (256 MiB memset, untimed)", "physics_substeps_per_s": value * fs,
                       "scaling_note": ("N=1 runs configs[1] (4096 envs); N>1 runs configs[3] (131072 envs per GPU, weak scaling). The "
                                        "single-GPU figure at the configs[3] batch is this line's aux.value at N=1") if not args.envs and not scene else None},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                         "peak_source": how, "kernel": "step_kernel", "algorithmic_bytes_per_env_step": algo, "algorithmic_bytes_per_launch": algo * n_envs,
                         "note": "compute/latency bound kernel: see DESIGN.md (HBM fraction is small by construction)", "issue": issue},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "solver": {"mean_newton_iters_per_env_step": stats["solver_iters"] / max(1, stats["env_steps"]),
                       "episodes": stats["episodes"], "unsupported_contact_steps": stats["unsupported"],
                       "mean_ncon_last_substep": stats["contacts_sum"] / max(1, stats["env_steps"]),
                       "obstacle_contact_step_fraction": stats["obstacle_steps"] / max(1, stats["env_steps"])},
            "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.envs and not args.no_aux and not scene:
            # the same kernel at the per-GPU batch of configs[3] (131072 envs), for context next to the 4096-env headline
            env.close()
            big = BatchedAckermannEnv(131072, device=dev, frame_skip=fs, dtype=args.dtype, seed=99, auto_reset=True, lanes_per_env=args.lanes)
            big.reset()
            big.set_episode(step_count=np.random.default_rng(7).integers(0, 1000, 131072).astype(np.int32))
            for _ in range((400 + fs - 1) // fs):
                big.step(None)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nb = 30
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(nb):
                big.step(None)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / nb
            line["aux"] = {"workload": f"131072 envs on 1 B200, frame_skip={fs}, back-to-back launches (state 85 MB < L2)",
                           "value": 131072 / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms,
                           "roofline_frac": algo * 131072 / (ms * 1e-3) / 1e9 / peak}
            big.close()
        if world == 1 and not args.no_cpu_baseline and not scene:
            cores = os.cpu_count() or 1
            per = max(50, args.cpu_steps // fs)
            v, tmax = cpu_rollout(per, fs, cores)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{per} env-steps (frame_skip={fs}) on each of {cores} processes, {tmax:.1f} s; C restatement of "
                                              "controller + mj_step (real mujoco not installable in this image)"}
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def run_ppo(args, rank, local_rank, world):
    """configs[4]: the PPO training loop (rollout on the batched env + fused learner + NCCL gradient all-reduce).
    A "step" is one PPO iteration = n_steps env steps of every environment followed by the update."""
    import torch
    import torch.distributed as dist
    from mujoco_playground_b200 import BatchedAckermannEnv
    from mujoco_playground_b200.ppo import PPOConfig, PPOTrainer
    from mujoco_playground_b200.shard import rank_seed
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_envs = args.envs or 65536
    env = BatchedAckermannEnv(n_envs, device=dev, frame_skip=1, seed=rank_seed(1234, rank))
    cfg = PPOConfig(n_steps=16)
    tr = PPOTrainer(env, cfg, seed=0)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        tr.collect(); tr.update()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    steps = min(args.steps, 50)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    roll = upd = 0.0
    e0.record()
    for _ in range(steps):
        roll += tr.collect()
        upd += tr.update()["update_s"]
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.summary() if sampler else None
    if rank == 0:
        per_iter = cfg.n_steps * n_envs * world
        print(json.dumps({
            "metric": "env-steps/sec", "value": per_iter * steps / (total_ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 simulator, TF32 tensor-core learner (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": f"configs[4]: PPO training loop, {n_envs} envs per GPU x {world} GPU, n_steps=16, 10 epochs x 4 minibatches, "
                                   "frame_skip=1, NCCL gradient all-reduce", "envs_per_gpu": n_envs},
            "split": {"rollout_ms_per_iteration": 1e3 * roll / steps, "update_ms_per_iteration": 1e3 * upd / steps,
                      "rollout_env_steps_per_s": per_iter * steps / roll},
            "clocks": clocks}), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="override environments per GPU")
    ap.add_argument("--frame-skip", type=int, default=4)
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--lanes", type=int, default=0, help="lanes per env: 0 = auto (1 for >= 32768 envs, else 4)")
    ap.add_argument("--no-aux", action="store_true", help="skip the extra large-batch measurement at N=1")
    ap.add_argument("--cpu-steps", type=int, default=800000, help="physics substeps per CPU process for the CPU arm sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="flat", choices=["flat", "scene", "ppo"],
                    help="flat = configs[1]/[3] (default), scene = configs[2], ppo = configs[4] (training loop)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "ppo":
        run_ppo(args, rank, local_rank, world)
    else:
        run_cuda(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
