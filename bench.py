#!/usr/bin/env python
"""Benchmark of the Ackermann env-step hot path (see BASELINE.json / SURVEY.md 8d).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the reference loop

A "step" is one env.step() of every environment of the batch = frame_skip physics substeps of 2 ms.

Headline (every N): configs[3] "ackermann flat-floor, 131072 envs per GPU" (weak scaling: the N=1 point of the
"env-steps/sec at 1/2/4/8 B200" series is the same per-GPU batch, no data-path collective).
Sub-records of the same JSON line (`sub`), each with its own roofline / e2e:
  configs[1]  4096 envs on 1 B200, frame_skip 4                       (N = 1 only)
  configs[2]  obstacle scene, 65536 envs, frame_skip 1 and 4          (N = 1 only)
  configs[4]  PPO training loop, 65536 envs per GPU, NCCL all-reduce  (every N; TF32/tcgen05 learner, fp32 learner at N = 1)
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

ALGO_BYTES_F32 = 650.0   # algorithmic HBM bytes per env-step, fp32 state, 79-float observation (SURVEY.md 8d / DESIGN.md)
ALGO_BYTES_F64 = 962.0


def measured_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the same configuration
    (profiles/traffic.json: key -> {dram_bytes_read, dram_bytes_write, note, profile}), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key)
        return (d["dram_bytes_read"] + d["dram_bytes_write"], f"{d['note']} [{d['profile']}]") if d else (None, None)
    except Exception:
        return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1628.1)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1628.1, "fallback (B200_PROFILING.md)"


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
    sim = OracleSim(M, fast=True)              # -O3 -march=native build of the same C source, compiled on this host
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
        _o.build_fast()
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
        "config": {"workload": f"configs[3]: ackermann flat-floor, random actions, frame_skip={fs} (the CPU arm steps one environment per host core; the per-GPU batch size does not apply)", "frame_skip": fs, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{per} env-steps x {cores} processes per step, {args.steps} steps (C restatement of controller + mj_step, -O3 -march=native; "
                                   "this repo's own CPU port -- real mujoco is not installable in this image)"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



def algo_bytes(dtype, obs_dim):
    """Algorithmic HBM bytes per env-step: state 37 reals r/w + goal/ref 4 reals read + counter r/w + action in, obs/reward/flags out."""
    return (ALGO_BYTES_F32 if dtype == "float32" else ALGO_BYTES_F64) - (79 - obs_dim) * 4


def make_barrier(world, dev):
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
    return barrier


def max_over_ranks(x, world, dev):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measure_env(name, traffic_key, *, dev, rank, world, local_rank, n_envs, fs, dtype, lanes, steps, warmup, flush, env_kw=None, e2e=True,
                env_id_base=0):
    """One env-step workload: device-timed throughput with the state resident in HBM (L2 flushed between timed launches) and the
    end-to-end figure through ackb_step_host with pinned HOST buffers.  Returns the record (rank 0) or None."""
    import numpy as np
    import torch
    from mujoco_playground_b200 import BatchedAckermannEnv
    barrier = make_barrier(world, dev)
    env = BatchedAckermannEnv(n_envs, device=dev, frame_skip=fs, dtype=dtype, seed=1234, auto_reset=True, lanes_per_env=lanes,
                              env_id_base=env_id_base, **(env_kw or {}))
    env.reset()
    # steady state of a long rollout: episode phases staggered uniformly (resets spread over time), robots landed on their wheels
    env.set_episode(step_count=np.random.default_rng(rank).integers(0, 1000, n_envs).astype(np.int32))
    for _ in range((400 + fs - 1) // fs):
        env.step(None)
    env.stats_reset()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(warmup):
        env.step(None)
    barrier()
    launches0 = env.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall = time.perf_counter()
    for s0, s1 in evs:
        flush.zero_()                     # evict the state from L2 between timed iterations (not timed)
        s0.record()
        env.step(None)
        s1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = env.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = max_over_ranks(sum(step_ms), world, dev)
    value = world * n_envs * steps / (total_ms * 1e-3)
    stats = env.stats()

    e2e_rec = None
    if e2e:
        # end to end through the C ABI with HOST buffers (H2D actions, D2H obs / reward / flags inside the timed region)
        h_act = torch.empty((n_envs, 2), dtype=torch.float32).uniform_(-1, 1).pin_memory()
        h_obs = torch.empty((n_envs, env.obs_dim), dtype=torch.float32).pin_memory()
        h_rew = torch.empty((n_envs,), dtype=torch.float32).pin_memory()
        h_term = torch.empty((n_envs,), dtype=torch.uint8).pin_memory()
        h_trunc = torch.empty((n_envs,), dtype=torch.uint8).pin_memory()
        for _ in range(warmup):
            env.step_host(h_act, h_obs, h_rew, h_term, h_trunc)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            env.step_host(h_act, h_obs, h_rew, h_term, h_trunc)
        torch.cuda.synchronize(dev)
        e2e_s = max_over_ranks(time.perf_counter() - t0, world, dev)
        d2h = n_envs * (env.obs_dim * 4 + 4 + 1 + 1)
        e2e_rec = {"value": world * n_envs * steps / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": n_envs * 2 * 4, "d2h_bytes_per_step": d2h,
                   "host_link_GBps_per_gpu": (d2h + n_envs * 8) * steps / e2e_s / 1e9,
                   "path": "ackb_step_host, pinned host buffers mapped into the kernel (zero-copy stores over PCIe)"}
    clocks = sampler.summary() if sampler else None
    obs_dim = env.obs_dim
    env.close()
    if rank != 0:
        return None
    peak, _, how = measured_peaks()
    algo = algo_bytes(dtype, obs_dim)
    avg_launch_s = (sum(step_ms) / len(step_ms)) * 1e-3
    achieved = algo * n_envs / avg_launch_s / 1e9
    traffic, traffic_note = measured_traffic(traffic_key)
    return {
        "workload": name, "value": value, "unit": "env-steps/s", "ms_per_step": total_ms / steps, "envs_per_gpu": n_envs, "frame_skip": fs,
        "physics_substeps_per_s": value * fs, "dtype": "f32" if dtype == "float32" else "f64",
        "l2": "flushed between timed iterations (256 MiB memset, untimed)",
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_note": traffic_note, "peak_source": how, "kernel": "step_kernel", "algorithmic_bytes_per_env_step": algo,
                     "algorithmic_bytes_per_launch": algo * n_envs,
                     "note": "issue/latency-bound kernel (about 3e4 instructions per environment and substep for 650 B): the HBM fraction is "
                             "small by construction, DESIGN.md section 4; the binding counters are in the ncu summaries under profiles/"},
        "e2e": e2e_rec, "gpu_launches": int(launches), "clocks": clocks,
        "solver": {"mean_newton_iters_per_env_step": stats["solver_iters"] / max(1, stats["env_steps"]), "episodes": stats["episodes"],
                   "unsupported_contact_steps": stats["unsupported"], "bad_state_resets": stats["bad_state"],
                   "mean_ncon_last_substep": stats["contacts_sum"] / max(1, stats["env_steps"]),
                   "obstacle_contact_step_fraction": stats["obstacle_steps"] / max(1, stats["env_steps"])},
        "wall_s_timed_region": t_wall,
    }


PPO_FLOP_PER_SAMPLE = 2 * 2 * (79 * 64 + 64 * 64) * 3 + 2 * 3 * 64 * 3   # both MLPs: forward + dX + dW (2 flop per MAC), heads


def measure_ppo(*, dev, rank, world, local_rank, n_envs, iters, warmup, mode, env_id_base=0):
    """configs[4]: PPO iterations (n_steps env steps of every environment + the update) with the learner arithmetic `mode`
    ("tf32" tensor cores or "fp32" CUDA cores).  Device-timed with CUDA events; e2e = wall clock of the same loop as a user
    runs it (PPOTrainer.collect / update, statistics read back to the host every iteration)."""
    import torch
    from mujoco_playground_b200 import BatchedAckermannEnv
    from mujoco_playground_b200.ppo import PPOConfig, PPOTrainer
    barrier = make_barrier(world, dev)
    env = BatchedAckermannEnv(n_envs, device=dev, frame_skip=1, seed=1234, env_id_base=env_id_base)
    cfg = PPOConfig(n_steps=16)
    tr = PPOTrainer(env, cfg, seed=0, learner_mode=mode)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(max(warmup, 3)):
        tr.collect(); tr.update()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    roll = upd = 0.0
    d2h = 0
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        roll += tr.collect()
        st = tr.update()
        upd += st["update_s"]
        d2h += 5 * 4
    e1.record()
    barrier()
    wall = max_over_ranks(time.perf_counter() - t0, world, dev)
    total_ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    clocks = sampler.summary() if sampler else None
    gk = tr.grad_kernel_seconds(reps=5) if hasattr(tr, "grad_kernel_seconds") else None
    env.close()
    if rank != 0:
        return None
    per_iter = cfg.n_steps * n_envs * world
    _, tf_peak, how = measured_peaks()
    rec = {
        "workload": f"configs[4]: PPO training loop, {n_envs} envs per GPU x {world} GPU, n_steps={cfg.n_steps}, {cfg.n_epochs} epochs x "
                    f"{cfg.minibatches} minibatches, frame_skip=1, NCCL gradient all-reduce, learner arithmetic {mode}",
        "value": per_iter * iters / (total_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": total_ms / iters, "envs_per_gpu": n_envs,
        "iterations": iters, "learner": mode,
        "split": {"rollout_ms_per_iteration": 1e3 * roll / iters, "update_ms_per_iteration": 1e3 * upd / iters,
                  "rollout_env_steps_per_s": per_iter * iters / roll},
        "e2e": {"value": per_iter * iters / wall, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h // max(1, iters),
                "path": "PPOTrainer.collect() + update(): the rollout never leaves the device; the update's diagnostics are read back"},
        "clocks": clocks,
    }
    if gk:
        mb = cfg.n_steps * n_envs // cfg.minibatches
        fl = PPO_FLOP_PER_SAMPLE * mb
        peak = tf_peak * 0.5 if mode != "fp32" else 74.5      # TF32 dense = half the measured bf16 rate; fp32 CUDA-core peak 148 SM x 128 x 2 x 1.965 GHz
        rec["roofline"] = {"bound": "tensor" if mode != "fp32" else "fp32-cuda-core", "achieved": fl / gk / 1e12, "peak": peak, "unit": "TFLOP/s",
                           "frac": fl / gk / 1e12 / peak, "traffic": measured_traffic(f"ppo_grad_{mode}")[0], "kernel": "ppo_grad_kernel",
                           "flop_per_sample": PPO_FLOP_PER_SAMPLE, "samples_per_launch": mb, "ms_per_launch": gk * 1e3, "peak_source": how,
                           "algorithmic_bytes_per_launch": mb * 336,
                           "note": "TF32 peak taken as half of the measured dense bf16 rate (no TF32 entry in MEASURED_PEAKS.json)"}
    return rec


def run_cuda(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fs = args.frame_skip
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    common = dict(dev=dev, rank=rank, world=world, local_rank=local_rank, dtype=args.dtype, lanes=args.lanes, steps=args.steps,
                  warmup=args.warmup, flush=flush)
    scene_kw = dict(model="scene", spawn_yaw_range=3.141592653589793, spawn_xy_jitter=0.12)
    sub = {}
    if args.workload == "ppo":
        n_envs = args.envs or 65536
        head = measure_ppo(dev=dev, rank=rank, world=world, local_rank=local_rank, n_envs=n_envs, iters=min(args.steps, 50), warmup=args.warmup,
                           mode=args.learner, env_id_base=rank * n_envs)
    elif args.workload == "scene":
        n_envs = args.envs or 65536
        head = measure_env(f"configs[2]: ackermann obstacle scene (ackermann_maze_flat.xml), {n_envs} envs per GPU, AckermannController, spawn yaw "
                           f"U(-pi,pi) + xy jitter 0.12 m, random actions, frame_skip={fs}", f"scene_{n_envs}_fs{fs}", n_envs=n_envs, fs=fs,
                           env_kw=scene_kw, env_id_base=rank * n_envs, **common)
    else:
        n_envs = args.envs or 131072
        nm = (f"configs[3]: ackermann flat-floor, 131072 envs per GPU sharded across {world} B200, random actions, frame_skip={fs}" if not args.envs
              else f"ackermann flat-floor, {n_envs} envs per GPU, random actions, frame_skip={fs}")
        head = measure_env(nm, f"flat_{n_envs}_fs{fs}", n_envs=n_envs, fs=fs, env_id_base=rank * n_envs, **common)
        if not args.no_sub and not args.envs:
            if world == 1:
                sub["configs[1]"] = measure_env("configs[1]: ackermann flat-floor, 4096 batched envs on 1 B200, random actions, frame_skip=4",
                                                "flat_4096_fs4", n_envs=4096, fs=4, **common)
                for f2 in (1, 4):
                    sub[f"configs[2] fs={f2}"] = measure_env(
                        f"configs[2]: ackermann obstacle scene (ackermann_maze_flat.xml), 65536 envs on 1 B200, AckermannController, spawn yaw "
                        f"U(-pi,pi) + xy jitter 0.12 m, random actions, frame_skip={f2}", f"scene_65536_fs{f2}", n_envs=65536, fs=f2,
                        env_kw=scene_kw, **common)
            pk = dict(dev=dev, rank=rank, world=world, local_rank=local_rank, n_envs=65536, iters=10, warmup=3)
            sub["configs[4]"] = measure_ppo(mode="tcgen05", env_id_base=rank * 65536, **pk)
            if world == 1:
                sub["configs[4] mma.sync tf32 learner"] = measure_ppo(mode="tf32", **pk)
                sub["configs[4] fp32 learner"] = measure_ppo(mode="fp32", **pk)

    if rank == 0:
        line = {
            "metric": "env-steps/sec", "value": head["value"], "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": head.get("dtype", "f32"), "data": "synthetic",
            "config": {"workload": head["workload"], "envs_per_gpu": head["envs_per_gpu"], "frame_skip": head.get("frame_skip", 1),
                       "lanes_per_env": args.lanes, "l2": head.get("l2"), "physics_substeps_per_s": head.get("physics_substeps_per_s"),
                       "rng": "Philox keyed by (seed, global environment id): the sharded batch equals the single-GPU batch"},
            "roofline": head.get("roofline"), "e2e": head.get("e2e"), "gpu_launches": head.get("gpu_launches"), "clocks": head.get("clocks"),
            "solver": head.get("solver"), "wall_s_timed_region": head.get("wall_s_timed_region"),
        }
        if "split" in head:
            line["split"] = head["split"]
        if sub:
            line["sub"] = {k: v for k, v in sub.items() if v is not None}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            per = max(50, args.cpu_steps // fs)
            v, tmax = cpu_rollout(per, fs, cores)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{per} env-steps (frame_skip={fs}) on each of {cores} processes, {tmax:.1f} s; C restatement of "
                                              "controller + mj_step, -O3 -march=native (real mujoco is not installable in this image, so this is "
                                              "this repo's own CPU port, not mujoco.mj_step)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="override environments per GPU")
    ap.add_argument("--frame-skip", type=int, default=4)
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--lanes", type=int, default=0, help="lanes per env: 0 = auto (1 for >= 32768 envs, else 4)")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (configs[1], [2], [4])")
    ap.add_argument("--cpu-steps", type=int, default=800000, help="physics substeps per CPU process for the CPU arm sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--learner", default="tcgen05", choices=["tcgen05", "tf32", "fp32"], help="--workload ppo: learner arithmetic")
    ap.add_argument("--workload", default="flat", choices=["flat", "scene", "ppo"],
                    help="flat = configs[3] headline + sub-records (default), scene = configs[2], ppo = configs[4] (training loop)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_cuda(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
