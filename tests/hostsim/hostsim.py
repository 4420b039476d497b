"""ctypes face of the TEST-ONLY host build of the per-environment device code (tests/hostsim/hostsim.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.join(_HERE, "..", "..")
_lib = None


def lib():
    global _lib
    if _lib is None:
        tgt = os.path.join(_HERE, "libhostsim.so")
        csrc = os.path.join(_ROOT, "mujoco_playground_b200", "csrc")
        srcs = [os.path.join(_HERE, "hostsim.cpp"), os.path.join(csrc, "ackb_core.cuh"), os.path.join(csrc, "ackb_env.cuh"),
                os.path.join(_ROOT, "include", "ackb_consts.def")]
        if not os.path.exists(tgt) or any(os.path.getmtime(s) > os.path.getmtime(tgt) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-I" + os.path.join(_ROOT, "include"),
                                   "-o", tgt, srcs[0]])
        _lib = ctypes.CDLL(tgt)
        _lib.hs_hs = None
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class HostSim:
    """One environment stepped by the host build of the CUDA per-environment code (LANES = 1)."""

    def __init__(self, consts: np.ndarray, f32: bool = False):
        self.L = lib()
        assert self.L.hs_nconsts() == len(consts)
        self.blob = np.ascontiguousarray(consts, dtype=np.float64)
        self.f32 = int(f32)
        self.qpos, self.qvel, self.warm = np.zeros(13), np.zeros(12), np.zeros(12)
        self.epd = np.zeros(4)                 # goal xy, odometry reference xy
        self.epi = np.zeros(2, np.int32)       # step counter, episode counter
        self.obs = np.zeros(int(round(consts_field(consts, "nbeam"))) + 7, np.float32)
        self.diag = np.zeros(4, np.int32)
        self.tap = np.zeros(50)

    def reset(self, seed=0, env_id=0):
        self.L.hs_env_reset(self.f32, _p(self.blob), _p(self.qpos), _p(self.qvel), _p(self.warm), _p(self.epd), _p(self.epi),
                            ctypes.c_ulonglong(seed), ctypes.c_uint(env_id), _p(self.obs))
        return self.obs.copy()

    def substep(self, ctrl, n=1):
        c4 = np.zeros(4)
        c4[: len(ctrl)] = ctrl
        self.L.hs_substep(self.f32, _p(self.blob), _p(self.qpos), _p(self.qvel), _p(self.warm), _p(c4), n, _p(self.tap), _p(self.diag))

    def step(self, action, frame_skip=1, lanes=1):
        a = np.ascontiguousarray(action, dtype=np.float32)
        out = np.zeros(6, np.float32)
        if lanes > 1:   # LANES host threads emulate the lanes of the environment
            self.L.hs_env_step_lanes(lanes, self.f32, _p(self.blob), _p(self.qpos), _p(self.qvel), _p(self.warm), _p(self.epd), _p(self.epi), _p(a),
                                     frame_skip, _p(self.obs), _p(out), _p(self.diag))
            return self.obs.copy(), float(out[0]), bool(out[1]), bool(out[2]), dict(collision=bool(out[3]), goal_distance=float(out[4]),
                                                                                     min_lidar=float(out[5]), ncon=int(self.diag[0]),
                                                                                     unsupported=int(self.diag[1]), niter=int(self.diag[2]), bad=int(self.diag[3]))
        self.L.hs_env_step(self.f32, _p(self.blob), _p(self.qpos), _p(self.qvel), _p(self.warm), _p(self.epd), _p(self.epi), _p(a),
                           frame_skip, _p(self.obs), _p(out), _p(self.diag))
        return self.obs.copy(), float(out[0]), bool(out[1]), bool(out[2]), dict(collision=bool(out[3]), goal_distance=float(out[4]),
                                                                                 min_lidar=float(out[5]), ncon=int(self.diag[0]),
                                                                                 unsupported=int(self.diag[1]), niter=int(self.diag[2]), bad=int(self.diag[3]))

    def observe(self):
        out = np.zeros(2)
        self.L.hs_observe(self.f32, _p(self.blob), _p(self.qpos), _p(self.qvel), _p(self.warm), _p(self.epd), _p(self.obs), _p(out))
        return self.obs.copy(), out[0], out[1]

    def reward(self, dist, min_lidar, step_count):
        out = np.zeros(5)
        self.L.hs_reward(self.f32, _p(self.blob), ctypes.c_double(dist), ctypes.c_double(min_lidar), int(step_count), _p(out))
        return out

    def action_to_ctrl(self, a0, a1):
        c = np.zeros(4)
        self.L.hs_action_to_ctrl(self.f32, _p(self.blob), ctypes.c_float(a0), ctypes.c_float(a1), _p(c))
        return c


def consts_field(consts, name):
    from mujoco_playground_b200.compiler.constants import consts_layout
    off, cnt = consts_layout()[name]
    return consts[off] if cnt == 1 else consts[off:off + cnt]
