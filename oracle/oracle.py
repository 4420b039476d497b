"""ctypes face of oracle/ackb_oracle.c (CPU fp64 restatement of mj_step; TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Parity unpinned for the physics (see the header of ackb_oracle.c).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liback_oracle.so")
def _cpu_tag() -> str:
    """Identity of this host's CPU (model + ISA flags): a -march=native binary built elsewhere must never be loaded here."""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        key = "".join(sorted({ln.split(":", 1)[1].strip() for ln in txt.splitlines() if ln.startswith(("model name", "flags"))}))
    except OSError:
        key = "unknown"
    return hashlib.sha1(key.encode()).hexdigest()[:10]


# -O3 -march=native build for the TIMED CPU baseline (bench.py only), one file per host CPU type
_FAST_PATH = os.path.join(_HERE, f"liback_oracle_native_{_cpu_tag()}.so")
_lib = None
_libs = {}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ackb_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liback_oracle.so"])
    return _LIB_PATH


def build_fast(force: bool = False) -> str:
    """Timed-baseline build: same source, -O3 -march=native (compiled on the machine that runs it).  The checker build keeps
    -O2 -ffp-contract=off so that parity tests do not depend on the host's vector ISA."""
    src = os.path.join(_HERE, "ackb_oracle.c")
    if force or not os.path.exists(_FAST_PATH) or os.path.getmtime(_FAST_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-fno-fast-math", "-shared", "-o", _FAST_PATH, src, "-lm"])
    return _FAST_PATH


def lib(fast: bool = False):
    global _lib
    if fast not in _libs:
        path = build_fast() if fast else build()
        L = ctypes.CDLL(path)
        L.orc_model_new.restype = ctypes.c_void_p
        L.orc_data_new.restype = ctypes.c_void_p
        for fn in (L.orc_forward, L.orc_step, L.orc_reset):
            fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_step_n.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_free.argtypes = [ctypes.c_void_p]
        for fn in (L.orc_model_field, L.orc_data_field):
            fn.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int)]
            fn.restype = ctypes.c_int
        L.orc_contact.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.orc_rollout.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_void_p]
        L.orc_rollout.restype = ctypes.c_long
        _libs[fast] = L
        if not fast:
            _lib = L
    return _libs[fast]


def _view(fn, handle, name):
    ptr, is_int = ctypes.c_void_p(), ctypes.c_int()
    n = fn(handle, name.encode(), ctypes.byref(ptr), ctypes.byref(is_int))
    if n < 0:
        raise KeyError(name)
    ctype = ctypes.c_int if is_int.value else ctypes.c_double
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(n,))


_SCALARS = ["nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "neq", "nsensordata"]


class OracleSim:
    """One environment's MjModel/MjData analogue driven by the C oracle."""

    def __init__(self, model: dict, tolerance: float | None = None, fast: bool = False):
        L = lib(fast)
        self._L = L
        self.model = model
        self.m = ctypes.c_void_p(L.orc_model_new())
        self.d = ctypes.c_void_p(L.orc_data_new())
        for k in _SCALARS:
            _view(L.orc_model_field, self.m, k)[0] = int(model[k])
        _view(L.orc_model_field, self.m, "nsensor")[0] = len(model["sensor_adr"])
        _view(L.orc_model_field, self.m, "nhullvert")[0] = len(model["hull_vert"])
        for k, v in model.items():
            if k in _SCALARS or not isinstance(v, np.ndarray):
                continue
            try:
                dst = _view(L.orc_model_field, self.m, k)
            except KeyError:
                continue
            flat = np.asarray(v).reshape(-1)
            if flat.size > dst.size:
                raise ValueError(f"oracle capacity exceeded for {k}: {flat.size} > {dst.size}")
            dst[: flat.size] = flat
        if tolerance is not None:
            _view(L.orc_model_field, self.m, "opt_tolerance")[0] = tolerance
        self.nq, self.nv, self.nu = int(model["nq"]), int(model["nv"]), int(model["nu"])
        self.reset()

    def __del__(self):
        try:
            self._L.orc_free(self.m)
            self._L.orc_free(self.d)
        except Exception:
            pass

    def f(self, name):
        """NumPy view (no copy) of a data field at full capacity."""
        return _view(self._L.orc_data_field, self.d, name)

    def mf(self, name):
        return _view(self._L.orc_model_field, self.m, name)

    # --- convenience views ------------------------------------------------------------------
    @property
    def qpos(self):
        return self.f("qpos")[: self.nq]

    @property
    def qvel(self):
        return self.f("qvel")[: self.nv]

    @property
    def ctrl(self):
        return self.f("ctrl")[: self.nu]

    @property
    def qacc(self):
        return self.f("qacc")[: self.nv]

    @property
    def qacc_warmstart(self):
        return self.f("qacc_warmstart")[: self.nv]

    @property
    def sensordata(self):
        return self.f("sensordata")[: int(self.model["nsensordata"])]

    @property
    def xpos(self):
        return self.f("xpos")[: 3 * int(self.model["nbody"])].reshape(-1, 3)

    @property
    def xquat(self):
        return self.f("xquat")[: 4 * int(self.model["nbody"])].reshape(-1, 4)

    @property
    def ncon(self):
        return int(self.f("ncon")[0])

    @property
    def nefc(self):
        return int(self.f("nefc")[0])

    def qM(self):
        maxv = self._L.orc_maxv()
        return self.f("qM").reshape(maxv, maxv)[: self.nv, : self.nv].copy()

    def efc(self, name):
        n = self.nefc
        if name == "J":
            maxv = self._L.orc_maxv()
            return self.f("efc_J")[: n * maxv].reshape(n, maxv)[:, : self.nv].copy()
        return self.f("efc_" + name)[:n].copy()

    def contacts(self):
        out = []
        buf = (ctypes.c_double * 18)()
        for i in range(self.ncon):
            self._L.orc_contact(self.d, i, buf)
            a = np.array(buf)
            out.append(dict(geom1=int(a[0]), geom2=int(a[1]), dim=int(a[2]), exclude=int(a[3]), dist=a[4],
                            pos=a[5:8].copy(), frame=a[8:17].reshape(3, 3).copy(), mu=a[17]))
        return out

    # --- stepping ---------------------------------------------------------------------------
    def reset(self):
        self._L.orc_reset(self.m, self.d)

    def forward(self):
        self._L.orc_forward(self.m, self.d)

    def step(self, n: int = 1):
        self._L.orc_step_n(self.m, self.d, n)

    def rollout(self, n_steps: int, frame_skip: int = 1, max_steps: int = 1000, seed: int = 0, spawn_qpos=None) -> int:
        """C loop: random action -> BicycleController -> frame_skip x step, with episode resets (bench CPU baseline)."""
        sq = np.ascontiguousarray(self.model["qpos0"] if spawn_qpos is None else spawn_qpos, dtype=np.float64)
        return int(self._L.orc_rollout(self.m, self.d, n_steps, frame_skip, max_steps, seed, sq.ctypes.data_as(ctypes.c_void_p)))
