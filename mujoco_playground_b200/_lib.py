"""ctypes binding of libackb.so (include/ackb.h).  There is no CPU fallback: a missing or unloadable
CUDA library is an error."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACKB_LIB", os.path.join(_HERE, "libackb.so"))   # ACKB_LIB: tuning variants only
_lib = None


class AckbStats(ctypes.Structure):
    _fields_ = [("episodes", ctypes.c_ulonglong), ("successes", ctypes.c_ulonglong), ("env_steps", ctypes.c_ulonglong),
                ("collisions", ctypes.c_ulonglong), ("unsupported", ctypes.c_ulonglong), ("solver_iters", ctypes.c_ulonglong),
                ("return_sum", ctypes.c_double), ("length_sum", ctypes.c_double),
                ("obstacle_steps", ctypes.c_ulonglong), ("contacts_sum", ctypes.c_ulonglong), ("bad_state", ctypes.c_ulonglong)]


# every symbol include/ackb.h declares: (restype, argtypes)
_vp, _i, _u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
SYMBOLS = {
    "ackb_consts_len": (_i, []),
    "ackb_create": (_i, [_vp, ctypes.c_size_t, _i, _i, _i, _u64, _i, ctypes.POINTER(_vp)]),
    "ackb_destroy": (_i, [_vp]),
    "ackb_set_env_id_base": (_i, [_vp, _u64]),
    "ackb_set_obs_pitch": (_i, [_vp, _i]),
    "ackb_num_envs": (_i, [_vp]),
    "ackb_obs_dim": (_i, [_vp]),
    "ackb_dtype": (_i, [_vp]),
    "ackb_reset": (_i, [_vp, _vp, _vp, _vp]),
    "ackb_step": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ackb_step_host": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "ackb_get_state": (_i, [_vp, _vp, _vp, _vp]),
    "ackb_set_state": (_i, [_vp, _vp, _vp, _vp]),
    "ackb_get_episode": (_i, [_vp, _vp, _vp, _vp]),
    "ackb_set_episode": (_i, [_vp, _vp, _vp, _vp]),
    "ackb_stats": (_i, [_vp, ctypes.POINTER(AckbStats)]),
    "ackb_stats_reset": (_i, [_vp]),
    "ackb_launch_count": (ctypes.c_ulonglong, [_vp]),
    "ackb_random_actions": (_i, [_vp, _vp, _vp]),
    "ackb_last_error": (ctypes.c_char_p, [_vp]),
    # include/ackb_ppo.h
    "ackb_ppo_num_params": (_i, [_i]),
    "ackb_ppo_set_mode": (_i, [_i]),
    "ackb_ppo_act": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _u64, ctypes.c_uint32, _i, _vp]),
    "ackb_ppo_bootstrap": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, ctypes.c_float, _vp, _vp, _vp]),
    "ackb_ppo_permutation": (_i, [_vp, ctypes.c_longlong, _u64, ctypes.c_uint32, _vp]),
    "ackb_ppo_adv_stats": (_i, [_vp, _vp, _i, _vp, _vp]),
    "ackb_ppo_clip_adam": (_i, [_vp, _vp, _vp, _vp, _vp, _i] + [ctypes.c_float] * 5 + [_vp]),
    "ackb_ppo_clip_adam_allreduce": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i] + [ctypes.c_float] * 5 + [_vp]),
    "ackb_ppo_gae": (_i, [_vp, _vp, _vp, _vp, _i, _i, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp]),
    "ackb_ppo_adv_stats_ws": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "ackb_ppo_minibatch_grad_mode": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, _i, _vp]),
    "ackb_ppo_minibatch_grad_pitched": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, _i, _vp]),
    "ackb_ppo_act_pitched": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _u64, ctypes.c_uint32, _i, _vp]),
    "ackb_ppo_minibatch_grad_stats": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, _i, _vp]),
    "ackb_ppo_bootstrap_pitched": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _vp, ctypes.c_float, _vp, _vp, _vp]),
    "ackb_ppo_minibatch_grad": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, _vp]),
}


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the simulator has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, handle=None):
    if rc != 0:
        msg = load().ackb_last_error(handle)
        raise RuntimeError(f"ackb error {rc}: {msg.decode() if msg else '?'}")
