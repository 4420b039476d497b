#!/usr/bin/env python
"""Generate tests/golden/reference_glue.json by RUNNING THE REFERENCE'S OWN PYTHON (unmodified, imported from
/root/reference) for the non-physics half of the hot path, with `mujoco` and `gymnasium` replaced by stubs:

  * BicycleController / AckermannController.cmd_vel_to_controls + apply_cmd_vel   (src/core/controller.py)
  * Odometry.calculate_odom / _quat_to_yaw                                        (src/core/odometry.py)
  * AckermannRobotEnv._setup_lidar / reset / step glue / _get_observation / _calculate_reward
                                                                                  (src/rl/envs/ackermann_env.py)

mujoco.mj_step is a no-op in the stub: the test vectors set data.xpos / xquat / sensordata by hand, so what is pinned
is exactly the arithmetic the reference performs around the physics.  Run in the build container only
(/root/reference does not exist on the GPU box):   python tests/golden/gen_reference_glue.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from mujoco_playground_b200.models import load_model  # noqa: E402

M = load_model("v2")
S = load_model("scene")

# ---- stubs ----------------------------------------------------------------------------------------------------
mj = types.ModuleType("mujoco")


class _Obj:
    mjOBJ_SENSOR, mjOBJ_SITE, mjOBJ_BODY, mjOBJ_ACTUATOR = "sensor", "site", "body", "actuator"


mj.mjtObj = _Obj


def mj_name2id(model, kind, name):
    names = {"sensor": model._t["sensor_names"], "site": model._t["site_names"], "body": model._t["body_names"],
             "actuator": model._t["actuator_names"]}[kind]
    return names.index(name) if name in names else -1     # MuJoCo returns -1, it does not raise


mj.mj_name2id = mj_name2id
mj.mj_step = lambda model, data: None
mj.mj_forward = lambda model, data: None
mj.viewer = types.ModuleType("mujoco.viewer")
sys.modules["mujoco"] = mj
sys.modules["mujoco.viewer"] = mj.viewer
gym = types.ModuleType("gymnasium")


class _Env:
    def reset(self, seed=None, options=None):
        return None


gym.Env = _Env
spaces = types.ModuleType("gymnasium.spaces")


class Box:
    def __init__(self, low, high, shape, dtype):
        self.low = np.full(shape, low, dtype=dtype)
        self.high = np.full(shape, high, dtype=dtype)
        self.shape, self.dtype = shape, dtype


spaces.Box = Box
gym.spaces = spaces
sys.modules["gymnasium"] = gym
sys.modules["gymnasium.spaces"] = spaces


class FakeModel:
    def __init__(self, t):
        self._t = t
        self.sensor_adr = np.asarray(t["sensor_adr"])


class FakeData:
    def __init__(self, t):
        self.ctrl = np.zeros(t["nu"])
        self.qpos = np.zeros(t["nq"])
        self.xpos = np.zeros((t["nbody"], 3))
        self.xquat = np.tile([1.0, 0, 0, 0], (t["nbody"], 1))
        self.sensordata = np.zeros(t["nsensordata"])


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


sys.path.insert(0, os.path.join(REF, "src"))
controller = load(os.path.join(REF, "src", "core", "controller.py"), "core.controller")
envmod = load(os.path.join(REF, "src", "rl", "envs", "ackermann_env.py"), "ref_ackermann_env")

out = {"source": "reference python under mujoco/gymnasium stubs", "controller_bicycle": [], "controller_ackermann": [],
       "obs_reward": [], "step_glue": []}
rng = np.random.default_rng(2024)

# ---- controllers ------------------------------------------------------------------------------------------------------
bc = controller.BicycleController(FakeModel(M), FakeData(M))
ac = controller.AckermannController(FakeModel(S), FakeData(S))
cases = [(1, 0), (1, 1), (1, -1), (0.5, 1), (0, 1), (0, -1), (-1, 1), (-0.3, -0.7), (1e-6, 0.5), (0.2, 1e-7), (0, 0), (-1, 0),
         (1, 1e-6), (1, 9.9e-7), (0.3, 2e-5), (1.0, 4.0e-5), (0.5, -3.0e-5), (1e-5, 1), (1.0001e-5, -1), (-1e-5, 0.3)]
cases += [tuple(np.float32(rng.uniform(-1, 1, 2)).tolist()) for _ in range(60)]
for v, w in cases:
    v32, w32 = np.float32(v), np.float32(w)
    with np.errstate(all="ignore"):
        bc.apply_cmd_vel(v32, w32)
    out["controller_bicycle"].append({"v": float(v32), "omega": float(w32), "ctrl": [None if not np.isfinite(x) else float(x) for x in bc.data.ctrl]})
    try:
        with np.errstate(all="ignore"):
            ac.apply_cmd_vel(v32, w32)
        out["controller_ackermann"].append({"v": float(v32), "omega": float(w32), "ctrl": [float(x) for x in ac.data.ctrl]})
    except ZeroDivisionError:
        out["controller_ackermann"].append({"v": float(v32), "omega": float(w32), "ctrl": None})


# ---- env glue ---------------------------------------------------------------------------------------------------------------
class Spawner:
    def load_random_environment(self, robot_pos=None, robot_quat=None):
        m, d = FakeModel(M), FakeData(M)
        d.qpos[0:3] = robot_pos
        d.xpos[M["body_names"].index("chassis")] = robot_pos      # what mj_forward would produce
        d.sensordata[:] = -1.0
        return m, d, "simple_floor"


import builtins  # noqa: E402
_print = builtins.print
builtins.print = lambda *a, **k: None       # Odometry prints on initialisation
env = envmod.AckermannRobotEnv(map_spawner=Spawner())
obs, info = env.reset(seed=0)
out["lidar_addrs"] = [int(a) for a in env.lidar_addrs]
out["reset"] = {"obs": [float(x) for x in obs], "goal": [float(x) for x in env.goal_position], "info_keys": sorted(info.keys()),
                "obs_dtype": str(obs.dtype)}
ch = M["body_names"].index("chassis")
for i in range(40):
    env.goal_position = rng.uniform(-8, 8, 2)
    pos = np.array([rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(0.05, 0.12)])
    q = rng.normal(size=4) * np.array([1, 0.05, 0.05, 1]) if i % 2 else rng.normal(size=4)
    q /= np.linalg.norm(q)
    sd = rng.uniform(-1, 12, M["nsensordata"])
    if i % 3 == 0:
        sd[5:] = -1.0
    if i % 5 == 0:
        env.goal_position = pos[:2] - env.odometry.reference_position[:2] + rng.uniform(-0.3, 0.3, 2)   # near the goal
    env.data.xpos[ch] = pos
    env.data.xquat[ch] = q
    env.data.sensordata[:] = sd
    o = env._get_observation()
    r, te, tr, inf = env._calculate_reward()
    out["obs_reward"].append({"xpos": pos.tolist(), "xquat": q.tolist(), "sensordata": sd.tolist(), "goal": env.goal_position.tolist(),
                              "reference_position": env.odometry.reference_position.tolist(), "obs": [float(x) for x in o],
                              "reward": float(r), "terminated": bool(te), "truncated": bool(tr), "collision": bool(inf["collision"]),
                              "min_lidar": float(inf["min_lidar"]), "goal_distance": float(inf["goal_distance"])})
# step glue: counter / truncation / ctrl written, with mj_step a no-op
env2 = envmod.AckermannRobotEnv(map_spawner=Spawner(), max_episode_steps=3, max_linear_velocity=0.7, max_angular_velocity=0.9)
env2.reset()
for a in ([0.5, 0.25], [2.0, -3.0], [-0.1, 0.9], [0.0, 0.0]):
    o, r, te, tr, inf = env2.step(np.array(a, dtype=np.float32))
    out["step_glue"].append({"action": a, "ctrl": [float(x) for x in env2.data.ctrl], "step": int(inf["step"]), "truncated": bool(tr),
                             "terminated": bool(te), "linear_velocity": float(inf["linear_velocity"]),
                             "angular_velocity": float(inf["angular_velocity"]), "info_keys": sorted(inf.keys())})
builtins.print = _print
json.dump(out, open(os.path.join(HERE, "reference_glue.json"), "w"), indent=0)
print("wrote reference_glue.json:", {k: (len(v) if isinstance(v, list) else "…") for k, v in out.items()})
