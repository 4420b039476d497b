#!/usr/bin/env python
"""Golden trajectories for the physics parity tests ("Plan when MuJoCo becomes available").

    python tools/dump_golden.py --reference /path/to/ulusoyn/mujoco_playground --out tests/golden/mujoco_golden.npz

needs the third-party `mujoco` package (absent from the build image and from the GPU box) and writes, for every scenario
below, the compiled model (tools/dump_mjmodel.py table, prefix "<scenario>/model/") and per step
    act[T,2] f32, ctrl[T,nu], qpos[T+1,nq], qvel[T+1,nv], qacc[T,nv], qacc_warmstart[T,nv], ncon[T], nefc[T], sensordata[T,ns],
    xpos_chassis[T,3], xquat_chassis[T,4]           (what the env reads after mj_step: pre-integration kinematics, quirk Q3)
and, for the first `--efc-steps` steps, the contact list (dist, pos, frame, geoms) and efc_type / efc_pos / efc_D / efc_R /
efc_aref / efc_force / efc_J (dense) -- the intermediates that settle the VERIFY items of SURVEY Appendix B (DESIGN.md section 2
lists which field settles which item).  `tests/test_mujoco_golden.py` consumes the file; without it those tests report
SKIPPED (no oracle).

    python tools/dump_golden.py --backend oracle --out /tmp/self.npz

writes the SAME FORMAT from this repo's own CPU oracle (marked backend = "oracle").  That file proves nothing about MuJoCo; the
test-suite uses it to exercise the consumer code path so that a real golden file drops in without surprises.

Scenarios
  cfg1      BASELINE configs[0]: ackermann_robot_v2.xml, reset state of simple_map_spawner.py:43-50, 1000 steps, actions
            U(-1,1)^2 float32 from numpy default_rng(0) through BicycleController (max velocities 1.0 / 1.0)
  scene     models/environments/ackermann_in_mushr_maze.xml, XML spawn pose, 400 steps, AckermannController, default_rng(1)
  rollover  v2 upside down (180 deg about x) dropped from z = 0.12: wheel tops / plate hulls vs floor
  nosedown  v2 pitched 80 deg nose down from z = 0.20: plate hull (mjc_PlaneConvex) contacts
  onside    v2 rolled 90 deg from z = 0.13: wheel caps on the floor (the two extra mjc_PlaneCylinder "triangle" points)
  wallhit   v2 + one box wall 0.45 m ahead (worldbody patched), full throttle for 500 steps: plate hull vs box, wheel vs box
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from dump_mjmodel import model_table, save_table  # noqa: E402


def _quat(axis, deg):
    a = np.deg2rad(deg) / 2
    return np.array([np.cos(a), *(np.sin(a) * np.asarray(axis, float))])


def scenarios(ref):
    v2 = os.path.join(ref, "models", "ackermann_robot_v2.xml")
    sc = os.path.join(ref, "models", "environments", "ackermann_in_mushr_maze.xml")
    return {
        "cfg1": dict(xml=v2, kind="v2", steps=1000, rng=0, pos=[0, 0, 0.1], quat=[1, 0, 0, 0]),
        "scene": dict(xml=sc, kind="scene", steps=400, rng=1),
        "rollover": dict(xml=v2, kind="v2", steps=250, rng=None, pos=[0, 0, 0.12], quat=_quat([1, 0, 0], 180)),
        "nosedown": dict(xml=v2, kind="v2", steps=400, rng=None, pos=[0, 0, 0.20], quat=_quat([0, 1, 0], 80)),
        "onside": dict(xml=v2, kind="v2", steps=250, rng=None, pos=[0, 0, 0.13], quat=_quat([1, 0, 0], 90)),
        "wallhit": dict(xml=v2, kind="v2", steps=500, rng="throttle", pos=[0, 0, 0.1], quat=[1, 0, 0, 0],
                        patch=dict(name="wall", type="box", size="0.5 0.5 0.5", pos="1.1 0 0.5")),
    }


def patched_xml(sc) -> str:
    """Path of the scenario's XML; with a `patch` a copy with absolute mesh paths and the extra world geom (temp dir)."""
    if "patch" not in sc:
        return sc["xml"]
    tree = ET.parse(sc["xml"])
    root = tree.getroot()
    base = os.path.dirname(os.path.abspath(sc["xml"]))
    for mesh in root.iter("mesh"):
        if "file" in mesh.attrib:
            mesh.attrib["file"] = os.path.normpath(os.path.join(base, mesh.attrib["file"]))
    ET.SubElement(root.find("worldbody"), "geom", sc["patch"])
    d = tempfile.mkdtemp(prefix="ackb_golden_")
    out = os.path.join(d, "patched.xml")
    tree.write(out)
    return out


def actions(sc):
    T = sc["steps"]
    if sc["rng"] is None:
        return np.zeros((T, 2), np.float32)
    if sc["rng"] == "throttle":
        return np.tile(np.array([[1.0, 0.0]], np.float32), (T, 1))
    return np.random.default_rng(sc["rng"]).uniform(-1, 1, size=(T, 2)).astype(np.float32)


def ctrl_of(kind, a):
    from oracle.env_oracle import ackermann_ctrl, bicycle_ctrl      # pinned against the reference's controller.py (tests/golden)
    a = np.clip(np.asarray(a, np.float32), np.float32(-1), np.float32(1))
    c = bicycle_ctrl(a[0] * np.float32(1.0), a[1] * np.float32(1.0)) if kind == "v2" else ackermann_ctrl(a[0] * np.float32(1.0), a[1] * np.float32(1.0))
    return c if np.all(np.abs(c) <= 1e10) else np.zeros_like(c)


# ----------------------------------------------------------------------------------------------------------------------
class MujocoBackend:
    name = "mujoco"

    def __init__(self, xml):
        import mujoco
        self.mj = mujoco
        self.m = mujoco.MjModel.from_xml_path(xml)
        self.d = mujoco.MjData(self.m)
        self.chassis = mujoco.mj_name2id(self.m, mujoco.mjtObj.mjOBJ_BODY, "chassis")

    def table(self):
        return model_table(self.m)

    def set_pose(self, pos, quat):
        self.d.qpos[0:3] = pos; self.d.qpos[3:7] = quat
        self.mj.mj_forward(self.m, self.d)

    def step(self, ctrl):
        self.d.ctrl[:] = ctrl
        self.mj.mj_step(self.m, self.d)

    def state(self):
        d = self.d
        return d.qpos.copy(), d.qvel.copy()

    def after_step(self):
        d, m = self.d, self.m
        return dict(qacc=d.qacc.copy(), qacc_warmstart=d.qacc_warmstart.copy(), ncon=int(d.ncon), nefc=int(d.nefc), sensordata=d.sensordata.copy(),
                    xpos_chassis=d.xpos[self.chassis].copy(), xquat_chassis=d.xquat[self.chassis].copy())

    def efc(self):
        d, m = self.d, self.m
        n, nv = int(d.nefc), m.nv
        J = np.array(d.efc_J).reshape(-1)
        if J.size == n * nv:
            J = J.reshape(n, nv)
        else:       # sparse Jacobian storage
            J = np.zeros((n, nv))
            self.mj.mju_sparse2dense(J, d.efc_J, d.efc_J_rownnz, d.efc_J_rowadr, d.efc_J_colind)
        con = np.zeros((int(d.ncon), 16))
        for i in range(int(d.ncon)):
            c = d.contact[i]
            g = getattr(c, "geom", None)
            g1, g2 = (int(g[0]), int(g[1])) if g is not None else (int(c.geom1), int(c.geom2))
            con[i] = [g1, g2, c.dim, c.exclude, c.dist, *c.pos, *c.frame[:6], c.mu if hasattr(c, "mu") else 0.0, c.includemargin]
        return dict(efc_type=np.array(d.efc_type[:n]), efc_pos=np.array(d.efc_pos[:n]), efc_D=np.array(d.efc_D[:n]), efc_R=np.array(d.efc_R[:n]),
                    efc_aref=np.array(d.efc_aref[:n]), efc_force=np.array(d.efc_force[:n]), efc_J=J, contact=con)


class OracleBackend:
    name = "oracle"

    def __init__(self, xml):
        from mujoco_playground_b200.compiler.mjcf import compile_mjcf
        from oracle.oracle import OracleSim
        self.T = compile_mjcf(xml)
        self.s = OracleSim(self.T, tolerance=1e-12)
        self.chassis = self.T["body_names"].index("chassis")

    def table(self):
        return self.T

    def set_pose(self, pos, quat):
        self.s.qpos[0:3] = pos; self.s.qpos[3:7] = quat
        self.s.forward()

    def step(self, ctrl):
        self.s.ctrl[:] = ctrl
        self.s.step()

    def state(self):
        return self.s.qpos.copy(), self.s.qvel.copy()

    def after_step(self):
        s = self.s
        return dict(qacc=s.qacc.copy(), qacc_warmstart=s.qacc_warmstart.copy(), ncon=s.ncon, nefc=s.nefc, sensordata=s.sensordata.copy(),
                    xpos_chassis=s.xpos[self.chassis].copy(), xquat_chassis=s.xquat[self.chassis].copy())

    def efc(self):
        s = self.s
        con = np.array([[c["geom1"], c["geom2"], c["dim"], c["exclude"], c["dist"], *c["pos"], *c["frame"].reshape(-1)[:6], c["mu"], 0.0]
                        for c in s.contacts()]).reshape(-1, 16)
        return dict(efc_type=s.efc("type"), efc_pos=s.efc("pos"), efc_D=s.efc("D"), efc_R=s.efc("R"), efc_aref=s.efc("aref"),
                    efc_force=s.efc("force"), efc_J=s.efc("J"), contact=con)


def run_scenario(name, sc, backend_cls, efc_steps):
    be = backend_cls(patched_xml(sc))
    out = {}
    if "pos" in sc:
        be.set_pose(sc["pos"], sc["quat"])
    acts = actions(sc)
    T = len(acts)
    q0, v0 = be.state()
    rec = dict(qpos=[q0], qvel=[v0], ctrl=[], qacc=[], qacc_warmstart=[], ncon=[], nefc=[], sensordata=[], xpos_chassis=[], xquat_chassis=[])
    for t in range(T):
        c = ctrl_of(sc["kind"], acts[t])
        be.step(c)
        rec["ctrl"].append(c)
        a = be.after_step()
        for k, v in a.items():
            rec[k].append(v)
        q, v = be.state()
        rec["qpos"].append(q); rec["qvel"].append(v)
        if t < efc_steps:
            for k, v in be.efc().items():
                out[f"{name}/efc/{t}/{k}"] = np.asarray(v)
    out[f"{name}/act"] = acts
    for k, v in rec.items():
        out[f"{name}/{k}"] = np.asarray(v)
    out[f"{name}/kind"] = np.array([sc["kind"]], dtype="U16")
    return out, be.table()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "mujoco_golden.npz"))
    ap.add_argument("--backend", default="mujoco", choices=["mujoco", "oracle"])
    ap.add_argument("--efc-steps", type=int, default=50)
    ap.add_argument("--only", default="", help="comma-separated scenario names (default: all)")
    ap.add_argument("--max-steps", type=int, default=0, help="truncate every scenario to this many steps (small fixtures)")
    a = ap.parse_args(argv)
    be = MujocoBackend if a.backend == "mujoco" else OracleBackend
    out = {"backend": np.array([a.backend], dtype="U16")}
    if a.backend == "mujoco":
        import mujoco
        out["mujoco_version"] = np.array([mujoco.__version__], dtype="U16")
    tmp = tempfile.mkdtemp(prefix="ackb_tbl_")
    names = []
    for name, sc in scenarios(a.reference).items():
        if a.only and name not in a.only.split(","):
            continue
        if a.max_steps:
            sc = dict(sc, steps=min(sc["steps"], a.max_steps))
        rec, table = run_scenario(name, sc, be, a.efc_steps)
        out.update(rec)
        p = os.path.join(tmp, name + ".npz")
        save_table(table, p, prefix=f"{name}/model/")
        with np.load(p) as z:
            out.update({k: z[k] for k in z.files})
        names.append(name)
        print(f"{name}: {len(rec[name + '/act'])} steps, final ncon {int(rec[name + '/ncon'][-1])}")
    out["scenarios"] = np.array(names, dtype="U16")
    np.savez_compressed(a.out, **out)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
