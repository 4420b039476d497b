#!/usr/bin/env python
"""Dump a compiled MuJoCo model into this repo's table format (the "override blob" of SURVEY.md 7.1 / 8c).

Runs only where the third-party `mujoco` package is installed (it is NOT in the build image nor on the GPU box):

    python tools/dump_mjmodel.py /path/to/reference/models/ackermann_robot_v2.xml  v2_mujoco.npz

The output has exactly the keys `mujoco_playground_b200.models.load_model` returns (MuJoCo's own field names), so it can
replace a table compiled by `mujoco_playground_b200/compiler/mjcf.py`:

    env = BatchedAckermannEnv(n, model="v2", model_table=load_table("v2_mujoco.npz"))     # kernels
    OracleSim(load_table("v2_mujoco.npz"))                                                  # CPU oracle

Every version-dependent constant (mesh inertia mode, hull vertices, `<replicate>` names, invweight0, meaninertia) then comes
from MuJoCo itself.  `tests/test_mujoco_golden.py` compares such a table with the compiled one field by field.
"""
from __future__ import annotations

import sys

import numpy as np

# our sensor codes (compiler/mjcf.py): 0 jointpos, 1 jointvel, 2 rangefinder
_SENSOR_CODES = {"mjSENS_JOINTPOS": 0, "mjSENS_JOINTVEL": 1, "mjSENS_RANGEFINDER": 2}


def model_table(m) -> dict:
    """`m`: mujoco.MjModel.  Returns the table dict (numpy arrays, lists of names, int scalars)."""
    import mujoco
    T = {}
    o = m.opt
    T["opt_timestep"] = np.array([o.timestep]); T["opt_gravity"] = np.array(o.gravity, float)
    T["opt_impratio"] = np.array([o.impratio]); T["opt_tolerance"] = np.array([o.tolerance])
    T["opt_ls_tolerance"] = np.array([o.ls_tolerance])
    T["opt_iterations"] = np.array([o.iterations], np.int32); T["opt_ls_iterations"] = np.array([o.ls_iterations], np.int32)
    for k in ("nq", "nv", "nbody", "njnt", "ngeom", "nsite", "neq", "nu", "nsensordata"):
        T[k] = int(getattr(m, k))

    def names(obj, n):
        return [mujoco.mj_id2name(m, obj, i) or "" for i in range(n)]
    O = mujoco.mjtObj
    for k in ("body_parentid", "body_jntnum", "body_jntadr", "body_dofnum", "body_dofadr", "body_rootid", "body_weldid"):
        T[k] = np.array(getattr(m, k), np.int32)
    for k in ("body_pos", "body_quat", "body_mass", "body_ipos", "body_iquat", "body_inertia", "body_invweight0"):
        T[k] = np.array(getattr(m, k), float)
    T["body_names"] = names(O.mjOBJ_BODY, m.nbody)
    for k in ("jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_limited"):
        T[k] = np.array(getattr(m, k), np.int32)
    for k in ("jnt_pos", "jnt_axis", "jnt_range", "jnt_margin", "jnt_solref", "jnt_solimp", "qpos0"):
        T[k] = np.array(getattr(m, k), float)
    T["jnt_names"] = names(O.mjOBJ_JOINT, m.njnt)
    for k in ("dof_bodyid", "dof_jntid", "dof_parentid"):
        T[k] = np.array(getattr(m, k), np.int32)
    for k in ("dof_armature", "dof_damping", "dof_frictionloss", "dof_solref", "dof_solimp", "dof_invweight0"):
        T[k] = np.array(getattr(m, k), float)
    for k in ("geom_type", "geom_bodyid", "geom_contype", "geom_conaffinity", "geom_condim", "geom_priority"):
        T[k] = np.array(getattr(m, k), np.int32)
    for k in ("geom_size", "geom_pos", "geom_quat", "geom_friction", "geom_solref", "geom_solimp", "geom_solmix", "geom_margin", "geom_gap",
              "geom_rbound"):
        T[k] = np.array(getattr(m, k), float)
    T["geom_alpha"] = np.array(m.geom_rgba[:, 3], float)
    T["geom_names"] = names(O.mjOBJ_GEOM, m.ngeom)
    # convex hulls of the mesh geoms: vertices (mesh frame, as MuJoCo re-centred it) + the hull graph MuJoCo walks in mjc_PlaneConvex
    hadr, hnum, hv, graphs = [], [], [], []
    for g in range(m.ngeom):
        did = int(m.geom_dataid[g])
        if int(m.geom_type[g]) != int(mujoco.mjtGeom.mjGEOM_MESH) or did < 0 or int(m.mesh_graphadr[did]) < 0:
            hadr.append(0); hnum.append(0)
            continue
        ga = int(m.mesh_graphadr[did])
        nvert, nface = int(m.mesh_graph[ga]), int(m.mesh_graph[ga + 1])
        gid = np.array(m.mesh_graph[ga + 2 + nvert: ga + 2 + 2 * nvert])
        verts = np.array(m.mesh_vert[int(m.mesh_vertadr[did]): int(m.mesh_vertadr[did]) + int(m.mesh_vertnum[did])], float)
        hadr.append(sum(len(v) for v in hv)); hnum.append(nvert)
        hv.append(verts[gid])
        graphs.append(np.array(m.mesh_graph[ga: ga + 2 + 3 * nvert + 6 * nface], np.int32))
    T["geom_hulladr"] = np.array(hadr, np.int32); T["geom_hullnum"] = np.array(hnum, np.int32)
    T["hull_vert"] = np.concatenate(hv) if hv else np.zeros((0, 3))
    T["hull_graph"] = np.concatenate(graphs) if graphs else np.zeros(0, np.int32)     # concatenated mesh_graph blocks, hull order
    T["site_bodyid"] = np.array(m.site_bodyid, np.int32); T["site_pos"] = np.array(m.site_pos, float); T["site_quat"] = np.array(m.site_quat, float)
    T["site_names"] = names(O.mjOBJ_SITE, m.nsite)
    T["sensor_names"] = names(O.mjOBJ_SENSOR, m.nsensor)
    codes = {int(getattr(mujoco.mjtSensor, k)): v for k, v in _SENSOR_CODES.items()}
    T["sensor_type"] = np.array([codes.get(int(t), -1) for t in m.sensor_type], np.int32)
    T["sensor_objid"] = np.array(m.sensor_objid, np.int32); T["sensor_cutoff"] = np.array(m.sensor_cutoff, float)
    T["sensor_adr"] = np.array(m.sensor_adr, np.int32)
    T["eq_obj1id"] = np.array(m.eq_obj1id, np.int32); T["eq_obj2id"] = np.array(m.eq_obj2id, np.int32)
    T["eq_data"] = np.array(m.eq_data, float)[:, :5]; T["eq_solref"] = np.array(m.eq_solref, float); T["eq_solimp"] = np.array(m.eq_solimp, float)
    T["actuator_names"] = names(O.mjOBJ_ACTUATOR, m.nu)
    T["actuator_trnid"] = np.array(m.actuator_trnid[:, 0], np.int32); T["actuator_gear"] = np.array(m.actuator_gear[:, 0], float)
    T["actuator_gainprm"] = np.array(m.actuator_gainprm[:, 0], float); T["actuator_biasprm"] = np.array(m.actuator_biasprm[:, :3], float)
    T["actuator_ctrllimited"] = np.array(m.actuator_ctrllimited, np.int32); T["actuator_ctrlrange"] = np.array(m.actuator_ctrlrange, float)
    T["actuator_forcelimited"] = np.array(m.actuator_forcelimited, np.int32); T["actuator_forcerange"] = np.array(m.actuator_forcerange, float)
    T["stat_meaninertia"] = np.array([m.stat.meaninertia])
    d = mujoco.MjData(m)
    mujoco.mj_forward(m, d)
    qM = np.zeros((m.nv, m.nv))
    mujoco.mj_fullM(m, qM, d.qM)
    T["qM0"] = qM
    T["mujoco_version"] = [mujoco.__version__]
    return T


def save_table(T: dict, path: str, prefix: str = "") -> None:
    out = {}
    for k, v in T.items():
        out[prefix + k] = np.array(v, dtype="U64") if isinstance(v, list) else np.asarray(v)
    np.savez_compressed(path, **out)


def load_table(path_or_npz, prefix: str = "") -> dict:
    """Inverse of save_table (also accepts an opened npz): the dict `load_model` would return."""
    z = np.load(path_or_npz, allow_pickle=False) if isinstance(path_or_npz, str) else path_or_npz
    scalars = ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "neq", "nsensordata")
    T = {}
    for k in z.files:
        if not k.startswith(prefix):
            continue
        name, v = k[len(prefix):], z[k]
        T[name] = int(v) if name in scalars else ([str(s) for s in v] if v.dtype.kind == "U" else v)
    return T


def main(argv):
    if len(argv) != 3:
        raise SystemExit(__doc__)
    import mujoco
    m = mujoco.MjModel.from_xml_path(argv[1])
    save_table(model_table(m), argv[2])
    print(f"wrote {argv[2]} (mujoco {mujoco.__version__})")


if __name__ == "__main__":
    main(sys.argv)
