"""Table model (mjcf.compile_mjcf) -> flat constants block for the CUDA kernels.

The kernels are specialised to the Ackermann topology; this module checks every structural
assumption they make and raises if a model violates one, rather than letting the kernels be
silently wrong.  Layout of the block: include/ackb_consts.def.
"""
from __future__ import annotations

import os
import re
from typing import Dict, Optional, Tuple

import numpy as np

from .setconst import forward_kinematics, _q2m

_DEF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "include", "ackb_consts.def")
MINVAL = 1e-15


def consts_layout(path: str = _DEF) -> Dict[str, Tuple[int, int]]:
    """name -> (offset, count) parsed from the X-macro file."""
    out, off = {}, 0
    for name, cnt in re.findall(r"^ACKB_FIELD\((\w+),\s*(\d+)\)", open(path).read(), flags=re.M):
        out[name] = (off, int(cnt))
        off += int(cnt)
    return out


def _check_power(solimp):
    assert solimp[4] in (1, 2), "solimp power must be 1 or 2 (other powers are not compiled into the kernels)"


def _impedance0(solimp):
    _check_power(solimp)
    return float(np.clip(solimp[0], 0.0001, 0.9999))


def _KB(solref, solimp, timestep):
    _check_power(solimp)
    tc = max(solref[0], 2 * timestep)
    dmax = float(np.clip(solimp[1], 0.0001, 0.9999))
    if solref[0] <= 0:
        raise NotImplementedError("direct (negative) solref")
    return 1.0 / max(MINVAL, dmax * dmax * tc * tc * solref[1] * solref[1]), 2.0 / max(MINVAL, dmax * tc)


def _mix(M, g1, g2):
    s1, s2 = M["geom_solmix"][g1], M["geom_solmix"][g2]
    mix = s1 / (s1 + s2)
    fr = np.maximum(M["geom_friction"][g1], M["geom_friction"][g2])
    solref = mix * M["geom_solref"][g1] + (1 - mix) * M["geom_solref"][g2]
    solimp = mix * M["geom_solimp"][g1] + (1 - mix) * M["geom_solimp"][g2]
    return fr, solref, solimp


def consts_get(blob, lay, name):
    off, cnt = lay[name]
    return blob[off] if cnt == 1 else blob[off:off + cnt]


def _box_grid(centres: np.ndarray, half: np.ndarray):
    """Occupancy grid of identical axis-aligned boxes, or None if they do not sit on a lattice of box-sized cells
    listed in (x, y) lexicographic order (the order fixes the contact slot order of wheel-box contacts)."""
    pitch = 2.0 * float(half[0])
    if not np.isclose(half[0], half[1]) or pitch <= 0:
        return None
    x0, y0 = centres[:, 0].min() - half[0], centres[:, 1].min() - half[1]
    fx, fy = (centres[:, 0] - x0) / pitch - 0.5, (centres[:, 1] - y0) / pitch - 0.5
    ix, iy = np.rint(fx).astype(int), np.rint(fy).astype(int)
    if not (np.allclose(fx, ix, atol=1e-9) and np.allclose(fy, iy, atol=1e-9)):
        return None
    nx, ny = int(ix.max()) + 1, int(iy.max()) + 1
    if nx > 24 or ny > 16 or len(set(zip(ix.tolist(), iy.tolist()))) != len(ix):
        return None
    order = sorted(range(len(ix)), key=lambda i: (ix[i], iy[i]))
    if order != list(range(len(ix))):
        return None
    # cell centres must reproduce the box centres exactly (the kernel recomputes them from the indices)
    if not (np.array_equal(x0 + (ix + 0.5) * pitch, centres[:, 0]) and np.array_equal(y0 + (iy + 0.5) * pitch, centres[:, 1])):
        return None
    rows = np.zeros(16)
    for a, b in zip(ix, iy):
        rows[b] += float(1 << int(a))
    return dict(x0=x0, y0=y0, pitch=pitch, nx=nx, ny=ny, rows=rows)


def build_consts(M: dict, *, model_kind: int, max_episode_steps: int = 1000, goal_distance_threshold: float = 0.5,
                 collision_threshold: float = 0.15, max_linear_velocity: float = 1.0, max_angular_velocity: float = 1.0,
                 spawn_qpos: Optional[np.ndarray] = None, lidar_index_map: str = "reference",
                 spawn_yaw_range: float = 0.0, spawn_xy_jitter: float = 0.0,
                 tolerance: Optional[float] = None, ls_iterations: Optional[int] = None,
                 ls_fast_cap: int = 1, ls_fast_iters: int = 2, ls_mid_cap: int = 2, ls_mid_iters: int = 4,
                 use_box_grid: bool = True) -> np.ndarray:
    lay = consts_layout()
    total = sum(c for _, c in lay.values())
    blob = np.zeros(total)

    def put(name, val):
        off, cnt = lay[name]
        v = np.asarray(val, float).reshape(-1)
        assert v.size <= cnt, (name, v.size, cnt)
        blob[off:off + v.size] = v

    names = M["body_names"]
    bid = {n: i for i, n in enumerate(names)}
    ch = bid["chassis"]
    wheels = [bid[n] for n in ("rear_left", "rear_right", "front_left", "front_right")]
    steers = [bid[n] for n in ("front_left_steer", "front_right_steer")]
    jn = M["jnt_names"]
    jid = {n: i for i, n in enumerate(jn)}
    hinge_j = [jid["front_left_steer"], jid["front_right_steer"], jid["rear_left_wheel"], jid["rear_right_wheel"],
               jid["front_left_wheel"], jid["front_right_wheel"]]
    hinge_dof = [int(M["jnt_dofadr"][j]) for j in hinge_j]

    # ---- structural checks ----------------------------------------------------------------------
    assert M["nq"] == 13 and M["nv"] == 12, "free chassis + 6 hinges expected"
    assert M["body_jntnum"][ch] == 1 and M["jnt_type"][M["body_jntadr"][ch]] == 0
    fj = M["body_jntadr"][ch]
    assert M["jnt_qposadr"][fj] == 0 and M["jnt_dofadr"][fj] == 0
    assert [int(M["jnt_qposadr"][j]) for j in hinge_j] == [9, 11, 7, 8, 10, 12]
    assert hinge_dof == [8, 10, 6, 7, 9, 11]
    for b in wheels + steers:
        I = M["body_inertia"][b]
        assert np.allclose(I, I[0]) and np.allclose(M["body_ipos"][b], 0), "hinge bodies must be isotropic with COM at origin"
    for j in hinge_j:
        assert np.allclose(M["jnt_pos"][j], 0) and M["jnt_margin"][j] == 0
    for j in hinge_j[:2]:
        assert np.allclose(M["jnt_axis"][j], [0, 0, 1])
    for j in hinge_j[2:]:
        assert np.allclose(M["jnt_axis"][j], [0, 1, 0])
    assert M["body_parentid"][wheels[2]] == steers[0] and M["body_parentid"][wheels[3]] == steers[1]
    for b in (wheels[2], wheels[3]):
        assert np.allclose(M["body_pos"][b], 0) and np.allclose(M["body_quat"][b], [1, 0, 0, 0])
    for b in wheels[:2] + steers:
        assert M["body_parentid"][b] == ch and np.allclose(M["body_quat"][b], [1, 0, 0, 0])
    assert np.allclose(M["dof_damping"][:6], 0) and np.allclose(M["dof_frictionloss"][:6], 0) and np.allclose(M["dof_armature"][:6], 0)
    assert float(M["opt_impratio"][0]) > 0

    h = float(M["opt_timestep"][0])
    put("model_kind", model_kind)
    put("timestep", h)
    put("gravity", M["opt_gravity"])
    put("tolerance", float(M["opt_tolerance"][0]) if tolerance is None else tolerance)
    put("iterations", int(M["opt_iterations"][0]))
    put("ls_iterations", int(M["opt_ls_iterations"][0]) if ls_iterations is None else int(ls_iterations))
    put("ls_fast_cap", ls_fast_cap)
    put("ls_fast_iters", ls_fast_iters)
    put("ls_mid_cap", ls_mid_cap)
    put("ls_mid_iters", ls_mid_iters)
    put("solver_scale", 1.0 / (float(M["stat_meaninertia"][0]) * max(1, M["nv"])))

    # ---- composite inertia about the chassis origin, chassis frame (hinges at 0) -------------------
    qpos = M["qpos0"].copy()
    qpos[0:3] = 0
    qpos[3:7] = [1, 0, 0, 0]
    kin = forward_kinematics(M, qpos)
    sub = [b for b in range(1, M["nbody"]) if M["body_rootid"][b] == ch]
    mass = sum(M["body_mass"][b] for b in sub)
    mcom = sum(M["body_mass"][b] * kin["xipos"][b] for b in sub)
    IO = np.zeros((3, 3))
    for b in sub:
        r = kin["xipos"][b]
        Ib = kin["ximat"][b] @ np.diag(M["body_inertia"][b]) @ kin["ximat"][b].T
        IO += Ib + M["body_mass"][b] * (np.dot(r, r) * np.eye(3) - np.outer(r, r))
    put("mass", mass)
    put("mcom", mcom)
    put("inertiaO", [IO[0, 0], IO[1, 1], IO[2, 2], IO[0, 1], IO[0, 2], IO[1, 2]])

    Jw = [float(M["body_inertia"][b][0]) for b in wheels]
    Js = [float(M["body_inertia"][b][0]) for b in steers]
    put("h_inertia", [Js[0] + Jw[2], Js[1] + Jw[3], Jw[0], Jw[1], Jw[2], Jw[3]])
    put("h_armature", [M["dof_armature"][d] for d in hinge_dof])
    put("h_damping", [M["dof_damping"][d] for d in hinge_dof])
    put("h_floss", [M["dof_frictionloss"][d] for d in hinge_dof])
    flR, flB = [], []
    for d in hinge_dof:
        imp = _impedance0(M["dof_solimp"][d])
        _, B = _KB(M["dof_solref"][d], M["dof_solimp"][d], h)
        flR.append(max(MINVAL, (1 - imp) / imp * M["dof_invweight0"][d]) if M["dof_frictionloss"][d] > 0 else 1.0)
        flB.append(B)
    put("h_flR", flR)
    put("h_flB", flB)
    put("h_flD", [1.0 / r for r in flR])
    put("h_invweight", [M["dof_invweight0"][d] for d in hinge_dof])

    # ---- implicit joint damping: constant part of the 8x8 Schur complement of (M~ + h B) and its inverse ------------
    hin = np.array([Js[0] + Jw[2], Js[1] + Jw[3], Jw[0], Jw[1], Jw[2], Jw[3]])
    arm = np.array([M["dof_armature"][d] for d in hinge_dof])
    dmp = np.array([M["dof_damping"][d] for d in hinge_dof])
    Sc = np.zeros((8, 8))
    Sc[:3, :3] = mass * np.eye(3)
    cx, cy, cz = mcom
    Sc[3:6, :3] = np.array([[0, -cz, cy], [cz, 0, -cx], [-cy, cx, 0]])
    Sc[:3, 3:6] = Sc[3:6, :3].T
    Sc[3:6, 3:6] = IO
    for i in range(2):
        Sc[6 + i, 5] = Sc[5, 6 + i] = hin[i]
        Sc[6 + i, 6 + i] = hin[i] + arm[i] + h * dmp[i]
    cE = hin[2:] + arm[2:] + h * dmp[2:]
    for wI in (0, 1):                               # rear wheels: constant spin axis e_y
        Sc[4, 4] -= hin[2 + wI] ** 2 / cE[wI]
    P = np.linalg.inv(Sc)
    put("eulerP", [P[i, j] for i in range(8) for j in range(i + 1)])
    put("euler_kappa", [hin[4] ** 2 / cE[2], hin[5] ** 2 / cE[3]])
    put("w_cEinv", 1.0 / cE)

    # ---- steer limits -------------------------------------------------------------------------------
    put("st_limited", [M["jnt_limited"][j] for j in hinge_j[:2]])
    put("st_lo", [M["jnt_range"][j][0] for j in hinge_j[:2]])
    put("st_hi", [M["jnt_range"][j][1] for j in hinge_j[:2]])
    KB = [_KB(M["jnt_solref"][j], M["jnt_solimp"][j], h) for j in hinge_j[:2]]
    put("lim_K", [k for k, _ in KB])
    put("lim_B", [b for _, b in KB])
    put("lim_solimp", np.concatenate([M["jnt_solimp"][j] for j in hinge_j[:2]]))
    for j in hinge_j[2:]:
        assert not M["jnt_limited"][j]

    # ---- steering equality ------------------------------------------------------------------------------
    if M["neq"] == 1:
        assert M["eq_obj1id"][0] == hinge_j[0] and M["eq_obj2id"][0] == hinge_j[1]
        assert np.allclose(M["eq_data"][0], [0, 1, 0, 0, 0]), "only polycoef='0 1' (sL - sR = 0) is supported"
        K, B = _KB(M["eq_solref"][0], M["eq_solimp"][0], h)
        put("has_eq", 1)
        put("eq_K", K)
        put("eq_B", B)
        put("eq_solimp", M["eq_solimp"][0])
        put("eq_invweight", M["dof_invweight0"][hinge_dof[0]] + M["dof_invweight0"][hinge_dof[1]])
    else:
        assert M["neq"] == 0

    # ---- wheels and their contacts -------------------------------------------------------------------------
    floor = [g for g in range(M["ngeom"]) if M["geom_type"][g] == 0]
    assert len(floor) == 1
    floor = floor[0]
    assert M["geom_bodyid"][floor] == 0 and np.allclose(M["geom_quat"][floor], [1, 0, 0, 0])
    assert np.allclose(M["geom_pos"][floor][:2], 0)
    put("plane_z", M["geom_pos"][floor][2])
    put("plane_half", M["geom_size"][floor][:2])
    impratio = float(M["opt_impratio"][0])
    boxes = [g for g in range(M["ngeom"]) if M["geom_type"][g] == 6]
    wc, wr, whl, wmu, wmr, wK, wB, wsi, wtr = [], [], [], [], [], [], [], [], []
    bmu, bmr, bK, bB, bsi = [], [], [], [], []
    for b in wheels:
        gs = [g for g in range(M["ngeom"]) if M["geom_bodyid"][g] == b]
        assert len(gs) == 1 and M["geom_type"][gs[0]] == 5
        g = gs[0]
        assert np.allclose(M["geom_pos"][g], 0) and M["geom_condim"][g] == 3 and M["geom_margin"][g] == 0 and M["geom_gap"][g] == 0
        axis = _q2m(M["geom_quat"][g])[:, 2]
        assert np.allclose(np.abs(axis), [0, 1, 0]), "cylinder axis must be the spin axis"
        can = (M["geom_contype"][g] & M["geom_conaffinity"][floor]) or (M["geom_contype"][floor] & M["geom_conaffinity"][g])
        assert can, "wheel must collide with the floor"
        wc.append(kin["xpos"][b])
        wr.append(M["geom_size"][g][0])
        whl.append(M["geom_size"][g][1])
        fr, solref, solimp = _mix(M, floor, g)
        K, B = _KB(solref, solimp, h)
        wmu.append(fr[0]); wmr.append(fr[0] ** 2 / impratio); wK.append(K); wB.append(B); wsi.append(solimp)
        wtr.append(M["body_invweight0"][b][0] + M["body_invweight0"][0][0])
        if boxes:
            fr, solref, solimp = _mix(M, g, boxes[0])
            K, B = _KB(solref, solimp, h)
            bmu.append(fr[0]); bmr.append(fr[0] ** 2 / impratio); bK.append(K); bB.append(B); bsi.append(solimp)
    put("w_center", np.concatenate(wc)); put("w_radius", wr); put("w_halflen", whl)
    put("w_mu", wmu); put("w_mureg2", wmr); put("w_K", wK); put("w_B", wB); put("w_solimp", np.concatenate(wsi)); put("w_tran", wtr)
    if boxes:
        put("wb_mu", bmu); put("wb_mureg2", bmr); put("wb_K", bK); put("wb_B", bB); put("wb_solimp", np.concatenate(bsi))
        hs = M["geom_size"][boxes[0]]
        for g in boxes:
            assert np.allclose(M["geom_size"][g], hs) and np.allclose(M["geom_quat"][g], [1, 0, 0, 0])
            assert M["body_weldid"][M["geom_bodyid"][g]] == 0 and M["geom_pos"][g][2] == M["geom_pos"][boxes[0]][2]
        put("nbox", len(boxes)); put("box_half", hs); put("box_z", M["geom_pos"][boxes[0]][2])
        if len(boxes) <= 40:
            put("box_cx", [M["geom_pos"][g][0] + M["body_pos"][M["geom_bodyid"][g]][0] for g in boxes])
            put("box_cy", [M["geom_pos"][g][1] + M["body_pos"][M["geom_bodyid"][g]][1] for g in boxes])
        grid = _box_grid(np.array([[M["geom_pos"][g][0] + M["body_pos"][M["geom_bodyid"][g]][0],
                                    M["geom_pos"][g][1] + M["body_pos"][M["geom_bodyid"][g]][1]] for g in boxes]), np.asarray(hs, float))
        if grid is not None and use_box_grid:
            put("grid_on", 1); put("grid_x0", grid["x0"]); put("grid_y0", grid["y0"]); put("grid_pitch", grid["pitch"])
            put("grid_nx", grid["nx"]); put("grid_ny", grid["ny"]); put("grid_rows", grid["rows"])
        else:
            assert len(boxes) <= 40, "more than 40 boxes need the occupancy grid (boxes on a lattice, listed in (x, y) order)"

    # ---- chassis plates: hull vertices + hull graph in the chassis frame, contact parameters of the plate pairs
    plates = [g for g in range(M["ngeom"]) if M["geom_type"][g] == 7]
    def _collides(g1, g2):
        return bool((M["geom_contype"][g1] & M["geom_conaffinity"][g2]) or (M["geom_contype"][g2] & M["geom_conaffinity"][g1]))
    pl_floor = [g for g in plates if _collides(g, floor)]
    pl_box = [g for g in plates if boxes and _collides(g, boxes[0])]
    active = [g for g in plates if g in pl_floor or g in pl_box]
    if active:
        assert len(active) == 2 and (not pl_floor or pl_floor == active) and (not pl_box or pl_box == active), "both plates must share their collision masks"
        assert "hull_adj" in M, "model table without hull graph: recompile it (tools/compile_models.py)"
        put("pl_count", 2); put("pl_floor", 1 if pl_floor else 0); put("pl_box", 1 if pl_box else 0)
        verts, adjs, tols, cens, rads, trans, prm = [], [], [], [], [], [], []
        for g in active:
            b = M["geom_bodyid"][g]
            assert M["body_weldid"][b] == ch, "plates must be welded to the chassis"
            Rg = kin["xmat"][b] @ _q2m(M["geom_quat"][g])
            adr, nvv = int(M["geom_hulladr"][g]), int(M["geom_hullnum"][g])
            assert nvv <= 32
            pw = kin["xpos"][b] + (kin["xmat"][b] @ M["geom_pos"][g]) + M["hull_vert"][adr:adr + nvv] @ Rg.T
            v32 = np.zeros((32, 3)); v32[:nvv] = pw
            verts.append(v32)
            packed = np.zeros((32, 6))
            for i in range(nvv):
                lst = [int(x) for x in M["hull_adj"][adr + i] if x >= 0] + [63] * 24
                for w_ in range(6):
                    packed[i, w_] = sum(lst[4 * w_ + q] << (6 * q) for q in range(4))
            adjs.append(packed)
            tols.append(0.3 * float(M["geom_rbound"][g]))
            cen = 0.5 * (pw.min(0) + pw.max(0))
            cens.append(cen); rads.append(float(np.linalg.norm(pw - cen, axis=1).max()))
            trans.append(M["body_invweight0"][b][0] + M["body_invweight0"][0][0])
            for other in ([floor] if pl_floor else []) + ([boxes[0]] if pl_box else []):
                fr, solref, solimp = _mix(M, other, g)
                K, B = _KB(solref, solimp, h)
                prm.append((fr[0], K, B, tuple(solimp)))
                assert M["geom_condim"][g] == 3 and M["geom_margin"][g] == 0 and M["geom_gap"][g] == 0
        assert all(p_ == prm[0] for p_ in prm), "plate-floor and plate-box contacts must share friction / solref / solimp (one parameter set in the kernels)"
        put("pl_nvert", [int(M["geom_hullnum"][g]) for g in active]); put("pl_vert", np.concatenate(verts)); put("pl_adj", np.concatenate(adjs))
        put("pl_tol", tols); put("pl_center", np.concatenate(cens)); put("pl_radius", rads); put("pl_tran", trans)
        put("pl_mu", prm[0][0]); put("pl_mureg2", prm[0][0] ** 2 / impratio); put("pl_K", prm[0][1]); put("pl_B", prm[0][2]); put("pl_solimp", prm[0][3])
    # plate bounding points (cheap "can a plate touch the floor in this pose" test): the 8 corners of the box around both plates
    allv = []
    for g in range(M["ngeom"]):
        if M["geom_type"][g] != 7:
            continue
        if not ((M["geom_contype"][g] & M["geom_conaffinity"][floor]) or (M["geom_contype"][floor] & M["geom_conaffinity"][g])):
            continue
        b = M["geom_bodyid"][g]
        Rg = kin["xmat"][b] @ _q2m(M["geom_quat"][g])
        hv = M["hull_vert"][M["geom_hulladr"][g]:M["geom_hulladr"][g] + M["geom_hullnum"][g]]
        allv.append(kin["xpos"][b] + (kin["xmat"][b] @ M["geom_pos"][g]) + hv @ Rg.T)
    pts = []
    if allv:
        pw = np.concatenate(allv)
        lo, hi = pw.min(0), pw.max(0)
        pts = [[sx, sy, sz] for sx in (lo[0], hi[0]) for sy in (lo[1], hi[1]) for sz in (lo[2], hi[2])]
    put("nhull", len(pts))
    if pts:
        put("hull_pts", np.concatenate(pts))

    # ---- actuators -----------------------------------------------------------------------------------------
    nu = M["nu"]
    assert nu <= 4
    put("nact", nu)
    put("act_hinge", [hinge_j.index(int(j)) for j in M["actuator_trnid"]])
    assert np.allclose(M["actuator_gear"], 1)
    put("act_gain", M["actuator_gainprm"]); put("act_bias", M["actuator_biasprm"])
    put("act_ctrllimited", M["actuator_ctrllimited"]); put("act_ctrlrange", M["actuator_ctrlrange"])
    put("act_forcelimited", M["actuator_forcelimited"]); put("act_forcerange", M["actuator_forcerange"])
    an = M["actuator_names"]
    if an == ["steering_servo", "rear_left_drive", "rear_right_drive"]:
        put("ctrl_kind", 0)   # BicycleController (src/core/controller.py:94-96)
    elif an == ["front_steer_left", "front_steer_right", "rear_left_drive", "rear_right_drive"]:
        put("ctrl_kind", 1)   # AckermannController (src/core/controller.py:37-40)
    else:
        raise ValueError(f"unknown actuator set {an}")

    # ---- lidar ---------------------------------------------------------------------------------------------------
    rf = [i for i in range(len(M["sensor_type"])) if M["sensor_type"][i] == 2]
    nbeam = len(rf)
    assert 0 < nbeam <= 72
    lb = bid["lidar_360"]
    assert np.allclose(M["body_quat"][lb], [1, 0, 0, 0]) and M["body_parentid"][lb] == ch
    cos_, sin_, rad = [], [], []
    for s in rf:
        sid = int(M["sensor_objid"][s])
        assert M["site_bodyid"][sid] == lb
        zax = _q2m(M["site_quat"][sid])[:, 2]
        assert abs(zax[2]) < 1e-9, "beams must be horizontal in the lidar frame"
        cos_.append(zax[0]); sin_.append(zax[1])
        sp = M["site_pos"][sid]
        assert abs(sp[2]) < 1e-12 and abs(sp[0] * zax[1] - sp[1] * zax[0]) < 1e-9, "site offset must lie along the beam"
        rad.append(sp[0] * zax[0] + sp[1] * zax[1])
    assert np.allclose(rad, rad[0])
    cut = M["sensor_cutoff"][rf]
    assert np.allclose(cut, cut[0])
    put("nbeam", nbeam); put("lidar_pos", M["body_pos"][lb]); put("lidar_r", rad[0]); put("lidar_cutoff", cut[0])
    put("lidar_cos", cos_); put("lidar_sin", sin_)
    if lidar_index_map == "reference" and model_kind in (0, 2):   # same robot XML and the same _setup_lidar code in both env classes
        # ackermann_env.py:126-141: mj_name2id("lidar-{i}") is -1 for i < 10 (names are zero padded),
        # and model.sensor_adr[-1] is the last sensor's address  => slots 0..9 all read the last beam.
        sn = M["sensor_names"]
        first = int(M["sensor_adr"][rf[0]])
        amap = []
        for i in range(72):
            name = f"lidar-{i}"
            sidx = sn.index(name) if name in sn else -1
            amap.append(int(M["sensor_adr"][sidx]) - first)
        put("lidar_map", amap)
    else:
        put("lidar_map", np.arange(nbeam))

    # ---- env semantics -----------------------------------------------------------------------------------------------
    put("wheel_radius", 0.0325); put("wheelbase", 0.20); put("track_width", 0.174)  # controller.py:28-29,85-86 defaults
    put("max_linear_velocity", max_linear_velocity); put("max_angular_velocity", max_angular_velocity)
    put("max_episode_steps", max_episode_steps); put("goal_threshold", goal_distance_threshold)
    put("collision_threshold", collision_threshold)
    put("goal_dmin", 2.0); put("goal_dmax", 8.0)   # ackermann_env.py:167
    if spawn_qpos is None:
        spawn_qpos = M["qpos0"].copy()
        if model_kind == 0:
            spawn_qpos[0:3] = [0, 0, 0.1]          # ackermann_env.py:151, simple_map_spawner.py:43-50
            spawn_qpos[3:7] = [1, 0, 0, 0]
        if model_kind == 2:
            spawn_qpos[0:3] = [0, 0, float(M["maze_spawn_z"][0])]   # xy: start cell + noise, drawn per episode
            spawn_qpos[3:7] = [1, 0, 0, 0]
    put("spawn_qpos", spawn_qpos)
    put("qpos0", M["qpos0"])
    put("spawn_yaw_range", spawn_yaw_range); put("spawn_xy_jitter", spawn_xy_jitter)
    if model_kind == 2:
        g = M["maze_grid"]
        assert consts_get(blob, lay, "grid_on") == 1 and np.allclose([consts_get(blob, lay, "grid_x0"), consts_get(blob, lay, "grid_y0"),
                                                                        consts_get(blob, lay, "grid_pitch")], g[:3])
        put("maze_on", 1); put("maze_free_rows", M["maze_free_rows"]); put("maze_xy_noise", float(M["maze_xy_noise"][0]))
        put("settle_steps", int(M["maze_settle_steps"][0]))
    return blob
