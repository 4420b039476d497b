"""Compile-time constants evaluated at qpos0 (dof_invweight0, body_invweight0, meaninertia).

Restates what MuJoCo's model compiler derives once per model at the reference
configuration: the diagonal of M^-1 per dof (averaged over the 3 translational /
3 rotational dofs of a free joint), the mean translational / rotational diagonal
of J_b M^-1 J_b^T with the body-COM Jacobian, and mean(diag M).  These feed the
constraint regulariser R.

The mass matrix here is built from body-COM Jacobians (sum_b m Jp^T Jp + Jr^T I Jr);
this is deliberately a different algorithm from the CRB recursion used by the oracle
and the CUDA kernel, so the three implementations cross-check one another.
"""
from __future__ import annotations

import numpy as np

JNT_FREE, JNT_HINGE = 0, 3


def _qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def _q2m(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def forward_kinematics(M: dict, qpos: np.ndarray) -> dict:
    nb = M["nbody"]
    xpos = np.zeros((nb, 3))
    xquat = np.tile([1.0, 0, 0, 0], (nb, 1))
    xmat = np.tile(np.eye(3), (nb, 1, 1))
    xanchor = np.zeros((M["njnt"], 3))
    xaxis = np.zeros((M["njnt"], 3))
    for b in range(1, nb):
        p = M["body_parentid"][b]
        jn, ja = M["body_jntnum"][b], M["body_jntadr"][b]
        if jn == 1 and M["jnt_type"][ja] == JNT_FREE:
            qa = M["jnt_qposadr"][ja]
            xpos[b] = qpos[qa:qa + 3]
            q = qpos[qa + 3:qa + 7]
            xquat[b] = q / np.linalg.norm(q)
            xanchor[ja] = xpos[b]
            xaxis[ja] = M["jnt_axis"][ja]
        else:
            xpos[b] = xpos[p] + xmat[p] @ M["body_pos"][b]
            xquat[b] = _qmul(xquat[p], M["body_quat"][b])
            for j in range(ja, ja + jn):
                R = _q2m(xquat[b])
                xaxis[j] = R @ M["jnt_axis"][j]
                xanchor[j] = R @ M["jnt_pos"][j] + xpos[b]
                if M["jnt_type"][j] != JNT_HINGE:
                    raise NotImplementedError
                ang = qpos[M["jnt_qposadr"][j]] - M["qpos0"][M["jnt_qposadr"][j]]
                ax = M["jnt_axis"][j]
                ql = np.array([np.cos(ang / 2), *(ax * np.sin(ang / 2))])
                xquat[b] = _qmul(xquat[b], ql)
                xpos[b] = xanchor[j] - _q2m(xquat[b]) @ M["jnt_pos"][j]
        xquat[b] /= np.linalg.norm(xquat[b])
        xmat[b] = _q2m(xquat[b])
    xipos = np.array([xpos[b] + xmat[b] @ M["body_ipos"][b] for b in range(nb)])
    ximat = np.array([_q2m(_qmul(xquat[b], M["body_iquat"][b])) for b in range(nb)])
    return dict(xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, ximat=ximat, xanchor=xanchor, xaxis=xaxis)


def jac_point(M: dict, kin: dict, body: int, point: np.ndarray):
    """(jacp [3,nv], jacr [3,nv]) of a point rigidly attached to `body`."""
    nv = M["nv"]
    jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
    b = body
    while b > 0:
        ja, jn = M["body_jntadr"][b], M["body_jntnum"][b]
        for j in range(ja, ja + jn):
            da = M["jnt_dofadr"][j]
            if M["jnt_type"][j] == JNT_FREE:
                jp[:, da:da + 3] = np.eye(3)
                for k in range(3):
                    ax = kin["xmat"][b][:, k]
                    jr[:, da + 3 + k] = ax
                    jp[:, da + 3 + k] = np.cross(ax, point - kin["xpos"][b])
            else:
                ax = kin["xaxis"][j]
                jr[:, da] = ax
                jp[:, da] = np.cross(ax, point - kin["xanchor"][j])
        b = M["body_parentid"][b]
    return jp, jr


def mass_matrix(M: dict, kin: dict) -> np.ndarray:
    nv = M["nv"]
    Mm = np.zeros((nv, nv))
    for b in range(1, M["nbody"]):
        if M["body_mass"][b] == 0 and not np.any(M["body_inertia"][b]):
            continue
        jp, jr = jac_point(M, kin, b, kin["xipos"][b])
        Iw = kin["ximat"][b] @ np.diag(M["body_inertia"][b]) @ kin["ximat"][b].T
        Mm += M["body_mass"][b] * jp.T @ jp + jr.T @ Iw @ jr
    Mm[np.diag_indices(nv)] += M["dof_armature"]
    return Mm


def set_const(M: dict) -> None:
    nv, nb = M["nv"], M["nbody"]
    kin = forward_kinematics(M, M["qpos0"])
    Mm = mass_matrix(M, kin)
    Minv = np.linalg.inv(Mm)
    M["stat_meaninertia"] = np.array([np.mean(np.diag(Mm))]) if nv else np.array([1.0])
    d = np.diag(Minv).copy()
    dof_inv = np.zeros(nv)
    for j in range(M["njnt"]):
        da = M["jnt_dofadr"][j]
        if M["jnt_type"][j] == JNT_FREE:
            dof_inv[da:da + 3] = d[da:da + 3].mean()
            dof_inv[da + 3:da + 6] = d[da + 3:da + 6].mean()
        else:
            dof_inv[da] = d[da]
    M["dof_invweight0"] = dof_inv
    binv = np.zeros((nb, 2))
    for b in range(1, nb):
        if M["body_weldid"][b] == 0:
            continue
        jp, jr = jac_point(M, kin, b, kin["xipos"][b])
        binv[b, 0] = np.trace(jp @ Minv @ jp.T) / 3
        binv[b, 1] = np.trace(jr @ Minv @ jr.T) / 3
    M["body_invweight0"] = binv
    M["qM0"] = Mm
