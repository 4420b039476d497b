"""MJCF subset compiler: XML (+ binary STL) -> table-form rigid-body model.

This is host-side, offline, fp64 NumPy code.  It understands exactly the MJCF
features used by the two models on the hot path of ulusoyn/mujoco_playground:

  * models/ackermann_robot_v2.xml                (reference: loaded at
    src/rl/envs/simple_map_spawner.py:37)
  * models/environments/ackermann_maze_flat.xml  (obstacle scene)

The output is a plain dict of NumPy arrays that uses MuJoCo's field naming
(body_parentid, jnt_axis, dof_damping, geom_size ...), so that the CPU oracle
(oracle/ackb_oracle.c) can be table driven and so that a dump of a real
``mujoco.MjModel`` (tools/dump_mjmodel.py) can override any entry.

Nothing here is executed on the product path at step time: the CUDA kernels
consume the constants extracted from this table by ``constants.py``.
"""
from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional

import numpy as np

from . import mesh as meshlib

# MuJoCo enum values (documented API constants)
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
_GEOM_TYPES = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE,
               "ellipsoid": GEOM_ELLIPSOID, "cylinder": GEOM_CYLINDER, "box": GEOM_BOX,
               "mesh": GEOM_MESH}

MJ_MINVAL = 1e-15


# --------------------------------------------------------------------------- #
# small quaternion / rotation helpers (w, x, y, z convention)
# --------------------------------------------------------------------------- #
def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_from_axis_angle(axis, ang):
    axis = np.asarray(axis, float)
    s = math.sin(ang * 0.5)
    return np.array([math.cos(ang * 0.5), axis[0] * s, axis[1] * s, axis[2] * s])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat_to_quat(R):
    """Rotation matrix -> unit quaternion (w >= 0)."""
    R = np.asarray(R, float)
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    if q[0] < 0:
        q = -q
    return q / np.linalg.norm(q)


def quat_from_euler_xyz(e):
    """Intrinsic x-y-z Euler angles (MJCF default eulerseq="xyz"), radians."""
    q = np.array([1.0, 0, 0, 0])
    for ax, a in zip(np.eye(3), e):
        q = quat_mul(q, quat_from_axis_angle(ax, a))
    return q


def quat_from_zaxis(v):
    """Minimal rotation taking (0,0,1) onto v (MJCF ``zaxis`` attribute)."""
    v = np.asarray(v, float)
    v = v / np.linalg.norm(v)
    axis = np.cross([0, 0, 1.0], v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        axis = np.array([1.0, 0, 0])
    else:
        axis = axis / s
    return quat_from_axis_angle(axis, math.atan2(s, v[2]))


def _floats(s: Optional[str], n: Optional[int] = None, default=None):
    if s is None:
        return None if default is None else np.array(default, float)
    a = np.array([float(t) for t in s.split()], float)
    if n is not None and len(a) < n and default is not None:
        d = np.array(default, float)
        d[: len(a)] = a
        a = d
    return a


# --------------------------------------------------------------------------- #
class _Ctx:
    def __init__(self, xml_path: str, mesh_inertia: str):
        self.xml_path = xml_path
        self.dir = os.path.dirname(os.path.abspath(xml_path))
        self.mesh_inertia = mesh_inertia
        self.deg = False
        self.defaults: Dict[str, Dict[str, str]] = {}
        self.meshes: Dict[str, dict] = {}
        self.bodies: List[dict] = []
        self.joints: List[dict] = []
        self.geoms: List[dict] = []
        self.sites: List[dict] = []

    def ang(self, a):
        return np.deg2rad(a) if self.deg else a

    def attrs(self, el) -> Dict[str, str]:
        d = dict(self.defaults.get(el.tag, {}))
        d.update(el.attrib)
        return d

    def frame_quat(self, a: Dict[str, str]):
        if "quat" in a:
            q = _floats(a["quat"])
            return q / np.linalg.norm(q)
        if "euler" in a:
            return quat_from_euler_xyz(self.ang(_floats(a["euler"])))
        if "zaxis" in a:
            return quat_from_zaxis(_floats(a["zaxis"]))
        if "axisangle" in a:
            v = _floats(a["axisangle"])
            return quat_from_axis_angle(v[:3] / np.linalg.norm(v[:3]), self.ang(v[3]))
        return np.array([1.0, 0, 0, 0])


def _expand_includes(root: ET.Element, base_dir: str) -> None:
    """<include file=.../> : splice the included file's top-level children in place."""
    for parent in list(root.iter()):
        for i, ch in enumerate(list(parent)):
            if ch.tag == "include":
                inc = ET.parse(os.path.join(base_dir, ch.attrib["file"])).getroot()
                _expand_includes(inc, base_dir)
                parent.remove(ch)
                for k, sub in enumerate(list(inc)):
                    parent.insert(i + k, sub)


def _merge_sections(root: ET.Element, tag: str) -> List[ET.Element]:
    out = []
    for sec in root.findall(tag):
        out.extend(list(sec))
    return out


def _parse_body(ctx: _Ctx, el: ET.Element, parent_id: int, rep_pos=None, rep_quat=None):
    """Depth-first traversal; ids follow XML order exactly as MuJoCo assigns them."""
    a = ctx.attrs(el)
    bid = len(ctx.bodies)
    body = dict(name=a.get("name", ""), parent=parent_id,
                pos=_floats(a.get("pos"), 3, [0, 0, 0]), quat=ctx.frame_quat(a),
                joints=[], geoms=[], inertial=None)
    ctx.bodies.append(body)
    _parse_body_children(ctx, el, bid, np.zeros(3), np.array([1.0, 0, 0, 0]), "")
    return bid


def _parse_body_children(ctx: _Ctx, el: ET.Element, bid: int, fpos, fquat, suffix: str):
    """fpos/fquat: accumulated <replicate> frame applied to direct child elements."""
    body = ctx.bodies[bid]
    for ch in el:
        a = ctx.attrs(ch)
        if ch.tag == "freejoint" or ch.tag == "joint":
            jtype = JNT_FREE if ch.tag == "freejoint" else \
                {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}[a.get("type", "hinge")]
            rng = _floats(a.get("range"))
            limited = a.get("limited", "auto")
            lim = (rng is not None) if limited == "auto" else (limited == "true")
            if rng is None:
                rng = np.zeros(2)
            elif jtype == JNT_HINGE:
                rng = ctx.ang(rng)
            axis = _floats(a.get("axis"), 3, [0, 0, 1])
            axis = axis / np.linalg.norm(axis)
            j = dict(name=a.get("name", "") + suffix, type=jtype, body=bid,
                     pos=_floats(a.get("pos"), 3, [0, 0, 0]), axis=axis,
                     range=rng, limited=lim,
                     damping=float(a.get("damping", 0)), frictionloss=float(a.get("frictionloss", 0)),
                     armature=float(a.get("armature", 0)), margin=float(a.get("margin", 0)),
                     ref=float(a.get("ref", 0)),
                     solreflimit=_floats(a.get("solreflimit"), 2, [0.02, 1]),
                     solimplimit=_floats(a.get("solimplimit"), 5, [0.9, 0.95, 0.001, 0.5, 2]),
                     solreffriction=_floats(a.get("solreffriction"), 2, [0.02, 1]),
                     solimpfriction=_floats(a.get("solimpfriction"), 5, [0.9, 0.95, 0.001, 0.5, 2]))
            body["joints"].append(len(ctx.joints))
            ctx.joints.append(j)
        elif ch.tag == "inertial":
            full = _floats(a.get("fullinertia"))
            if full is not None:
                raise NotImplementedError("fullinertia")
            body["inertial"] = dict(pos=_floats(a.get("pos"), 3, [0, 0, 0]), quat=ctx.frame_quat(a),
                                    mass=float(a["mass"]), diag=_floats(a.get("diaginertia"), 3, [0, 0, 0]))
        elif ch.tag == "geom":
            gtype = _GEOM_TYPES[a.get("type", "sphere")]
            pos = _floats(a.get("pos"), 3, [0, 0, 0])
            quat = ctx.frame_quat(a)
            pos = fpos + quat_to_mat(fquat) @ pos
            quat = quat_mul(fquat, quat)
            rgba = _floats(a.get("rgba"), 4, [0.5, 0.5, 0.5, 1])
            g = dict(name=a.get("name", "") + (suffix if a.get("name") else ""), type=gtype, body=bid,
                     size=_floats(a.get("size"), 3, [0, 0, 0]), pos=pos, quat=quat,
                     contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
                     condim=int(a.get("condim", 3)), group=int(a.get("group", 0)),
                     friction=_floats(a.get("friction"), 3, [1, 0.005, 0.0001]),
                     solref=_floats(a.get("solref"), 2, [0.02, 1]),
                     solimp=_floats(a.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2]),
                     solmix=float(a.get("solmix", 1)), margin=float(a.get("margin", 0)),
                     gap=float(a.get("gap", 0)), priority=int(a.get("priority", 0)),
                     mass=(float(a["mass"]) if "mass" in a else None),
                     density=float(a.get("density", 1000)), alpha=float(rgba[3]),
                     mesh=a.get("mesh"))
            body["geoms"].append(len(ctx.geoms))
            ctx.geoms.append(g)
        elif ch.tag == "site":
            pos = _floats(a.get("pos"), 3, [0, 0, 0])
            quat = ctx.frame_quat(a)
            pos = fpos + quat_to_mat(fquat) @ pos
            quat = quat_mul(fquat, quat)
            ctx.sites.append(dict(name=a.get("name", "") + suffix, base=a.get("name", ""), body=bid,
                                  pos=pos, quat=quat))
        elif ch.tag == "replicate":
            # MJCF <replicate count sep offset euler>: copy i is placed in the frame obtained by
            # applying the (offset, euler) transform i times; names get "<sep><zero padded i>".
            count = int(a["count"])
            sep = a.get("sep", "")
            off = _floats(a.get("offset"), 3, [0, 0, 0])
            dq = quat_from_euler_xyz(ctx.ang(_floats(a.get("euler"), 3, [0, 0, 0])))
            width = len(str(count - 1))
            p, q = fpos.copy(), fquat.copy()
            for i in range(count):
                _parse_body_children(ctx, ch, bid, p, q, suffix + f"{sep}{i:0{width}d}")
                p = p + quat_to_mat(q) @ off
                q = quat_mul(q, dq)
        elif ch.tag == "body":
            if np.any(fpos != 0) or fquat[0] != 1.0:
                raise NotImplementedError("replicated bodies")
            _parse_body(ctx, ch, bid)
        elif ch.tag in ("light", "camera"):
            pass
        else:
            raise NotImplementedError(f"MJCF element <{ch.tag}> inside <body>")


def compile_mjcf(xml_path: str, mesh_inertia: str = "legacy", root: Optional[ET.Element] = None) -> dict:
    """Compile the MJCF file into a MuJoCo-style table model (dict of ndarrays).

    mesh_inertia: "legacy" | "exact" | "convex"  (MuJoCo mesh ``inertia`` modes; the default of the
    MuJoCo 3.x releases contemporary with the reference's checkpoints is "legacy").
    """
    if root is None:           # `root`: an already parsed / edited tree; xml_path then only anchors relative mesh paths
        root = ET.parse(xml_path).getroot()
    ctx = _Ctx(xml_path, mesh_inertia)
    _expand_includes(root, ctx.dir)

    comp = root.find("compiler")
    if comp is not None:
        ctx.deg = comp.attrib.get("angle", "degree") == "degree"
    else:
        ctx.deg = True
    for c in root.findall("compiler"):
        ctx.deg = c.attrib.get("angle", "degree" if ctx.deg else "radian") == "degree"

    opt = dict(timestep=0.002, gravity=np.array([0, 0, -9.81]), impratio=1.0, tolerance=1e-8,
               ls_tolerance=0.01, iterations=100, ls_iterations=50)
    for o in root.findall("option"):
        if "timestep" in o.attrib:
            opt["timestep"] = float(o.attrib["timestep"])
        if "gravity" in o.attrib:
            opt["gravity"] = _floats(o.attrib["gravity"])
        for k in ("impratio", "tolerance", "ls_tolerance"):
            if k in o.attrib:
                opt[k] = float(o.attrib[k])
        for k in ("iterations", "ls_iterations"):
            if k in o.attrib:
                opt[k] = int(o.attrib[k])
        for k in ("integrator", "cone", "solver"):
            if k in o.attrib and o.attrib[k] not in ("Euler", "pyramidal", "Newton"):
                raise NotImplementedError(f"option {k}={o.attrib[k]}")

    for d in root.findall("default"):
        for el in d:
            if el.tag == "default":
                raise NotImplementedError("nested default classes")
            ctx.defaults.setdefault(el.tag, {}).update(el.attrib)

    for m in _merge_sections(root, "asset"):
        if m.tag == "mesh":
            path = os.path.normpath(os.path.join(ctx.dir, m.attrib["file"]))
            if not os.path.exists(path):
                # models/environments/ackermann_maze_flat.xml:12,16 keeps the "../CAD Models/" path of the
                # file it was copied from (one directory too shallow): retry one level further up.
                alt = os.path.normpath(os.path.join(ctx.dir, "..", m.attrib["file"]))
                if os.path.exists(alt):
                    path = alt
            name = m.attrib.get("name", os.path.splitext(os.path.basename(path))[0])
            scale = _floats(m.attrib.get("scale"), 3, [1, 1, 1])
            ctx.meshes[name] = meshlib.process_mesh(path, scale, mesh_inertia)

    # world body
    ctx.bodies.append(dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                           joints=[], geoms=[], inertial=None))
    for wb in root.findall("worldbody"):
        _parse_body_children(ctx, wb, 0, np.zeros(3), np.array([1.0, 0, 0, 0]), "")

    return _assemble(ctx, root, opt)


# --------------------------------------------------------------------------- #
def _geom_mass_inertia(ctx: _Ctx, g: dict):
    """Return (mass, com offset in geom frame (pre-adjust), inertia 3x3 about COM in geom frame)."""
    t = g["type"]
    s = g["size"]
    if t == GEOM_MESH:
        me = ctx.meshes[g["mesh"]]
        vol, I_unit = me["volume"], me["inertia_unit_density"]  # about COM, mesh axes
        mass = g["mass"] if g["mass"] is not None else g["density"] * vol
        return mass, me["com"], I_unit * (mass / vol)
    if t == GEOM_CYLINDER:
        r, h = s[0], s[1]
        vol = math.pi * r * r * 2 * h
        mass = g["mass"] if g["mass"] is not None else g["density"] * vol
        ixy = mass * (3 * r * r + (2 * h) ** 2) / 12
        return mass, np.zeros(3), np.diag([ixy, ixy, mass * r * r / 2])
    if t == GEOM_BOX:
        vol = 8 * s[0] * s[1] * s[2]
        mass = g["mass"] if g["mass"] is not None else g["density"] * vol
        return mass, np.zeros(3), np.diag([mass * (s[1] ** 2 + s[2] ** 2) / 3,
                                           mass * (s[0] ** 2 + s[2] ** 2) / 3,
                                           mass * (s[0] ** 2 + s[1] ** 2) / 3])
    if t == GEOM_SPHERE:
        vol = 4 / 3 * math.pi * s[0] ** 3
        mass = g["mass"] if g["mass"] is not None else g["density"] * vol
        return mass, np.zeros(3), np.eye(3) * (0.4 * mass * s[0] ** 2)
    if t == GEOM_PLANE:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    raise NotImplementedError(f"inertia of geom type {t}")


def _assemble(ctx: _Ctx, root: ET.Element, opt: dict) -> dict:
    nb = len(ctx.bodies)
    M: dict = {}
    M["opt_timestep"] = np.array([opt["timestep"]])
    M["opt_gravity"] = np.asarray(opt["gravity"], float)
    M["opt_impratio"] = np.array([opt["impratio"]])
    M["opt_tolerance"] = np.array([opt["tolerance"]])
    M["opt_ls_tolerance"] = np.array([opt["ls_tolerance"]])
    M["opt_iterations"] = np.array([opt["iterations"]], np.int32)
    M["opt_ls_iterations"] = np.array([opt["ls_iterations"]], np.int32)

    # ---- mesh geoms: shift geom frame to mesh COM / principal axes like MuJoCo does ----------
    for g in ctx.geoms:
        if g["type"] == GEOM_MESH:
            me = ctx.meshes[g["mesh"]]
            Rg = quat_to_mat(g["quat"])
            g["user_pos"], g["user_quat"] = g["pos"].copy(), g["quat"].copy()
            g["pos"] = g["pos"] + Rg @ me["com"]
            g["quat"] = quat_mul(g["quat"], me["quat"])
            g["size"] = me["aabb_half"].copy()

    # ---- joints / dofs ------------------------------------------------------------------------
    njnt = len(ctx.joints)
    jnt_qposadr, jnt_dofadr = [], []
    nq = nv = 0
    for j in ctx.joints:
        jnt_qposadr.append(nq)
        jnt_dofadr.append(nv)
        nq += {JNT_FREE: 7, JNT_BALL: 4, JNT_SLIDE: 1, JNT_HINGE: 1}[j["type"]]
        nv += {JNT_FREE: 6, JNT_BALL: 3, JNT_SLIDE: 1, JNT_HINGE: 1}[j["type"]]
    M["nq"], M["nv"], M["nbody"], M["njnt"] = nq, nv, nb, njnt
    M["ngeom"], M["nsite"] = len(ctx.geoms), len(ctx.sites)

    body_parentid = np.array([b["parent"] for b in ctx.bodies], np.int32)
    body_jntnum = np.array([len(b["joints"]) for b in ctx.bodies], np.int32)
    body_jntadr = np.array([b["joints"][0] if b["joints"] else -1 for b in ctx.bodies], np.int32)
    body_dofnum = np.zeros(nb, np.int32)
    body_dofadr = -np.ones(nb, np.int32)
    for bi, b in enumerate(ctx.bodies):
        for jid in b["joints"]:
            if body_dofadr[bi] < 0:
                body_dofadr[bi] = jnt_dofadr[jid]
            body_dofnum[bi] += {JNT_FREE: 6, JNT_BALL: 3}.get(ctx.joints[jid]["type"], 1)
    body_rootid = np.zeros(nb, np.int32)
    body_weldid = np.zeros(nb, np.int32)
    for bi in range(1, nb):
        p = body_parentid[bi]
        body_rootid[bi] = bi if p == 0 else body_rootid[p]
        body_weldid[bi] = bi if body_jntnum[bi] > 0 else body_weldid[p]
    M.update(body_parentid=body_parentid, body_jntnum=body_jntnum, body_jntadr=body_jntadr,
             body_dofnum=body_dofnum, body_dofadr=body_dofadr, body_rootid=body_rootid,
             body_weldid=body_weldid)
    M["body_pos"] = np.array([b["pos"] for b in ctx.bodies])
    M["body_quat"] = np.array([b["quat"] for b in ctx.bodies])
    M["body_names"] = [b["name"] for b in ctx.bodies]

    # ---- body inertial frames -----------------------------------------------------------------
    body_mass = np.zeros(nb)
    body_ipos = np.zeros((nb, 3))
    body_iquat = np.tile([1.0, 0, 0, 0], (nb, 1))
    body_inertia = np.zeros((nb, 3))
    for bi, b in enumerate(ctx.bodies):
        if bi == 0:
            continue
        if b["inertial"] is not None:
            it = b["inertial"]
            body_mass[bi], body_ipos[bi], body_iquat[bi], body_inertia[bi] = it["mass"], it["pos"], it["quat"], it["diag"]
            continue
        if body_weldid[bi] == 0:
            continue  # static body: inertia irrelevant
        # inertiafromgeom="auto": accumulate geoms (COM-weighted, parallel axis) in body frame
        ms, coms, Is = [], [], []
        for gid in b["geoms"]:
            g = ctx.geoms[gid]
            m_, c_, I_ = _geom_mass_inertia(ctx, g)
            if g["type"] == GEOM_MESH:
                R = quat_to_mat(g["user_quat"])
                c_w = g["user_pos"] + R @ c_
            else:
                R = quat_to_mat(g["quat"])
                c_w = g["pos"] + R @ c_
            ms.append(m_)
            coms.append(c_w)
            Is.append(R @ I_ @ R.T)
        mt = float(sum(ms))
        if mt <= 0:
            continue
        com = sum(m_ * c for m_, c in zip(ms, coms)) / mt
        I = np.zeros((3, 3))
        for m_, c, I_ in zip(ms, coms, Is):
            d = c - com
            I += I_ + m_ * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        w, V = np.linalg.eigh(I)
        order = np.argsort(-w)  # MuJoCo sorts principal moments in decreasing order
        w, V = w[order], V[:, order]
        if np.linalg.det(V) < 0:
            V[:, 2] = -V[:, 2]
        body_mass[bi], body_ipos[bi], body_iquat[bi], body_inertia[bi] = mt, com, mat_to_quat(V), w
    M.update(body_mass=body_mass, body_ipos=body_ipos, body_iquat=body_iquat, body_inertia=body_inertia)

    # ---- joints -------------------------------------------------------------------------------
    J = ctx.joints
    M["jnt_type"] = np.array([j["type"] for j in J], np.int32)
    M["jnt_qposadr"] = np.array(jnt_qposadr, np.int32)
    M["jnt_dofadr"] = np.array(jnt_dofadr, np.int32)
    M["jnt_bodyid"] = np.array([j["body"] for j in J], np.int32)
    M["jnt_pos"] = np.array([j["pos"] for j in J]).reshape(njnt, 3)
    M["jnt_axis"] = np.array([j["axis"] for j in J]).reshape(njnt, 3)
    M["jnt_limited"] = np.array([int(j["limited"]) for j in J], np.int32)
    M["jnt_range"] = np.array([j["range"] for j in J]).reshape(njnt, 2)
    M["jnt_margin"] = np.array([j["margin"] for j in J])
    M["jnt_solref"] = np.array([j["solreflimit"] for j in J]).reshape(njnt, 2)
    M["jnt_solimp"] = np.array([j["solimplimit"] for j in J]).reshape(njnt, 5)
    M["jnt_names"] = [j["name"] for j in J]

    qpos0 = np.zeros(nq)
    dof_bodyid = np.zeros(nv, np.int32)
    dof_jntid = np.zeros(nv, np.int32)
    dof_parentid = -np.ones(nv, np.int32)
    dof_armature = np.zeros(nv)
    dof_damping = np.zeros(nv)
    dof_frictionloss = np.zeros(nv)
    dof_solref = np.zeros((nv, 2))
    dof_solimp = np.zeros((nv, 5))
    last_dof_of_body = -np.ones(nb, np.int32)
    for jid, j in enumerate(J):
        qa, da = jnt_qposadr[jid], jnt_dofadr[jid]
        b = j["body"]
        if j["type"] == JNT_FREE:
            qpos0[qa:qa + 3] = ctx.bodies[b]["pos"]
            qpos0[qa + 3:qa + 7] = ctx.bodies[b]["quat"]
            n = 6
        elif j["type"] == JNT_HINGE or j["type"] == JNT_SLIDE:
            qpos0[qa] = j["ref"]
            n = 1
        else:
            raise NotImplementedError("ball joint")
        for k in range(n):
            d = da + k
            dof_bodyid[d], dof_jntid[d] = b, jid
            dof_armature[d], dof_damping[d], dof_frictionloss[d] = j["armature"], j["damping"], j["frictionloss"]
            dof_solref[d], dof_solimp[d] = j["solreffriction"], j["solimpfriction"]
            if last_dof_of_body[b] >= 0:
                dof_parentid[d] = last_dof_of_body[b]
            else:
                p = body_parentid[b]
                while p > 0 and last_dof_of_body[p] < 0:
                    p = body_parentid[p]
                dof_parentid[d] = last_dof_of_body[p] if p > 0 else -1
            last_dof_of_body[b] = d
    M.update(qpos0=qpos0, dof_bodyid=dof_bodyid, dof_jntid=dof_jntid, dof_parentid=dof_parentid,
             dof_armature=dof_armature, dof_damping=dof_damping, dof_frictionloss=dof_frictionloss,
             dof_solref=dof_solref, dof_solimp=dof_solimp)

    # ---- geoms / sites ------------------------------------------------------------------------
    G = ctx.geoms
    ng = len(G)
    M["geom_type"] = np.array([g["type"] for g in G], np.int32)
    M["geom_bodyid"] = np.array([g["body"] for g in G], np.int32)
    M["geom_contype"] = np.array([g["contype"] for g in G], np.int32)
    M["geom_conaffinity"] = np.array([g["conaffinity"] for g in G], np.int32)
    M["geom_condim"] = np.array([g["condim"] for g in G], np.int32)
    M["geom_priority"] = np.array([g["priority"] for g in G], np.int32)
    M["geom_size"] = np.array([g["size"] for g in G]).reshape(ng, 3)
    M["geom_pos"] = np.array([g["pos"] for g in G]).reshape(ng, 3)
    M["geom_quat"] = np.array([g["quat"] for g in G]).reshape(ng, 4)
    M["geom_friction"] = np.array([g["friction"] for g in G]).reshape(ng, 3)
    M["geom_solref"] = np.array([g["solref"] for g in G]).reshape(ng, 2)
    M["geom_solimp"] = np.array([g["solimp"] for g in G]).reshape(ng, 5)
    M["geom_solmix"] = np.array([g["solmix"] for g in G])
    M["geom_margin"] = np.array([g["margin"] for g in G])
    M["geom_gap"] = np.array([g["gap"] for g in G])
    M["geom_alpha"] = np.array([g["alpha"] for g in G])
    M["geom_names"] = [g["name"] for g in G]
    # convex-hull vertices of mesh geoms, expressed in the (COM-centred, principal) geom frame
    hull_adr = -np.ones(ng, np.int32)
    hull_num = np.zeros(ng, np.int32)
    hull_vert: List[np.ndarray] = []
    hull_adj: List[np.ndarray] = []
    rbound = np.zeros(ng)
    n_h = 0
    for gi, g in enumerate(G):
        if g["type"] == GEOM_MESH:
            me = ctx.meshes[g["mesh"]]
            hv = me["hull_vert_local"]
            hull_adr[gi], hull_num[gi] = n_h, len(hv)
            hull_vert.append(hv)
            hull_adj.append(me["hull_adj_local"])         # neighbour lists (indices local to this geom's hull), -1 padded
            rbound[gi] = float(np.linalg.norm(me["aabb_half"]))   # mjCGeom::GetRBound for meshes: norm of the half sizes of the aligned box
            n_h += len(hv)
    M["geom_hulladr"], M["geom_hullnum"] = hull_adr, hull_num
    M["hull_vert"] = np.concatenate(hull_vert) if hull_vert else np.zeros((0, 3))
    M["hull_adj"] = np.concatenate(hull_adj) if hull_adj else -np.ones((0, 24), np.int32)
    M["geom_rbound"] = rbound

    S = ctx.sites
    M["site_bodyid"] = np.array([s["body"] for s in S], np.int32)
    M["site_pos"] = np.array([s["pos"] for s in S]).reshape(len(S), 3)
    M["site_quat"] = np.array([s["quat"] for s in S]).reshape(len(S), 4)
    M["site_names"] = [s["name"] for s in S]

    # ---- sensors (sensor list expands over replicated sites, names get the same suffix) --------
    sens_names, sens_type, sens_obj, sens_cutoff, sens_adr = [], [], [], [], []
    adr = 0
    jn = M["jnt_names"]
    for s in _merge_sections(root, "sensor"):
        a = s.attrib
        cutoff = float(a.get("cutoff", 0))
        if s.tag in ("jointpos", "jointvel"):
            sens_names.append(a.get("name", ""))
            sens_type.append(0 if s.tag == "jointpos" else 1)
            sens_obj.append(jn.index(a["joint"]))
            sens_cutoff.append(cutoff)
            sens_adr.append(adr)
            adr += 1
        elif s.tag == "rangefinder":
            matches = [i for i, si in enumerate(S) if si["name"] == a["site"]]
            if matches:
                targets = [(a.get("name", ""), matches[0])]
            else:  # replicated site: one sensor per copy, suffix carried over to the sensor name
                targets = [(a.get("name", "") + si["name"][len(si["base"]):], i)
                           for i, si in enumerate(S) if si["base"] == a["site"]]
            for nm, sid in targets:
                sens_names.append(nm)
                sens_type.append(2)
                sens_obj.append(sid)
                sens_cutoff.append(cutoff)
                sens_adr.append(adr)
                adr += 1
        else:
            raise NotImplementedError(f"sensor <{s.tag}>")
    M["sensor_names"] = sens_names
    M["sensor_type"] = np.array(sens_type, np.int32)      # 0 jointpos, 1 jointvel, 2 rangefinder
    M["sensor_objid"] = np.array(sens_obj, np.int32)
    M["sensor_cutoff"] = np.array(sens_cutoff)
    M["sensor_adr"] = np.array(sens_adr, np.int32)
    M["nsensordata"] = adr

    # ---- equality -----------------------------------------------------------------------------
    eqs = []
    for e in _merge_sections(root, "equality"):
        if e.tag != "joint":
            raise NotImplementedError(f"equality <{e.tag}>")
        a = e.attrib
        eqs.append(dict(j1=jn.index(a["joint1"]), j2=(jn.index(a["joint2"]) if "joint2" in a else -1),
                        poly=_floats(a.get("polycoef"), 5, [0, 1, 0, 0, 0]),
                        solref=_floats(a.get("solref"), 2, [0.02, 1]),
                        solimp=_floats(a.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2])))
    M["neq"] = len(eqs)
    M["eq_obj1id"] = np.array([e["j1"] for e in eqs], np.int32)
    M["eq_obj2id"] = np.array([e["j2"] for e in eqs], np.int32)
    M["eq_data"] = np.array([e["poly"] for e in eqs]).reshape(len(eqs), 5)
    M["eq_solref"] = np.array([e["solref"] for e in eqs]).reshape(len(eqs), 2)
    M["eq_solimp"] = np.array([e["solimp"] for e in eqs]).reshape(len(eqs), 5)

    # ---- actuators ----------------------------------------------------------------------------
    acts = []
    for e in _merge_sections(root, "actuator"):
        a = e.attrib
        gain, bias = 1.0, np.zeros(3)
        if e.tag == "position":
            kp, kv = float(a.get("kp", 1)), float(a.get("kv", 0))
            gain, bias = kp, np.array([0, -kp, -kv])
        elif e.tag == "velocity":
            kv = float(a.get("kv", 1))
            gain, bias = kv, np.array([0, 0, -kv])
        elif e.tag == "motor":
            pass
        else:
            raise NotImplementedError(f"actuator <{e.tag}>")
        cr, fr = _floats(a.get("ctrlrange")), _floats(a.get("forcerange"))
        acts.append(dict(name=a.get("name", ""), jnt=jn.index(a["joint"]), gear=float(a.get("gear", "1").split()[0]),
                         gain=gain, bias=bias,
                         ctrllimited=cr is not None, ctrlrange=cr if cr is not None else np.zeros(2),
                         forcelimited=fr is not None, forcerange=fr if fr is not None else np.zeros(2)))
    nu = len(acts)
    M["nu"] = nu
    M["actuator_names"] = [a["name"] for a in acts]
    M["actuator_trnid"] = np.array([a["jnt"] for a in acts], np.int32)
    M["actuator_gear"] = np.array([a["gear"] for a in acts])
    M["actuator_gainprm"] = np.array([a["gain"] for a in acts])
    M["actuator_biasprm"] = np.array([a["bias"] for a in acts]).reshape(nu, 3)
    M["actuator_ctrllimited"] = np.array([int(a["ctrllimited"]) for a in acts], np.int32)
    M["actuator_ctrlrange"] = np.array([a["ctrlrange"] for a in acts]).reshape(nu, 2)
    M["actuator_forcelimited"] = np.array([int(a["forcelimited"]) for a in acts], np.int32)
    M["actuator_forcerange"] = np.array([a["forcerange"] for a in acts]).reshape(nu, 2)

    from .setconst import set_const
    set_const(M)
    return M
