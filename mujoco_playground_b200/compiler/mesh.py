"""Binary-STL mesh processing for the model compiler (offline, fp64).

Only what the Ackermann models need from the two chassis plates
(``CAD Models/Base.stl``, ``CAD Models/Ceiling.stl``; referenced from
models/ackermann_robot_v2.xml:8-14): volume, centre of mass and inertia under
MuJoCo's three mesh-inertia conventions, the principal frame, and the convex
hull vertices used by plane-vs-hull collision.
"""
from __future__ import annotations

import struct

import numpy as np


def load_stl(path: str):
    """Binary STL -> (vertices [nv,3] deduplicated, faces [nf,3] int)."""
    with open(path, "rb") as f:
        buf = f.read()
    ntri = struct.unpack("<I", buf[80:84])[0]
    rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("attr", "<u2")])
    tri = np.frombuffer(buf[84:84 + 50 * ntri], dtype=rec)["v"].astype(np.float64)
    flat = tri.reshape(-1, 3)
    verts, inverse = np.unique(flat, axis=0, return_inverse=True)
    faces = inverse.reshape(-1, 3)
    # drop degenerate faces
    ok = (faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 0] != faces[:, 2])
    return verts, faces[ok]


def _tetra_sums(a, b, c, vol):
    """Second-moment integral of tetrahedra (0,a,b,c) with (signed) volumes vol.

    Returns sum_i vol_i * E_i[x x^T] where E_i is the mean over the tetrahedron: for a tetrahedron with
    vertices v0..v3 the mean of x x^T is (sum_k v_k v_k^T + (sum_k v_k)(sum_k v_k)^T) / 20.
    """
    s = a + b + c
    P = (np.einsum("n,ni,nj->ij", vol, a, a) + np.einsum("n,ni,nj->ij", vol, b, b)
         + np.einsum("n,ni,nj->ij", vol, c, c) + np.einsum("n,ni,nj->ij", vol, s, s)) / 20.0
    return P


def _mass_properties(verts, faces, mode: str):
    """Return (volume, com, inertia about com for unit density)."""
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    if mode == "legacy":
        # legacy: tetrahedra from the area-weighted centroid of the surface, |volume| per tetrahedron
        n = np.cross(b - a, c - a)
        area = 0.5 * np.linalg.norm(n, axis=1)
        cen = (a + b + c) / 3.0
        origin = (area[:, None] * cen).sum(0) / area.sum()
    else:
        origin = np.zeros(3)
    a0, b0, c0 = a - origin, b - origin, c - origin
    vol = np.einsum("ni,ni->n", a0, np.cross(b0, c0)) / 6.0
    if mode == "legacy":
        vol = np.abs(vol)
    V = vol.sum()
    com0 = (vol[:, None] * (a0 + b0 + c0) / 4.0).sum(0) / V
    P = _tetra_sums(a0, b0, c0, vol)            # second moment about `origin`
    P = P - V * np.outer(com0, com0)            # about COM
    I = np.trace(P) * np.eye(3) - P
    return V, com0 + origin, I


def process_mesh(path: str, scale, mode: str = "legacy") -> dict:
    verts, faces = load_stl(path)
    verts = verts * np.asarray(scale, float)
    from scipy.spatial import ConvexHull
    hull = ConvexHull(verts)
    hull_idx = np.unique(hull.simplices)
    if mode == "convex":
        # orient hull triangles outward
        hv = verts
        hf = hull.simplices.copy()
        cen = verts[hull_idx].mean(0)
        nrm = np.cross(hv[hf[:, 1]] - hv[hf[:, 0]], hv[hf[:, 2]] - hv[hf[:, 0]])
        flip = np.einsum("ni,ni->n", nrm, hv[hf[:, 0]] - cen) < 0
        hf[flip] = hf[flip][:, ::-1]
        V, com, I = _mass_properties(hv, hf, "exact")
    elif mode in ("exact", "legacy"):
        V, com, I = _mass_properties(verts, faces, mode)
    else:
        raise ValueError(f"mesh inertia mode {mode!r}")
    w, U = np.linalg.eigh(I)
    order = np.argsort(-w)
    w, U = w[order], U[:, order]
    if np.linalg.det(U) < 0:
        U[:, 2] = -U[:, 2]
    from .mjcf import mat_to_quat
    quat = mat_to_quat(U)
    local = (verts - com) @ U          # vertices in COM-centred principal frame
    # hull graph as MuJoCo builds it (user_mesh.cc MakeGraph): walk the triangulated hull faces in qhull's facet order; every vertex
    # of a face gets the face's other two vertices appended to its edge list unless already there.  mjc_PlaneConvex walks these
    # lists in order.  (qhull "Qt" through scipy; the facet order of MuJoCo's own qhull build may differ: VERIFY with hull_graph
    # of tools/dump_mjmodel.py.)
    loc = {int(g): i for i, g in enumerate(hull_idx)}
    adj = [[] for _ in hull_idx]
    for tri in hull.simplices:
        ids = [loc[int(v)] for v in tri]
        for a in ids:
            for b in ids:
                if b != a and b not in adj[a]:
                    adj[a].append(b)
    max_adj = 24
    assert max(len(x) for x in adj) <= max_adj, "hull vertex degree above the table width"
    hull_adj = -np.ones((len(hull_idx), max_adj), np.int32)
    for i, lst in enumerate(adj):
        hull_adj[i, :len(lst)] = lst
    return dict(hull_adj_local=hull_adj, volume=float(V), com=com, inertia_unit_density=I, quat=quat, principal=w,
                hull_vert_local=local[hull_idx].copy(), aabb_half=np.abs(local).max(0),
                nvert=len(verts), nface=len(faces))
