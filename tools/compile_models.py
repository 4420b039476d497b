#!/usr/bin/env python
"""Compile the reference's MJCF models into the table-form fixtures shipped with the package.

Runs wherever the reference checkout is available (it is NOT available on the GPU box):
    python tools/compile_models.py [/root/reference] [--mesh-inertia legacy|exact|convex]
Writes mujoco_playground_b200/models/{ackermann_v2,ackermann_scene}.npz.  Only derived constants are
stored (no XML, no mesh triangles).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mujoco_playground_b200.compiler.mjcf import compile_mjcf  # noqa: E402


def save(M, path):
    out = {}
    for k, v in M.items():
        if isinstance(v, list):
            out[k] = np.array(v, dtype=object if v and not isinstance(v[0], str) else "U64")
        else:
            out[k] = np.asarray(v)
    np.savez_compressed(path, **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference", nargs="?", default="/root/reference")
    ap.add_argument("--mesh-inertia", default="legacy")
    a = ap.parse_args()
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mujoco_playground_b200", "models")
    os.makedirs(dst, exist_ok=True)
    v2 = compile_mjcf(os.path.join(a.reference, "models", "ackermann_robot_v2.xml"), a.mesh_inertia)
    save(v2, os.path.join(dst, "ackermann_v2.npz"))
    sc = compile_mjcf(os.path.join(a.reference, "models", "environments", "ackermann_in_mushr_maze.xml"), a.mesh_inertia)
    save(sc, os.path.join(dst, "ackermann_scene.npz"))
    from mujoco_playground_b200.compiler.maze import MAZES, compile_maze
    for name in MAZES:   # v2 robot in the PointMaze layouts (compiler/maze.py)
        save(compile_maze(os.path.join(a.reference, "models", "ackermann_robot_v2.xml"), name, a.mesh_inertia),
             os.path.join(dst, f"ackermann_maze_{name}.npz"))
    print("wrote", dst)


if __name__ == "__main__":
    main()
