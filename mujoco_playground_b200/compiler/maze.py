"""PointMaze scenes for the Ackermann robot: the model the reference builds in
``AckermannGymnasiumMazeEnv._merge_maze_and_robot_xml`` (src/rl/envs/ackermann_gymnasium_maze_env.py:237-394) by merging
``models/ackermann_robot_v2.xml`` into the maze XML that the third-party package ``gymnasium_robotics`` generates.

That package is not vendored by the reference and is not installable here, so its published maze definition is RESTATED
(gymnasium_robotics.envs.maze.maps / maze.Maze, v1.2-1.3; requirements.txt pins nothing):
  * maps: U_MAZE, OPEN, MEDIUM_MAZE, LARGE_MAZE (1 = block, 0 = free), ``maze_size_scaling = 1``, ``maze_height = 0.4`` for PointMaze;
  * cell (i, j) has its centre at  x = (j + 0.5) s - W s / 2,  y = H s / 2 - (i + 0.5) s  (row 0 is the +y edge);
  * a block is a box of half sizes (0.5 s, 0.5 s, maze_height / 2 s);
  * reset: goal cell and start cell drawn uniformly from the free cells (start != goal: the package resamples while the two
    centres are closer than 0.5 s), each with uniform xy noise of +-0.25 s.
What the reference does on top (kept): ground plane forced to z = -0.5, blocks re-based to sit on it (centre z = -0.5 + half
height), robot spawned at z = -0.445 with identity orientation, 3 settle ``mj_step``s before the odometry reference is taken.
Known deviations (PARITY UNPINNED except for the lidar known answer of SURVEY Appendix D1): the maze XML's own ``<default>`` /
``<option>`` values and its point-mass ball are unknown here; blocks get MuJoCo's default contact parameters and the ball is absent.
"""
from __future__ import annotations

import copy
import xml.etree.ElementTree as ET
from typing import Dict, List, Tuple

import numpy as np

U_MAZE = [[1, 1, 1, 1, 1],
          [1, 0, 0, 0, 1],
          [1, 1, 1, 0, 1],
          [1, 0, 0, 0, 1],
          [1, 1, 1, 1, 1]]
OPEN = [[1, 1, 1, 1, 1, 1, 1],
        [1, 0, 0, 0, 0, 0, 1],
        [1, 0, 0, 0, 0, 0, 1],
        [1, 0, 0, 0, 0, 0, 1],
        [1, 1, 1, 1, 1, 1, 1]]
MEDIUM_MAZE = [[1, 1, 1, 1, 1, 1, 1, 1],
               [1, 0, 0, 1, 1, 0, 0, 1],
               [1, 0, 0, 1, 0, 0, 0, 1],
               [1, 1, 0, 0, 0, 1, 1, 1],
               [1, 0, 0, 1, 0, 0, 0, 1],
               [1, 0, 1, 0, 0, 1, 0, 1],
               [1, 0, 0, 0, 1, 0, 0, 1],
               [1, 1, 1, 1, 1, 1, 1, 1]]
LARGE_MAZE = [[1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1],
              [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1],
              [1, 0, 1, 1, 0, 1, 0, 1, 0, 1, 0, 1],
              [1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1],
              [1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 0, 1],
              [1, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0, 1],
              [1, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 1],
              [1, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1],
              [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]]

# The obstacle layout of models/environments/ackermann_maze_flat.xml (38 blocks of 1 m x 1 m x 0.2 m on an 8 x 8 lattice, floor at
# z = 0), for the "v2 robot in the obstacle scene" merge that the reference's SYSTEM_SUMMARY.md:33-39 describes (its MapSpawner is
# absent upstream).  Row 0 is the +y edge (y = 3), column 0 is x = -4.
MUSHR = [[1, 1, 1, 1, 1, 1, 1, 1],
         [1, 0, 0, 1, 0, 0, 0, 1],
         [1, 0, 1, 0, 0, 1, 0, 1],
         [1, 0, 0, 0, 1, 0, 0, 1],
         [1, 1, 1, 0, 0, 0, 1, 1],
         [1, 0, 0, 0, 1, 0, 0, 1],
         [1, 0, 0, 1, 1, 0, 0, 1],
         [1, 1, 1, 1, 1, 1, 1, 1]]

MAZES: Dict[str, list] = {"umaze": U_MAZE, "open": OPEN, "medium": MEDIUM_MAZE, "large": LARGE_MAZE, "mushr": MUSHR}
# per-layout overrides of (ground z, block half height, spawn z, settle steps, map centre offset)
PARAMS = {"mushr": dict(ground_z=0.0, half_height=0.1, block_z=0.05, spawn_z=0.1, settle=0, offset=(-0.5, -0.5))}
MAZE_ENV_IDS = {"PointMaze_UMaze-v3": "umaze", "PointMaze-Open-v3": "open", "PointMaze-Medium-v3": "medium", "PointMaze-Large-v3": "large"}

SCALING = 1.0
MAZE_HEIGHT = 0.4
GROUND_Z = -0.5              # ackermann_gymnasium_maze_env.py:322-338
SPAWN_Z = -0.445             # :424-441
SETTLE_STEPS = 3             # :227-232
XY_NOISE = 0.25              # gymnasium_robotics maze.add_xy_position_noise (fraction of the cell size)


def cell_xy(name: str, i: int, j: int) -> Tuple[float, float]:
    m = MAZES[name]
    H, W = len(m), len(m[0])
    ox, oy = PARAMS.get(name, {}).get("offset", (0.0, 0.0))
    return (j + 0.5) * SCALING - W * SCALING / 2.0 + ox, H * SCALING / 2.0 - (i + 0.5) * SCALING + oy


def maze_layout(name: str) -> dict:
    """Blocks and free cells in (x, y)-lexicographic order plus the occupancy grid (row iy = ascending y, bit ix = ascending x)."""
    m = MAZES[name]
    H, W = len(m), len(m[0])
    blocks, free = [], []
    for j in range(W):
        for i in range(H - 1, -1, -1):          # ascending y
            (blocks if m[i][j] == 1 else free).append(cell_xy(name, i, j))
    ox, oy = PARAMS.get(name, {}).get("offset", (0.0, 0.0))
    x0, y0 = -W * SCALING / 2.0 + ox, -H * SCALING / 2.0 + oy
    block_rows, free_rows = np.zeros(16), np.zeros(16)
    for i in range(H):
        iy = H - 1 - i
        for j in range(W):
            if m[i][j] == 1:
                block_rows[iy] += float(1 << j)
            else:
                free_rows[iy] += float(1 << j)
    return dict(blocks=blocks, free=free, nx=W, ny=H, x0=x0, y0=y0, pitch=SCALING, block_rows=block_rows, free_rows=free_rows)


def build_maze_root(robot_xml_path: str, name: str) -> ET.Element:
    """MJCF tree of the robot model with the maze blocks added and the floor lowered, following the reference's merge rules."""
    root = copy.deepcopy(ET.parse(robot_xml_path).getroot())
    wb = root.find("worldbody")
    prm = PARAMS.get(name, {})
    gz = prm.get("ground_z", GROUND_Z)
    for g in wb.findall("geom"):
        nm = g.get("name", "").lower()
        if "ground" in nm or "floor" in nm:
            p = g.get("pos", "0 0 0").split()
            g.set("pos", f"{p[0]} {p[1]} {gz}")
    hz = prm.get("half_height", MAZE_HEIGHT / 2.0 * SCALING)
    bz = prm.get("block_z", gz + hz)
    lay = maze_layout(name)
    for k, (x, y) in enumerate(lay["blocks"]):
        ET.SubElement(wb, "geom", dict(name=f"block_{k}", type="box", size=f"{0.5 * SCALING} {0.5 * SCALING} {hz}", pos=f"{x} {y} {bz}"))
    return root


def compile_maze(robot_xml_path: str, name: str, mesh_inertia: str = "legacy") -> dict:
    from .mjcf import compile_mjcf
    M = compile_mjcf(robot_xml_path, mesh_inertia, root=build_maze_root(robot_xml_path, name))
    lay = maze_layout(name)
    M["maze_free_rows"] = lay["free_rows"]
    M["maze_grid"] = np.array([lay["x0"], lay["y0"], lay["pitch"], lay["nx"], lay["ny"]], float)
    M["maze_spawn_z"] = np.array([PARAMS.get(name, {}).get("spawn_z", SPAWN_Z)])
    M["maze_settle_steps"] = np.array([PARAMS.get(name, {}).get("settle", SETTLE_STEPS)])
    M["maze_xy_noise"] = np.array([XY_NOISE * SCALING])
    return M
