"""PointMaze scenes: the v2 robot merged into the U / Open / Medium / Large mazes (compiler/maze.py).

Pins: (i) the reference's own recorded observation (rl_logs/ppo/ppo_model_10000_steps.zip -> _last_obs) is
reproduced by the compiled U-maze: that fixes cell size, map orientation, block placement and the lidar convention;
(ii) the kernel arithmetic (host build) against the oracle on maze states, including reset + settle steps.  The maze XML of the
un-vendored gymnasium_robotics package is restated, see compiler/maze.py for what stays unpinned.
"""
import numpy as np
import pytest

from mujoco_playground_b200.compiler.constants import build_consts
from mujoco_playground_b200.compiler.maze import MAZES, cell_xy, maze_layout
from mujoco_playground_b200.models import load_model
from oracle.env_oracle import OracleEnv
from tests.hostsim.hostsim import HostSim, consts_field
from tests.test_oracle_physics import D1


def _free_cells(name):
    return maze_layout(name)["free"]


def test_maze_layouts():
    for name, n_free, n_block in (("umaze", 7, 18), ("open", 15, 20), ("medium", 26, 38), ("large", 46, 62)):
        lay = maze_layout(name)
        assert (len(lay["free"]), len(lay["blocks"])) == (n_free, n_block)
        M = load_model("maze:" + name)
        boxes = [g for g in range(M["ngeom"]) if M["geom_type"][g] == 6]
        assert len(boxes) == n_block and (M["nq"], M["nv"], M["nu"]) == (13, 12, 3)
        assert np.allclose([M["geom_pos"][g][2] for g in boxes], -0.3) and np.allclose(M["geom_size"][boxes[0]], [0.5, 0.5, 0.2])
        floor = [g for g in range(M["ngeom"]) if M["geom_type"][g] == 0]
        assert len(floor) == 1 and M["geom_pos"][floor[0]][2] == -0.5
        blob = build_consts(M, model_kind=2)
        assert consts_field(blob, "grid_on") == 1 and consts_field(blob, "maze_on") == 1 and consts_field(blob, "settle_steps") == 3
        # free and block masks partition the map
        fr, br = consts_field(blob, "maze_free_rows"), consts_field(blob, "grid_rows")
        nx, ny = int(consts_field(blob, "grid_nx")), int(consts_field(blob, "grid_ny"))
        for iy in range(ny):
            assert int(fr[iy]) & int(br[iy]) == 0 and int(fr[iy]) | int(br[iy]) == (1 << nx) - 1
    # the U opens to the left: row 1 (top corridor) and row 3 (bottom corridor) are joined by the cell at the right (j = 3)
    assert cell_xy("umaze", 2, 3) == (1.0, 0.0) and (1.0, 0.0) in _free_cells("umaze") and (0.0, 0.0) not in _free_cells("umaze")


def test_reference_checkpoint_observation_is_reproduced_by_the_compiled_umaze():
    """Appendix D1: the robot stood in the right-hand cell of the U (cell centre (1, 0)) with start noise (-0.1539, -0.1781)."""
    M = load_model("maze:umaze")
    blob = build_consts(M, model_kind=2)
    yaw = 9e-5
    qpos = M["qpos0"].copy()
    qpos[0:3] = [1.0 - 0.1539, -0.1781, -0.435]      # wheels resting on the floor at z = -0.5
    qpos[3:7] = [np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
    o = OracleEnv(M, kind="v2")                        # plain reset (no settle steps): a level robot at the recorded pose
    want = o.reset(np.zeros(2), spawn_qpos=qpos)
    np.testing.assert_allclose(want[10:72], D1, atol=2e-4)
    assert np.allclose(want[:10], 0.62135, atol=2e-4), "slots 0..9 alias beam 71 (lidar name quirk), as in the checkpoint"
    h = HostSim(blob, False)
    h.qpos[:] = qpos
    h.epd[:] = [0, 0, qpos[0], qpos[1]]
    got, _, _ = h.observe()
    np.testing.assert_allclose(got[:72], want[:72], atol=2e-6)
    np.testing.assert_allclose(got[10:72], D1, atol=2e-4)


@pytest.mark.parametrize("name", ["umaze", "large"])
def test_maze_reset_and_rollout_match_oracle(name):
    """Reset (start / goal cells, noise, 3 settle steps, odometry reference from the last settle step's forward pass) and a
    60-step random-action rollout: host build of the kernel code vs the oracle."""
    M = load_model("maze:" + name)
    blob = build_consts(M, model_kind=2, tolerance=1e-13)
    free = _free_cells(name)
    rng = np.random.default_rng(5)
    for env_id in range(3):
        h = HostSim(blob, False)
        obs0 = h.reset(seed=11, env_id=env_id)
        goal, ref = h.epd[:2].copy(), h.epd[2:4].copy()
        # start and goal lie in distinct free cells, within the +-0.25 noise square
        cg = min(free, key=lambda c: (c[0] - goal[0]) ** 2 + (c[1] - goal[1]) ** 2)
        cs = min(free, key=lambda c: (c[0] - ref[0]) ** 2 + (c[1] - ref[1]) ** 2)
        assert cg != cs and max(abs(goal[0] - cg[0]), abs(goal[1] - cg[1])) <= 0.25 + 1e-9
        assert max(abs(ref[0] - cs[0]), abs(ref[1] - cs[1])) <= 0.25 + 1e-3
        # the oracle from the same spawn pose: qpos before settling = (ref_xy up to the settle drift, spawn z, identity)
        o = OracleEnv(M, kind="maze", tolerance=1e-13)
        qpos = M["qpos0"].copy()
        qpos[0:3] = [h.qpos[0], h.qpos[1], -0.445]
        qpos[3:7] = [1, 0, 0, 0]
        # recover the exact spawn xy: three settle steps move the robot by far less than 1e-3, so re-run the host reset pose
        h2 = HostSim(blob, False)
        h2.qpos[:] = M["qpos0"]; h2.qvel[:] = 0; h2.warm[:] = 0
        spawn_xy = _spawn_xy(blob, 11, env_id)
        qpos[0:2] = spawn_xy
        want0 = o.reset(goal, spawn_qpos=qpos)
        np.testing.assert_allclose(obs0, want0, atol=5e-6)
        assert np.abs(h.qpos - o.sim.qpos).max() < 1e-9 and np.abs(h.qvel - o.sim.qvel).max() < 1e-8
        assert np.abs(ref - o.reference_position[:2]).max() < 1e-9
        for t in range(60):
            a = rng.uniform(-1, 1, 2).astype(np.float32)
            ob, r, te, tr, info = h.step(a)
            oo, r2, te2, tr2, info2 = o.step(a)
            np.testing.assert_allclose(ob, oo, atol=2e-5)
            assert abs(r - r2) < 1e-4 and te == te2 and info["ncon"] == info2["ncon"]
        assert np.abs(h.qpos - o.sim.qpos).max() < 1e-7


def _spawn_xy(blob, seed, env_id):
    """Spawn xy of (seed, env, episode 0) = host reset with the settle steps switched off."""
    from mujoco_playground_b200.compiler.constants import consts_layout
    b = blob.copy()
    b[consts_layout()["settle_steps"][0]] = 0
    h = HostSim(b, False)
    h.reset(seed=seed, env_id=env_id)
    return h.qpos[:2].copy()


def test_mushr_layout_is_the_obstacle_scene_with_the_v2_robot():
    """'maze:mushr' (SURVEY 8f row 4): the v2 robot (79-float observation, BicycleController) among the 38 blocks of
    models/environments/ackermann_maze_flat.xml; block positions must equal the scene model's, floor at z = 0, no settle steps."""
    S, Mz = load_model("scene"), load_model("maze:mushr")
    bs = sorted((round(float(S["geom_pos"][g][0] + S["body_pos"][S["geom_bodyid"][g]][0]), 6), round(float(S["geom_pos"][g][1] + S["body_pos"][S["geom_bodyid"][g]][1]), 6),
                 round(float(S["geom_pos"][g][2]), 6)) for g in range(S["ngeom"]) if S["geom_type"][g] == 6)
    bm = sorted((round(float(Mz["geom_pos"][g][0]), 6), round(float(Mz["geom_pos"][g][1]), 6), round(float(Mz["geom_pos"][g][2]), 6))
                for g in range(Mz["ngeom"]) if Mz["geom_type"][g] == 6)
    assert bs == bm and len(bm) == 38
    blob = build_consts(Mz, model_kind=2)
    assert consts_field(blob, "settle_steps") == 0 and consts_field(blob, "plane_z") == 0 and consts_field(blob, "nbeam") == 72
    h = HostSim(blob, False)
    obs = h.reset(seed=3, env_id=5)
    assert obs.shape == (79,) and np.isfinite(obs).all() and obs[:72].max() < 9.0 and (obs[:72] > 0).all(), "every beam ends on a wall"
    o = OracleEnv(Mz, kind="v2")
    want = o.reset(h.epd[:2], spawn_qpos=np.concatenate([h.qpos[:3], [1, 0, 0, 0], np.zeros(6)]))
    np.testing.assert_allclose(obs, want, atol=5e-6)
