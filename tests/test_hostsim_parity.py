"""Arithmetic of the CUDA per-environment code, compiled for the host (tests/hostsim, LANES = 1), against the CPU oracle.
This is the CPU-side check of the kernel's numerics; the GPU run of the same comparison is tests/test_gpu_parity.py."""
import numpy as np
import pytest

from mujoco_playground_b200.compiler.constants import build_consts
from mujoco_playground_b200.models import load_model
from oracle.env_oracle import OracleEnv
from oracle.oracle import OracleSim
from tests.hostsim.hostsim import HostSim, consts_field

M = load_model("v2")


def _rand_state(rng, ground):
    qpos = M["qpos0"].copy()
    if ground:
        qpos[2] = 0.0645 + rng.uniform(-0.0005, 0.002)
        q = np.array([1.0, *(rng.normal(size=3) * 0.02)])
        qvel = rng.normal(size=12) * np.array([1, 1, .1, .3, .3, 1, 20, 20, 3, 20, 3, 20]) * 0.5
    else:
        qpos[2] = rng.uniform(0.25, 0.5)
        q = rng.normal(size=4)
        qvel = rng.normal(size=12) * np.array([1, 1, 1, 3, 3, 3, 20, 20, 3, 20, 3, 20])
    qpos[:2] = rng.uniform(-5, 5, 2)
    qpos[3:7] = q / np.linalg.norm(q)
    qpos[7:] = rng.uniform(-0.5, 0.5, 6)
    return qpos, qvel


@pytest.mark.parametrize("f32,tol", [(False, 1e-9), (True, 1e-3)])
def test_single_substep_matches_oracle(f32, tol):
    rng = np.random.default_rng(1)
    h = HostSim(build_consts(M, model_kind=0, tolerance=1e-13), f32)
    o = OracleSim(M, tolerance=1e-13)
    for i in range(40):
        qpos, qvel = _rand_state(rng, ground=i % 2 == 0)
        if i % 7 == 0:
            qpos[9] = 0.63                    # beyond the steer limit
        warm = rng.normal(size=12)
        ctrl = rng.uniform(-1, 1, 3) * np.array([0.61, 50, 50])
        o.reset(); o.qpos[:] = qpos; o.qvel[:] = qvel; o.qacc_warmstart[:] = warm; o.ctrl[:] = ctrl
        o.step()
        h.qpos[:], h.qvel[:], h.warm[:] = qpos, qvel, warm
        h.substep(ctrl)
        assert h.diag[0] == o.ncon and h.diag[1] == 0
        scale_v = max(1.0, np.abs(o.qvel).max())
        assert np.abs(h.qpos - o.qpos).max() < tol * max(1.0, np.abs(o.qpos).max())
        assert np.abs(h.qvel - o.qvel).max() < tol * scale_v
        assert np.abs(h.warm - o.qacc_warmstart).max() < (1e-6 if not f32 else 0.5) * max(1.0, np.abs(o.qacc_warmstart).max())


def test_trajectory_1000_steps_fp64():
    """config[0]-style reference trajectory: reset state, 1000 random-action steps; the two independent fp64
    implementations stay together to ~1e-8 (the system is chaotic, errors grow along the rollout)."""
    h = HostSim(build_consts(M, model_kind=0, tolerance=1e-12), False)
    h.reset(seed=0)
    o = OracleEnv(M, tolerance=1e-12)
    o.reset(h.epd[:2])
    rng = np.random.default_rng(0)
    for t in range(1000):
        a = rng.uniform(-1, 1, 2).astype(np.float32)
        obs, r, te, tr, info = h.step(a)
        oo, ro, teo, tro, io = o.step(a)
        assert info["ncon"] == io["ncon"], f"step {t}"
        assert te == teo and tr == tro
        assert abs(r - ro) < 1e-4 * max(1.0, abs(ro))
        if t < 100:
            assert np.abs(h.qpos - o.sim.qpos).max() < 1e-11
            assert np.abs(obs - oo).max() < 1e-5
    assert tr and np.abs(h.qpos - o.sim.qpos).max() < 1e-6


def test_frame_skip_is_repeated_substeps():
    c = build_consts(M, model_kind=0)
    h1, h4 = HostSim(c), HostSim(c)
    h1.reset(seed=3); h4.reset(seed=3)
    rng = np.random.default_rng(3)
    for _ in range(30):
        a = rng.uniform(-1, 1, 2).astype(np.float32)
        for _k in range(4):
            h1.step(a, frame_skip=1)
        h4.step(a, frame_skip=4)
    assert np.abs(h1.qpos - h4.qpos).max() == 0 and h1.epi[0] == 4 * h4.epi[0]


def test_plate_pushed_into_the_floor():
    """Chassis pushed 45 mm into the floor (upright, so the launcher's tilt guard does not route it): the fast NC = 2 kernel has no slot
    for plane-vs-hull contacts and must FLAG the step; the general kernel generates the contacts the oracle generates."""
    from tests.hostsim.hostsim import lib
    h = HostSim(build_consts(M, model_kind=0))
    h.qpos[:] = M["qpos0"]
    h.qpos[2] = 0.02
    h.substep(np.zeros(3))
    assert h.diag[1] == 1
    o = OracleSim(M)
    o.qpos[2] = 0.02
    o.forward()
    assert int(o.f("unsupported_contact")[0]) == 0 and any(c["geom2"] in (1, 2) for c in o.contacts())
    lib().hs_set_general(1)
    try:
        g = HostSim(build_consts(M, model_kind=0))
        g.qpos[:] = M["qpos0"]
        g.qpos[2] = 0.02
        g.substep(np.zeros(3))
        # deep interpenetration: every wheel also has its two cap-independent triangle points (dist <= -r / 2), so the two extra slots
        # of a wheel cannot hold the plate contacts as well -> flagged, but counted
        assert g.diag[0] == o.ncon
    finally:
        lib().hs_set_general(0)


def test_scene_single_substep_matches_oracle():
    """Obstacle scene (ackermann_maze_flat.xml): wheel-vs-box contacts, 4 actuators, no steering equality."""
    S = load_model("scene")
    h = HostSim(build_consts(S, model_kind=1, tolerance=1e-13), False)
    o = OracleSim(S, tolerance=1e-13)
    rng = np.random.default_rng(0)
    nbox = 0
    for t in range(150):
        qpos = S["qpos0"].copy()
        qpos[0] = rng.uniform(-3.45, -2.55); qpos[1] = rng.uniform(-3.45, -2.55); qpos[2] = 0.0648 + rng.uniform(-0.0003, 0.001)
        yaw = rng.uniform(-np.pi, np.pi)
        ang = rng.normal(size=2) * 0.01
        q = np.array([np.cos(yaw / 2), ang[0], ang[1], np.sin(yaw / 2)])
        qpos[3:7] = q / np.linalg.norm(q)
        qpos[7:] = rng.uniform(-0.4, 0.4, 6)
        qvel = rng.normal(size=12) * np.array([.5, .5, .05, .2, .2, .5, 10, 10, 2, 10, 2, 10])
        ctrl = rng.uniform(-1, 1, 4) * np.array([0.6, 0.6, 50, 50])
        o.reset(); o.qpos[:] = qpos; o.qvel[:] = qvel; o.ctrl[:] = ctrl
        o.step()
        h.qpos[:], h.qvel[:], h.warm[:] = qpos, qvel, 0
        h.substep(ctrl)
        nbox += int(any(c["geom1"] != 0 for c in o.contacts()))
        assert h.diag[0] == o.ncon and h.diag[1] == int(o.f("unsupported_contact")[0]) == 0
        assert np.abs(h.qpos - o.qpos).max() < 1e-10
        assert np.abs(h.qvel - o.qvel).max() < 1e-9 * max(1.0, np.abs(o.qvel).max())
    assert nbox > 20


def test_scene_observation_matches_oracle_rays():
    """36 beams against floor + 38 boxes: kernel ray code (host build) vs the oracle's generic ray caster."""
    from oracle.env_oracle import OracleEnv
    S = load_model("scene")
    h = HostSim(build_consts(S, model_kind=1), False)
    o = OracleEnv(S, kind="scene")
    rng = np.random.default_rng(4)
    for _ in range(20):
        qpos = S["qpos0"].copy()
        qpos[0], qpos[1] = rng.uniform(-3.3, -2.7), rng.uniform(-3.3, -2.7)
        yaw = rng.uniform(-np.pi, np.pi)
        qpos[3:7] = [np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
        goal = rng.uniform(-3, 3, 2)
        want = o.reset(goal, spawn_qpos=qpos)
        h.qpos[:] = qpos
        h.epd[:2] = goal
        h.epd[2:4] = qpos[:2]
        got, _, _ = h.observe()
        assert got.shape == (43,)
        np.testing.assert_allclose(got, want, atol=2e-6)


def test_scene_grid_ray_walk_matches_oracle_everywhere():
    """Occupancy-grid ray walk (DDA over the maze cells) vs the oracle's brute-force caster: poses all over the maze, tilted chassis,
    rays that hit the floor first, rays leaving the maze; and grid vs no-grid constants give the same wheel-box contacts."""
    from oracle.env_oracle import OracleEnv
    S = load_model("scene")
    h = HostSim(build_consts(S, model_kind=1), False)
    h0 = HostSim(build_consts(S, model_kind=1, use_box_grid=False), False)
    o = OracleEnv(S, kind="scene")
    rng = np.random.default_rng(11)
    occupied = {(int(round(x)), int(round(y))) for x, y in zip(consts_field(h.blob, "box_cx")[:38], consts_field(h.blob, "box_cy")[:38])}
    assert consts_field(h.blob, "grid_on") == 1 and consts_field(h0.blob, "grid_on") == 0
    n = 0
    while n < 120:
        x, y = rng.uniform(-5.0, 4.0, 2)
        if (int(round(x)), int(round(y))) in occupied:
            continue
        qpos = S["qpos0"].copy()
        qpos[0], qpos[1], qpos[2] = x, y, rng.uniform(0.06, 0.12)
        yaw = rng.uniform(-np.pi, np.pi)
        tilt = rng.normal(size=2) * 0.04
        q = np.array([np.cos(yaw / 2), tilt[0], tilt[1], np.sin(yaw / 2)])
        qpos[3:7] = q / np.linalg.norm(q)
        goal = rng.uniform(-3, 3, 2)
        want = o.reset(goal, spawn_qpos=qpos)
        for sim in (h, h0):
            sim.qpos[:] = qpos
            sim.epd[:2] = goal
            sim.epd[2:4] = qpos[:2]
            got, _, _ = sim.observe()
            np.testing.assert_allclose(got[:36], want[:36], atol=2e-6, err_msg=f"pose {qpos[:7]}")
        n += 1
    # wheel-box detection through the grid: same contacts, same order as the exhaustive loop
    for t in range(60):
        qpos = S["qpos0"].copy()
        qpos[0] = rng.uniform(-3.45, -2.55); qpos[1] = rng.uniform(-3.45, -2.55); qpos[2] = 0.0648 + rng.uniform(-0.0003, 0.001)
        yaw = rng.uniform(-np.pi, np.pi)
        qpos[3:7] = [np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
        qvel = rng.normal(size=12) * np.array([.5, .5, .05, .2, .2, .5, 10, 10, 2, 10, 2, 10])
        ctrl = rng.uniform(-1, 1, 4) * np.array([0.6, 0.6, 50, 50])
        for sim in (h, h0):
            sim.qpos[:], sim.qvel[:], sim.warm[:] = qpos, qvel, 0
            sim.substep(ctrl)
        assert h.diag[0] == h0.diag[0]
        assert np.array_equal(h.qpos, h0.qpos) and np.array_equal(h.qvel, h0.qvel)


@pytest.mark.parametrize("lanes", [4, 8])
def test_multi_lane_code_paths_agree_with_single_lane(lanes):
    """The 4-lane (one lane per wheel) and 8-lane (one lane per floor contact) decompositions, run on the host with one
    thread per lane and a barrier at every team collective, against the 1-lane path: same arithmetic up to summation order."""
    blob = build_consts(M, model_kind=0, tolerance=1e-13)
    h1, hL = HostSim(blob), HostSim(blob)
    h1.reset(seed=2); hL.reset(seed=2)
    rng = np.random.default_rng(1)
    for t in range(60):
        a = rng.uniform(-1, 1, 2).astype(np.float32)
        o1, r1, te1, tr1, i1 = h1.step(a, frame_skip=2)
        oL, rL, teL, trL, iL = hL.step(a, frame_skip=2, lanes=lanes)
        assert i1["ncon"] == iL["ncon"] and te1 == teL and tr1 == trL and abs(r1 - rL) < 1e-6
        assert np.abs(o1 - oL).max() < 1e-5
    assert np.abs(h1.qpos - hL.qpos).max() < 1e-9 and np.abs(h1.qvel - hL.qvel).max() < 1e-8


def test_bad_state_guard_matches_mj_reset_data():
    """mj_checkPos / mj_checkVel: a NaN or |x| > 1e10 in qpos / qvel makes mj_step reset the data to qpos0
    before its forward pass; the env above never notices.  Kernel code (host build) and oracle do the same and count it."""
    c = build_consts(M, model_kind=0, tolerance=1e-12)
    for poison in ("nan_qvel", "huge_qpos"):
        h = HostSim(c, False)
        h.reset(seed=5)
        o = OracleEnv(M, tolerance=1e-12)
        o.reset(h.epd[:2])
        rng = np.random.default_rng(5)
        for t in range(30):
            a = rng.uniform(-1, 1, 2).astype(np.float32)
            if t == 10:
                if poison == "nan_qvel":
                    h.qvel[7] = np.nan; o.sim.qvel[7] = np.nan
                else:
                    h.qpos[0] = 3e10; o.sim.qpos[0] = 3e10
            obs, r, te, tr, info = h.step(a)
            oo, ro, teo, tro, io = o.step(a)
            assert info["bad"] == (1 if t == 10 else 0)
            assert np.isfinite(obs).all() and np.abs(obs - oo).max() < 1e-5
            assert np.abs(h.qpos - o.sim.qpos).max() < 1e-9 and np.abs(h.qvel - o.sim.qvel).max() < 1e-7
        assert int(o.sim.f("nbad")[0]) == 1


def _quat(axis, deg):
    a = np.deg2rad(deg) / 2
    return np.array([np.cos(a), *(np.sin(a) * np.asarray(axis, float))])


@pytest.mark.parametrize("f32,tol", [(False, 1e-7), (True, 1e-3)])
def test_rollover_contacts_match_oracle(f32, tol):
    """Wheel caps on the floor (the two extra mjc_PlaneCylinder points), chassis-plate hulls on the floor (mjc_PlaneConvex) and the
    disk-parallel-to-the-floor corner: the general (NC = 4) kernel path generates them all -- contact counts exact, nothing flagged
    unsupported, single-step velocities within tolerance of the oracle, over tumbling trajectories from 11 poses."""
    from tests.hostsim.hostsim import lib
    c = build_consts(M, model_kind=0, tolerance=1e-12)
    lib().hs_set_general(1)
    try:
        rng = np.random.default_rng(0)
        cases = [(_quat([1, 0, 0], 180), 0.12), (_quat([0, 1, 0], 80), 0.2), (_quat([1, 0, 0], 90), 0.13), (_quat([0, 1, 0], 95), 0.25),
                 (_quat([0, 1, 0], -100), 0.25)]
        for _ in range(6):
            q = rng.normal(size=4)
            cases.append((q / np.linalg.norm(q), 0.25))
        seen = set()
        for q, z in cases:
            o = OracleSim(M, tolerance=1e-12)
            h = HostSim(c, f32)
            q0 = M["qpos0"].copy(); q0[:3] = [0, 0, z]; q0[3:7] = q
            o.reset(); o.qpos[:] = q0; o.qvel[3:6] = [1.0, -2.0, 0.5]
            for t in range(300):
                qo, vo, wo = o.qpos.copy(), o.qvel.copy(), o.qacc_warmstart.copy()
                o.ctrl[:] = 0
                o.step()
                h.qpos[:], h.qvel[:], h.warm[:] = qo, vo, wo
                h.substep(np.zeros(3))
                assert h.diag[0] == o.ncon and h.diag[1] == 0, f"step {t}: ncon {h.diag[0]} vs {o.ncon}, unsupported {h.diag[1]}"
                assert np.abs(h.qvel - o.qvel).max() < tol * max(1.0, np.abs(o.qvel).max())
                seen |= {(cc["geom1"], cc["geom2"]) for cc in o.contacts()}
                seen.add(("n", o.ncon))
        assert (0, 1) in seen and (0, 2) in seen, "both plates touched the floor"
        assert any(k[0] == "n" and k[1] >= 11 for k in seen), "wheel + plate contacts together"
    finally:
        lib().hs_set_general(0)


@pytest.mark.parametrize("name", ["maze:umaze", "maze:mushr"])
def test_plate_vs_wall_contacts_match_oracle(name):
    """Maze models: the chassis plates (contype 2 / conaffinity 1) collide with the wall blocks (1 / 1) and reach 2 cm further forward
    than the front wheels, so a head-on wall hit is a PLATE contact.  Kernel code and oracle generate it (box face normal, blended
    deepest hull vertices); driving full throttle into walls: contact counts exact, single-step parity, nothing unsupported."""
    from mujoco_playground_b200.compiler.constants import consts_layout
    Mz = load_model(name)
    c = build_consts(Mz, model_kind=2, tolerance=1e-12)
    c0 = c.copy()
    c0[consts_layout()["settle_steps"][0]] = 0
    plate_steps = 0
    gt = Mz["geom_type"]
    for seed in range(3):
        h = HostSim(c, False)
        h.reset(seed=seed, env_id=seed)
        h0 = HostSim(c0, False)
        h0.reset(seed=seed, env_id=seed)
        o = OracleEnv(Mz, kind="maze", tolerance=1e-12)
        o.reset(h.epd[:2], spawn_qpos=h0.qpos.copy())
        for t in range(500):
            a = np.array([1.0, 0.3 * np.sin(0.01 * t + seed)], np.float32)
            qo, vo, wo = o.sim.qpos.copy(), o.sim.qvel.copy(), o.sim.qacc_warmstart.copy()
            _, _, _, _, io = o.step(a)
            h.qpos[:], h.qvel[:], h.warm[:] = qo, vo, wo
            _, _, _, _, info = h.step(a)
            assert info["ncon"] == io["ncon"] and info["unsupported"] == 0, f"seed {seed} step {t}"
            assert np.abs(h.qvel - o.sim.qvel).max() < 1e-8 * max(1.0, np.abs(o.sim.qvel).max())
            plate_steps += int(any(gt[cc["geom1"]] == 6 and gt[cc["geom2"]] == 7 for cc in o.sim.contacts()))
    assert plate_steps > 100, "the drive must produce plate-vs-block contacts"


@pytest.mark.parametrize("f32,tol", [(False, 1e-9), (True, 2e-4)])
def test_flat_floor_lidar_matches_oracle_rays(f32, tol):
    """Flat-floor model (72 slots, beam origins on a 0.035 m circle, 12 m cutoff, 40 m floor plane): the kernel's floor test works from
    a per-environment base (lidar centre B, height h0: ray parameter -h0 / lvz - r, hit point B - (h0 / lvz) dw), the oracle casts every
    ray from its own origin like mj_ray.  Upright, tilted, tumbling and high poses, robots at the edge of the plane (hits beyond the
    half size do not count: -1) and rays above the horizon."""
    o = OracleEnv(M)
    h = HostSim(build_consts(M, model_kind=0), f32)
    rng = np.random.default_rng(21)
    seen_miss = seen_cut = seen_near = 0
    for i in range(300):
        qpos = M["qpos0"].copy()
        mode = i % 5
        if mode == 0:      # driving pose
            qpos[:2] = rng.uniform(-20, 20, 2); qpos[2] = 0.0645 + rng.uniform(0, 0.002)
            q = np.array([1.0, *(rng.normal(size=3) * 0.03)])
        elif mode == 1:    # tilted, in the air
            qpos[:2] = rng.uniform(-20, 20, 2); qpos[2] = rng.uniform(0.1, 0.6)
            q = np.array([1.0, *(rng.normal(size=3) * 0.3)])
        elif mode == 2:    # tumbling
            qpos[:2] = rng.uniform(-20, 20, 2); qpos[2] = rng.uniform(0.2, 1.0)
            q = rng.normal(size=4)
        elif mode == 3:    # at the edge of the 40 m plane, nose down: near hits inside, far hits beyond the edge
            qpos[0] = rng.choice([-1, 1]) * rng.uniform(39.0, 39.99); qpos[1] = rng.uniform(-39.9, 39.9); qpos[2] = rng.uniform(0.07, 0.3)
            q = np.array([1.0, *(rng.normal(size=3) * 0.1)])
        else:              # high up: every downward ray is longer than the cutoff
            qpos[:2] = rng.uniform(-5, 5, 2); qpos[2] = rng.uniform(3.0, 8.0)
            q = np.array([1.0, *(rng.normal(size=3) * 0.2)])
        qpos[3:7] = q / np.linalg.norm(q)
        goal = rng.uniform(-3, 3, 2)
        want = o.reset(goal, spawn_qpos=qpos)
        h.qpos[:] = qpos
        h.epd[:2] = goal
        h.epd[2:4] = qpos[:2]
        got, _, _ = h.observe()
        lid_w, lid_g = want[:72], got[:72]
        # a ray within rounding of the horizon or of the plane's edge may flip between hit and miss in fp32: compare where both agree
        miss_w, miss_g = lid_w < 0, lid_g < 0
        flips = int((miss_w != miss_g).sum())
        assert flips <= (2 if f32 else 0), f"pose {i}: {flips} hit / miss flips"
        both = ~miss_w & ~miss_g
        # (relative: a grazing ray's length is h0 / lvz with lvz close to zero)
        assert np.all(np.abs(lid_g[both] - lid_w[both]) <= tol * np.maximum(1.0, np.abs(lid_w[both]))), f"pose {i}"
        np.testing.assert_allclose(got[72:], want[72:], atol=2e-6 if not f32 else 2e-4)
        seen_miss += int(miss_w.sum()); seen_cut += int((lid_w == 12.0).sum()); seen_near += int((both & (lid_w < 1.0)).sum())
    assert seen_miss > 1000 and seen_cut > 1000 and seen_near > 300, (seen_miss, seen_cut, seen_near)
