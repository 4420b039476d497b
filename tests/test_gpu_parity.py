"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Run with -m gpu on a B200.

Tolerances (BASELINE.json north_star): single-step qpos/qvel within 1e-5 relative in fp64 mode and
1e-3 in fp32 mode; <=100-step trajectories within the tolerance stated in each test; integer contact
counts and done flags exact.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _models():
    from mujoco_playground_b200.models import load_model
    return load_model("v2")


def _random_states(M, n, rng, ground_frac=0.5):
    qpos = np.tile(M["qpos0"], (n, 1))
    qvel = np.zeros((n, 12))
    for i in range(n):
        if rng.uniform() < ground_frac:
            qpos[i, 2] = 0.0645 + rng.uniform(-0.0005, 0.002)
            ang = rng.normal(size=3) * 0.02
            q = np.array([1.0, *ang])
            qvel[i] = rng.normal(size=12) * np.array([1, 1, .1, .3, .3, 1, 20, 20, 3, 20, 3, 20]) * 0.5
        else:
            qpos[i, 2] = rng.uniform(0.25, 0.5)   # clear of the floor at any orientation (cap-down wheel contacts are out of scope)
            q = rng.normal(size=4)
            qvel[i] = rng.normal(size=12) * np.array([1, 1, 1, 3, 3, 3, 20, 20, 3, 20, 3, 20])
        qpos[i, 0:2] = rng.uniform(-5, 5, 2)
        qpos[i, 3:7] = q / np.linalg.norm(q)
        qpos[i, 7:] = rng.uniform(-0.5, 0.5, 6)
        if rng.uniform() < 0.2:
            qpos[i, 9] = 0.62   # beyond the steer limit
    return qpos, qvel


def _rel(a, b):
    return np.abs(a - b).max() / max(1e-9, np.abs(b).max())


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-5), ("float32", 1e-3)])
def test_single_step_matches_oracle(dtype, tol):
    from mujoco_playground_b200 import BatchedAckermannEnv
    from oracle.env_oracle import OracleEnv
    M = _models()
    n = 96
    rng = np.random.default_rng(5)
    qpos, qvel = _random_states(M, n, rng)
    warm = rng.normal(size=(n, 12))
    acts = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    env = BatchedAckermannEnv(n, dtype=dtype, auto_reset=False, solver_tolerance=1e-12 if dtype == "float64" else None)
    env.reset()
    env.set_state(qpos, qvel, warm)
    _, _, _, _, info = env.step(torch.from_numpy(acts).cuda())
    q2, v2, w2 = env.get_state()
    ncon = info["ncon"].cpu().numpy()
    o = OracleEnv(M, tolerance=1e-12)
    worst_q = worst_v = 0.0
    for i in range(n):
        o.reset(np.zeros(2))
        o.sim.qpos[:] = qpos[i]; o.sim.qvel[:] = qvel[i]; o.sim.qacc_warmstart[:] = warm[i]
        o.step(acts[i])
        assert ncon[i] == o.sim.ncon, f"env {i}: contact count {ncon[i]} != oracle {o.sim.ncon}"
        worst_q = max(worst_q, _rel(q2[i], o.sim.qpos))
        worst_v = max(worst_v, _rel(v2[i], o.sim.qvel))
    print(f"{dtype}: worst relative qpos {worst_q:.3e} qvel {worst_v:.3e}")
    assert worst_q < tol and worst_v < tol
    env.close()


@pytest.mark.parametrize("dtype,steps,tol_q,tol_obs", [("float64", 100, 1e-7, 1e-4), ("float32", 60, 5e-4, 5e-3)])
def test_trajectory_from_reset(dtype, steps, tol_q, tol_obs):
    """<=100-step trajectories from the reference reset state with U(-1,1) actions.  fp32 tolerance is absolute
    5e-4 on qpos over 60 steps; beyond ~150 steps fp32 diverges chaotically through contact switching."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    from oracle.env_oracle import OracleEnv
    M = _models()
    n = 6
    env = BatchedAckermannEnv(n, dtype=dtype, seed=3, auto_reset=False, solver_tolerance=1e-12 if dtype == "float64" else None)
    obs0 = env.reset().cpu().numpy().copy()
    goal, ref, _ = env.get_episode()
    rng = np.random.default_rng(11)
    oracles = [OracleEnv(M, tolerance=1e-12) for _ in range(n)]
    for k, o in enumerate(oracles):
        assert np.abs(o.reset(goal[k]) - obs0[k]).max() < 1e-5
    ncon_flips = flag_flips = 0
    worst_obs = 0.0
    for t in range(steps):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        obs, rew, term, trunc, info = env.step(torch.from_numpy(a).cuda())
        obs, rew, term, trunc, ncon = (x.cpu().numpy() for x in (obs, rew, term, trunc, info["ncon"]))
        for k, o in enumerate(oracles):
            oo, r, te, tr, inf = o.step(a[k])
            # north_star: "integer contact counts and done flags must match exactly".  Exact in fp64.  In fp32 a contact whose
            # distance is within rounding of zero can switch one step early / late: the flip RATE is measured and bounded here
            # (and reported), never skipped; the done flags must still match exactly.
            ncon_flips += int(ncon[k] != inf["ncon"])
            flag_flips += int(bool(term[k]) != te) + int(bool(trunc[k]) != tr)
            worst_obs = max(worst_obs, float(np.abs(oo - obs[k]).max()))
            if dtype == "float64":
                assert ncon[k] == inf["ncon"], f"step {t} env {k}: ncon {ncon[k]} vs {inf['ncon']}"
            assert abs(r - rew[k]) < (1e-4 if dtype == "float64" else 2e-3) * max(1.0, abs(r))
    total = steps * n
    print(f"{dtype}: contact-count flips {ncon_flips}/{total} ({100.0 * ncon_flips / total:.2f} %), done-flag flips {flag_flips}, "
          f"worst |obs - oracle| {worst_obs:.2e}")
    assert flag_flips == 0
    assert worst_obs < tol_obs
    assert ncon_flips == 0 if dtype == "float64" else ncon_flips <= 0.03 * total, "fp32 contact-count flip rate above 3 % of env steps"
    q, v, _ = env.get_state()
    for k, o in enumerate(oracles):
        assert np.abs(q[k] - o.sim.qpos).max() < tol_q, f"{dtype} env {k}: qpos error {np.abs(q[k] - o.sim.qpos).max()}"
    env.close()


def test_lane_layouts_agree():
    """1, 4 and 8 lanes per environment run the same arithmetic up to summation order (tight solver tolerance so that the
    different orders cannot stop the Newton iteration at different points)."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    res = []
    rng = np.random.default_rng(2)
    acts = rng.uniform(-1, 1, (40, 33, 2)).astype(np.float32)
    for lanes in (1, 4, 8):
        env = BatchedAckermannEnv(33, dtype="float64", seed=9, lanes_per_env=lanes, auto_reset=False, frame_skip=2, solver_tolerance=1e-12)
        env.reset()
        for t in range(40):
            obs, *_ = env.step(torch.from_numpy(acts[t]).cuda())
        res.append((env.get_state()[0], obs.cpu().numpy().copy()))
        env.close()
    for q, o in res[1:]:
        assert np.abs(q - res[0][0]).max() < 1e-9
        assert np.abs(o - res[0][1]).max() < 1e-5


def test_auto_reset_truncation_and_terminal_obs():
    from mujoco_playground_b200 import BatchedAckermannEnv
    n = 70
    env = BatchedAckermannEnv(n, dtype="float32", seed=4, max_episode_steps=5, auto_reset=True)
    obs0 = env.reset().cpu().numpy().copy()
    last = None
    for t in range(5):
        obs, rew, term, trunc, info = env.step(None)
        if t < 4:
            assert int(trunc.sum().item()) == 0
            last = obs.cpu().numpy().copy()
    assert int(trunc.sum().item()) == n, "every env must truncate at max_episode_steps (ackermann_env.py:219-220)"
    tobs = info["terminal_observation"].cpu().numpy()
    new = obs.cpu().numpy()
    # the new episode starts from the spawn pose: lidar all -1 (level robot), odometry zero
    assert np.allclose(new[:, 72:75], 0.0, atol=1e-6)
    assert np.allclose(new[:, :72], obs0[:, :72])
    # terminal observation continues the old episode (odometry moved a little, same goal as before the reset)
    assert np.allclose(tobs[:, 75:77] + tobs[:, 72:74], last[:, 75:77] + last[:, 72:74], atol=1e-5)
    _, _, sc = env.get_episode()
    assert (sc == 0).all()
    st = env.stats()
    assert st["episodes"] == n and st["env_steps"] == 5 * n
    env.close()


@pytest.mark.parametrize("pinned", [True, False])
def test_step_host_matches_device_step(pinned):
    """ackb_step_host with pinned caller buffers (zero-copy: the kernel reads / writes host memory directly) and with
    pageable ones (staged copies) returns exactly what the device-tensor step returns, auto-reset rows included."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    n = 50
    rng = np.random.default_rng(8)
    a = torch.from_numpy(rng.uniform(-1, 1, (n, 2)).astype(np.float32))
    e1 = BatchedAckermannEnv(n, seed=5, auto_reset=True, max_episode_steps=12)
    e2 = BatchedAckermannEnv(n, seed=5, auto_reset=True, max_episode_steps=12)
    e1.reset(); e2.reset()
    mk = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    h = [mk(torch.empty((n, e2.obs_dim))), mk(torch.empty(n)), mk(torch.empty(n, dtype=torch.uint8)), mk(torch.empty(n, dtype=torch.uint8))]
    ah = mk(a.clone())
    ntrunc = 0
    for _ in range(30):
        o1, r1, t1, u1, _ = e1.step(a.cuda())
        e2.step_host(ah, *h)
        assert torch.equal(o1.cpu(), h[0]) and torch.equal(r1.cpu(), h[1]) and torch.equal(t1.cpu(), h[2]) and torch.equal(u1.cpu(), h[3])
        ntrunc += int(h[3].sum())
    assert ntrunc == 2 * n, "two truncations (with auto-reset observations) per environment in 30 steps"
    e1.close(); e2.close()


def test_gym_adapter_signature():
    from mujoco_playground_b200 import AckermannRobotEnv
    env = AckermannRobotEnv()
    obs, info = env.reset(seed=0)
    assert obs.shape == (79,) and obs.dtype == np.float32 and set(info) == {"map_name", "goal_position", "start_position"}
    obs, r, te, tr, info = env.step(np.array([0.5, 0.1], np.float32))
    assert obs.shape == (79,) and isinstance(r, float) and isinstance(te, bool) and isinstance(tr, bool)
    assert {"goal_distance", "collision", "min_lidar", "step", "linear_velocity", "angular_velocity"} <= set(info)
    # reference quirk Q2: no lidar hit (-1) counts as a collision => -50 on an open floor
    assert info["collision"] and r < -50
    env.close()


def test_no_cpu_fallback_message():
    from mujoco_playground_b200 import _lib
    L = _lib.load()
    import ctypes
    h = ctypes.c_void_p()
    rc = L.ackb_create(None, 0, 1, 0, 0, 0, 4, ctypes.byref(h))
    assert rc < 0 and b"null" in L.ackb_last_error(None)


# ---------------------------------------------------------------------------------------------------------------
# obstacle scene (models/environments/ackermann_maze_flat.xml): 4 actuators, AckermannController, 38 boxes, 36 beams
# ---------------------------------------------------------------------------------------------------------------
def _scene_states(S, n, rng):
    qpos = np.tile(S["qpos0"], (n, 1))
    qvel = np.zeros((n, 12))
    for i in range(n):
        qpos[i, 0] = rng.uniform(-3.45, -2.55)       # spawn cell: walls at x = -3.5, y = -3.5 and the block at (-3, -2)
        qpos[i, 1] = rng.uniform(-3.45, -2.55)
        qpos[i, 2] = 0.0648 + rng.uniform(-0.0003, 0.001)
        yaw = rng.uniform(-np.pi, np.pi)
        ang = rng.normal(size=2) * 0.01
        q = np.array([np.cos(yaw / 2), ang[0], ang[1], np.sin(yaw / 2)])
        qpos[i, 3:7] = q / np.linalg.norm(q)
        qpos[i, 7:] = rng.uniform(-0.4, 0.4, 6)
        qvel[i] = rng.normal(size=12) * np.array([.5, .5, .05, .2, .2, .5, 10, 10, 2, 10, 2, 10])
    return qpos, qvel


@pytest.mark.parametrize("dtype,tol,lanes", [("float64", 1e-5, 4), ("float64", 1e-5, 1), ("float32", 1e-3, 4), ("float32", 1e-3, 1)])
def test_scene_single_step_matches_oracle(dtype, tol, lanes):
    from mujoco_playground_b200 import BatchedAckermannEnv
    from mujoco_playground_b200.models import load_model
    from oracle.env_oracle import OracleEnv
    S = load_model("scene")
    n = 128
    rng = np.random.default_rng(21)
    qpos, qvel = _scene_states(S, n, rng)
    acts = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    env = BatchedAckermannEnv(n, model="scene", dtype=dtype, auto_reset=False, lanes_per_env=lanes,
                              solver_tolerance=1e-12 if dtype == "float64" else None)
    assert env.obs_dim == 43
    env.reset()
    env.set_state(qpos, qvel, np.zeros((n, 12)))
    obs, _, _, _, info = env.step(torch.from_numpy(acts).cuda())
    obs = obs.cpu().numpy()
    q2, v2, _ = env.get_state()
    ncon = info["ncon"].cpu().numpy()
    o = OracleEnv(S, kind="scene", tolerance=1e-12)
    worst, nbox, flips, lid_err = 0.0, 0, 0, 0.0
    for i in range(n):
        o.reset(np.zeros(2))
        o.sim.qpos[:] = qpos[i]; o.sim.qvel[:] = qvel[i]
        oo, *_ = o.step(acts[i])
        nbox += int(any(c["geom1"] != 0 for c in o.sim.contacts()))
        flips += int(ncon[i] != o.sim.ncon)
        lid_err = max(lid_err, float(np.abs(oo[:36] - obs[i, :36]).max()))
        if dtype == "float64":
            assert ncon[i] == o.sim.ncon
        worst = max(worst, _rel(q2[i], o.sim.qpos), _rel(v2[i], o.sim.qvel))
    print(f"scene {dtype} lanes={lanes}: contact-count flips {flips}/{n}, worst lidar error {lid_err:.2e}, worst rel state error {worst:.2e}")
    assert nbox > 10, "the sample must exercise wheel-box contacts"
    assert lid_err < (1e-4 if dtype == "float64" else 2e-3), "36-beam lidar against the maze walls"
    assert flips == 0 if dtype == "float64" else flips <= 3, "fp32: a contact at distance ~0 may switch (bounded, reported)"
    assert worst < tol
    env.close()


def test_scene_rollout_with_spawn_jitter_stays_supported():
    """config[2]-style rollout: per-env yaw / xy jitter at spawn so that wheels hit the maze walls."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    env = BatchedAckermannEnv(2048, model="scene", dtype="float32", seed=3, frame_skip=4, spawn_yaw_range=np.pi, spawn_xy_jitter=0.25)
    env.reset()
    hit = 0
    for _ in range(150):
        obs, rew, term, trunc, info = env.step(None)
        hit = max(hit, int((info["ncon"] > 8).sum().item()))
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert hit > 0, "some environments must be in contact with obstacles"
    st = env.stats()
    assert st["unsupported"] < 0.01 * st["env_steps"]
    env.close()


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json's full sizes, through size-independent properties
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_flat_batch_is_layout_and_size_invariant():
    """configs[3] size (131072 envs, 1 lane per env, tail mode) against configs[1] size (4096 envs, 4 lanes per env): goals,
    synthetic actions and resets are keyed by (seed, env id, step), so environment i must follow the same trajectory in both
    batches whatever the batch size, lane layout or CTA it runs in.  fp32, 12 env-steps of 4 substeps: tolerance 2e-4 on qpos,
    flags and contact counts exact; plus run-to-run bit reproducibility of the large batch."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    steps, small_n, big_n = 12, 4096, 131072
    out = {}
    for tag, n in (("small", small_n), ("big", big_n), ("big2", big_n)):
        env = BatchedAckermannEnv(n, dtype="float32", seed=77, frame_skip=4, auto_reset=True, max_episode_steps=7)
        env.reset()
        flags, ncon = [], []
        for _ in range(steps):
            obs, rew, term, trunc, info = env.step(None)
            flags.append(torch.stack([term[:small_n], trunc[:small_n]]).cpu().numpy().copy())
            ncon.append(info["ncon"][:small_n].cpu().numpy().copy())
        q, v, _ = env.get_state()
        out[tag] = (q, v, obs.cpu().numpy().copy(), np.array(flags), np.array(ncon), rew.cpu().numpy().copy())
        st = env.stats()
        assert st["unsupported"] == 0 and st["episodes"] == n, "one truncation (auto-reset) per environment in 12 steps"
        assert np.isfinite(q).all() and np.isfinite(out[tag][2]).all()
        env.close()
    qs, vs, os_, fs_, ns, _ = out["small"]
    qb, vb, ob, fb, nb, _ = out["big"]
    assert np.array_equal(fs_, fb), "done flags must match exactly"
    assert np.array_equal(ns, nb), "contact counts must match exactly"
    assert np.abs(qs - qb[:small_n]).max() < 2e-4
    assert np.abs(os_ - ob[:small_n]).max() < 2e-3
    for a, b in zip(out["big"], out["big2"]):
        assert np.array_equal(a, b), "two runs of the same batch must be bit identical"


def test_full_size_scene_rollout_properties():
    """configs[2] size: 65536 obstacle-scene environments with spawn jitter.  Properties: finite state, robots stay above the
    floor and inside the maze walls, contact counts in range, a visible fraction of env-steps has wheel-box contacts, the share
    of env-steps with unsupported contact geometry stays below 1 %."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    n = 65536
    env = BatchedAckermannEnv(n, model="scene", dtype="float32", seed=5, frame_skip=4, spawn_yaw_range=np.pi, spawn_xy_jitter=0.12)
    env.reset()
    with_box = 0
    max_ncon = 0
    for _ in range(60):
        obs, rew, term, trunc, info = env.step(None)
        nc = info["ncon"]
        with_box += int((nc > 8).sum().item())
        max_ncon = max(max_ncon, int(nc.max().item()))
    q, v, _ = env.get_state()
    assert np.isfinite(q).all() and np.isfinite(v).all() and torch.isfinite(obs).all() and torch.isfinite(rew).all()
    # (larger jitters start robots overlapping a wall block; those are pushed out violently, as in MuJoCo, and are not a sane workload)
    assert (q[:, 2] > -0.01).all() and (q[:, 2] < 0.3).all(), f"chassis height stays physical: {q[:, 2].min()} .. {q[:, 2].max()}"
    assert ((q[:, 2] > 0.05) & (q[:, 2] < 0.09)).mean() > 0.995, "robots drive on the floor"
    assert (np.abs(q[:, 0]) < 4.5).all() and (np.abs(q[:, 1]) < 4.5).all(), "nobody leaves the maze"
    assert np.abs(np.linalg.norm(q[:, 3:7], axis=1) - 1).max() < 1e-3
    assert 8 <= max_ncon <= 16
    st = env.stats()
    frac = st["obstacle_steps"] / st["env_steps"]
    assert frac > 0.002, f"fraction of env-steps with a wheel-obstacle contact {frac}"
    # random +-50 rad/s wheel commands make the light robot hop: the oracle shows the same 0..8 contact distribution with mean ~2
    assert 1.0 < st["contacts_sum"] / st["env_steps"] <= 9.0, "mean contact count"
    assert st["unsupported"] < 0.001 * st["env_steps"]
    env.close()


def test_long_horizon_statistical_parity_fp32_vs_fp64():
    """Contact switching makes single trajectories chaotic (fp32 and fp64 decorrelate after ~150-300 steps),
    so beyond the short horizon parity is checked in distribution: 8192 environments, full 1000-step episodes, identical goals and
    action streams.  Episode statistics of the fp32 and fp64 batches must agree within sampling error."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    n, steps = 8192, 1000
    res = {}
    for dtype in ("float32", "float64"):
        env = BatchedAckermannEnv(n, dtype=dtype, seed=2024, frame_skip=1, auto_reset=True)
        env.reset()
        ret = torch.zeros(n, device="cuda")
        for _ in range(steps):
            obs, rew, term, trunc, info = env.step(None)
            ret += rew
        st = env.stats()
        q, v, _ = env.get_state()
        res[dtype] = dict(ret=ret.cpu().numpy(), st=st, dist=obs[:, 77].cpu().numpy().copy())
        assert np.isfinite(q).all() and st["unsupported"] == 0
        env.close()
    a, b = res["float32"], res["float64"]
    assert a["st"]["episodes"] == b["st"]["episodes"], "every environment truncates exactly once at step 1000 (goals are >= 2 m away)"
    assert a["st"]["successes"] == b["st"]["successes"] == 0 or abs(a["st"]["successes"] - b["st"]["successes"]) <= 5
    ma, mb = a["ret"].mean(), b["ret"].mean()
    se = np.sqrt(a["ret"].var() / n + b["ret"].var() / n)
    assert abs(ma - mb) < 5 * se + 1e-3 * abs(mb), f"mean 1000-step return fp32 {ma} vs fp64 {mb} (se {se})"
    assert abs(a["ret"].std() - b["ret"].std()) < 0.05 * b["ret"].std()
    assert abs(a["st"]["solver_iters"] - b["st"]["solver_iters"]) < 0.05 * b["st"]["solver_iters"]
    assert abs(a["st"]["contacts_sum"] - b["st"]["contacts_sum"]) < 0.05 * b["st"]["contacts_sum"]
    # per-environment: the two precisions stay correlated through the slow variables (goal distance after 8 s of random driving)
    assert np.corrcoef(a["ret"], b["ret"])[0, 1] > 0.95


# ---------------------------------------------------------------------------------------------------------------
# PointMaze scenes (v2 robot in the U / Open / Medium / Large mazes): reset with settle steps, deferred auto-reset
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,dtype,tol", [("umaze", "float64", 1e-7), ("large", "float32", 2e-3), ("mushr", "float64", 1e-7),
                                            ("mushr", "float32", 2e-3)])
def test_maze_env_matches_oracle(name, dtype, tol):
    """v2 robot in a block maze; "mushr" = the 38 blocks of ackermann_maze_flat.xml."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    from mujoco_playground_b200.models import load_model
    from oracle.env_oracle import OracleEnv
    M = load_model("maze:" + name)
    n, steps = 64, 40
    env = BatchedAckermannEnv(n, model="maze:" + name, dtype=dtype, seed=21, auto_reset=False,
                              solver_tolerance=1e-12 if dtype == "float64" else None)
    assert env.obs_dim == 79
    # spawn poses: the same reset with the settle steps switched off (state before settling)
    env0 = BatchedAckermannEnv(n, model="maze:" + name, dtype="float64", seed=21, auto_reset=False, settle_steps=0)
    env0.reset()
    spawn = env0.get_state()[0]
    env0.close()
    obs0 = env.reset().cpu().numpy().copy()
    goal, ref, _ = env.get_episode()
    rng = np.random.default_rng(3)
    acts = rng.uniform(-1, 1, (steps, n, 2)).astype(np.float32)
    ks = [0, 7, 33]
    oracles = []
    for k in ks:
        o = OracleEnv(M, kind="maze", tolerance=1e-12)
        want = o.reset(goal[k], spawn_qpos=spawn[k])
        assert np.abs(want - obs0[k]).max() < (1e-5 if dtype == "float64" else 2e-4), "reset observation after the settle steps"
        assert np.abs(o.reference_position[:2] - ref[k]).max() < 1e-6
        oracles.append(o)
    for t in range(steps):
        obs, rew, term, trunc, info = env.step(torch.from_numpy(acts[t]).cuda())
        for k, o in zip(ks, oracles):
            oo, r, te, tr, inf = o.step(acts[t, k])
            if dtype == "float64":
                assert int(info["ncon"][k].item()) == inf["ncon"] and bool(term[k].item()) == te
    q, v, _ = env.get_state()
    for k, o in zip(ks, oracles):
        assert _rel(q[k], o.sim.qpos) < tol, f"{name} {dtype} env {k}"
    env.close()


def test_maze_auto_reset_uses_the_deferred_reset_launch():
    """Models with settle steps: the step kernel marks finished environments and a masked reset launch (reset + 3 settle steps +
    observation) follows on the same stream; the caller sees the same contract as with the fused reset."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    n = 300
    env = BatchedAckermannEnv(n, model="maze:medium", dtype="float32", seed=8, max_episode_steps=6, auto_reset=True, frame_skip=2)
    env.reset()
    goal0, ref0, _ = env.get_episode()
    l0 = env.launch_count
    for t in range(6):
        obs, rew, term, trunc, info = env.step(None)
    assert env.launch_count - l0 == 12, "step kernel + masked reset kernel per step"
    tr = trunc.cpu().numpy().astype(bool)
    assert tr.sum() >= n - 10, "a few start/goal pairs in adjacent cells end by reaching the goal and restart earlier"
    new = obs.cpu().numpy()
    tobs = info["terminal_observation"].cpu().numpy()
    assert np.allclose(new[tr, 72:75], 0.0, atol=1e-5), "fresh episode: odometry restarts at zero after the settle steps"
    assert not np.allclose(tobs[tr, 72:74], 0.0, atol=1e-6), "terminal observation belongs to the old episode"
    goal1, ref1, sc = env.get_episode()
    assert (sc[tr] == 0).all() and np.abs(goal1 - goal0).max() > 0.5, "new goals were drawn"
    st = env.stats()
    assert st["episodes"] >= n and np.isfinite(new).all()
    q, _, _ = env.get_state()
    assert np.all(q[:, 2] > -0.46) and np.all(q[:, 2] < -0.40), "robots rest on the maze floor at z = -0.5"
    env.close()


@pytest.mark.parametrize("model,lanes", [("v2", 1), ("v2", 4), ("v2", 8), ("scene", 1), ("scene", 4), ("maze:umaze", 1), ("maze:umaze", 4)])
def test_odd_batch_sizes_and_layouts_run_clean(model, lanes):
    """Ragged batches (partial warps / CTAs, a single environment) in every layout, with auto-reset and host buffers: finite outputs,
    consistent counters, and environment 0 independent of the batch size."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    first = None
    for n in (1, 31, 33, 127, 1025):
        env = BatchedAckermannEnv(n, model=model, dtype="float32", seed=17, lanes_per_env=lanes, frame_skip=2, max_episode_steps=5, auto_reset=True)
        env.reset()
        for _ in range(7):
            obs, rew, term, trunc, info = env.step(None)
        h = [torch.empty((n, env.obs_dim)).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory(),
             torch.empty(n, dtype=torch.uint8).pin_memory()]
        env.step_host(torch.zeros(n, 2).pin_memory(), *h)
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and torch.isfinite(h[0]).all()
        st = env.stats()
        assert st["env_steps"] == 8 * n and st["episodes"] >= n
        q, v, _ = env.get_state()
        assert np.isfinite(q).all() and np.isfinite(v).all()
        if first is None:
            first = q[0].copy()
        else:
            # fp32: whether environment 0 finishes a solve in the plain or in the tail layout depends on its warp mates, which
            # changes summation order at the solver tolerance; 16 substeps of the hopping robot amplify that to < 1e-3
            assert np.abs(q[0] - first).max() < 5e-3, "environment 0 does not depend on the batch size (up to fp32 solver tolerance)"
        env.close()


def test_sharded_batch_equals_single_batch_bit_for_bit():
    """random streams are keyed by the GLOBAL environment id, so rank r of R owning envs [r N/R, (r+1) N/R) with the same
    seed reproduces the single-handle batch exactly (goals, synthetic actions, resets) -- results do not depend on R."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    from mujoco_playground_b200.shard import shard_range
    n, steps = 1000, 40
    kw = dict(dtype="float32", seed=77, frame_skip=2, max_episode_steps=15, lanes_per_env=4)     # 15-step episodes: auto-resets inside the run
    full = BatchedAckermannEnv(n, **kw)
    full.reset()
    shards = []
    for r in range(3):
        b, e = shard_range(n, r, 3)
        s = BatchedAckermannEnv(e - b, env_id_base=b, **kw)
        s.reset()
        shards.append((b, e, s))
    for t in range(steps):
        fo, fr, ft, ftr, _ = full.step(None)
        for b, e, s in shards:
            so, sr, st, strn, _ = s.step(None)
            assert torch.equal(fo[b:e], so) and torch.equal(fr[b:e], sr) and torch.equal(ft[b:e], st) and torch.equal(ftr[b:e], strn), f"step {t} shard {b}"
    qf = full.get_state()[0]
    for b, e, s in shards:
        assert np.array_equal(qf[b:e], s.get_state()[0])
        assert np.array_equal(full.get_episode()[0][b:e], s.get_episode()[0])
        s.close()
    assert full.stats()["episodes"] >= 2 * n
    full.close()


def test_interleaved_handles_of_different_models_and_precisions():
    """Each handle carries its own constants (kernel parameter): alternating calls on handles of different models / precisions on
    one device give the same results as running each handle alone (round 1 shared one __constant__ block per device)."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    specs = [dict(model="v2", dtype="float32"), dict(model="scene", dtype="float32"), dict(model="v2", dtype="float64"), dict(model="maze:umaze", dtype="float32")]
    n, steps = 130, 12
    alone = []
    for sp in specs:
        e = BatchedAckermannEnv(n, seed=5, **sp)
        e.reset()
        for _ in range(steps):
            obs = e.step(None)[0]
        alone.append(obs.clone())
        e.close()
    envs = [BatchedAckermannEnv(n, seed=5, **sp) for sp in specs]
    for e in envs:
        e.reset()
    for _ in range(steps):
        outs = [e.step(None)[0] for e in envs]
    for a, b, sp in zip(alone, outs, specs):
        assert torch.equal(a, b), sp
    for e in envs:
        e.close()


@pytest.mark.parametrize("lanes", [1, 4])
def test_bad_state_guard_on_device(lanes):
    """mj_checkPos / mj_checkVel: poisoned environments are reset to qpos0 inside the step (counted in stats.bad_state), everything
    stays finite and the other environments are untouched."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    n = 64
    ref = BatchedAckermannEnv(n, dtype="float32", seed=2, lanes_per_env=lanes, auto_reset=False)
    env = BatchedAckermannEnv(n, dtype="float32", seed=2, lanes_per_env=lanes, auto_reset=False)
    ref.reset(); env.reset()
    for _ in range(20):
        ref.step(None); env.step(None)
    q, v, w = env.get_state()
    q[3, 0] = np.nan; v[10, 7] = np.inf; q[40, 9] = 5e10
    env.set_state(q, v, w)
    o1 = env.step(None)[0]
    o0 = ref.step(None)[0]
    assert env.stats()["bad_state"] == 3 and ref.stats()["bad_state"] == 0
    assert torch.isfinite(o1).all()
    keep = [i for i in range(n) if i not in (3, 10, 40)]
    assert torch.equal(o1[keep], o0[keep])
    q2 = env.get_state()[0]
    assert np.isfinite(q2).all() and abs(q2[3, 2] - 0.065) < 1e-3, "reset to qpos0 (chassis at z = 0.065), then one substep"
    ref.close(); env.close()


def test_evaluate_agent_with_scripted_policies():
    """sb3_io.evaluate_agent (src/rl/utils.py:20-50) against policies whose outcome is known: a goal-seeking controller built from
    the observation's goal bearing reaches most goals on the empty floor; a policy that stands still reaches none and every episode
    runs into the time limit; episode statistics come from the device-side counters."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    from mujoco_playground_b200.sb3_io import evaluate_agent
    # 10 substeps (20 ms) per step, 800 steps = 16 s of driving at <= 1 m/s: enough for goals 2 - 8 m away
    env = BatchedAckermannEnv(512, dtype="float32", seed=6, frame_skip=10, max_episode_steps=800)

    def seek(obs):      # obs[:, 78] = wrapped bearing of the goal, obs[:, 77] = distance (ackermann_env.py:248-263)
        ang = obs[:, 78]
        return torch.stack([torch.where(ang.abs() < 1.0, torch.ones_like(ang), 0.4 * torch.ones_like(ang)), torch.clamp(2.0 * ang, -1, 1)], dim=1)

    good = evaluate_agent(env, seek, n_steps=800)
    assert good["episodes"] >= 512 and good["success_rate"] > 0.9, good
    assert good["mean_length"] < 700
    idle = evaluate_agent(env, lambda obs: torch.zeros((obs.shape[0], 2), device=obs.device), n_steps=800)
    assert idle["episodes"] == 512 and idle["success_rate"] == 0.0 and abs(idle["mean_length"] - 800) < 1e-6, idle
    # n_episodes mode (the reference's argument): stops once that many episodes have finished
    few = evaluate_agent(env, seek, n_steps=800, n_episodes=100)
    assert 100 <= few["episodes"] <= 512 * 3
    env.close()


def test_vec_env_contract():
    """AckermannB200VecEnv: the SB3 VecEnv contract of the reference's DummyVecEnv(Monitor(env)) (src/rl/train.py:70-76)."""
    from mujoco_playground_b200 import AckermannB200VecEnv
    venv = AckermannB200VecEnv(40, seed=3, max_episode_steps=7)
    assert venv.num_envs == 40 and venv.observation_space.shape == (79,) and venv.action_space.shape == (2,)
    assert venv.action_space.low.min() == -1 and venv.action_space.high.max() == 1 and np.isinf(venv.observation_space.high).all()
    obs = venv.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (40, 79) and obs.dtype == np.float32
    rng = np.random.default_rng(0)
    ret = np.zeros(40)
    for t in range(7):
        obs, rew, dones, infos = venv.step(rng.uniform(-2, 2, (40, 2)))       # out-of-range actions are clipped like SB3 does
        ret += rew
        assert rew.shape == (40,) and dones.dtype == bool and len(infos) == 40
        if t < 6:
            assert not dones.any() and all(i == {} for i in infos)
    assert dones.all(), "time limit after max_episode_steps"
    for i, info in enumerate(infos):
        assert info["TimeLimit.truncated"] is True and info["terminal_observation"].shape == (79,)
        assert info["episode"]["l"] == 7 and abs(info["episode"]["r"] - ret[i]) < 1e-3 * max(1.0, abs(ret[i]))
    assert np.allclose(obs[:, 72:75], 0.0, atol=1e-6), "auto-reset: the returned observation is the first of the new episode"
    assert venv.env_is_wrapped(None) == [False] * 40 and venv.get_attr("frame_skip")[0] == 1
    venv.close()


def test_single_env_adapter_spaces_and_reseed():
    """AckermannRobotEnv: spaces of the reference (ackermann_env.py:95-108); reset(seed=s) restarts the random streams."""
    from mujoco_playground_b200 import AckermannRobotEnv
    env = AckermannRobotEnv(dtype="float64")
    assert env.observation_space.shape == (79,) and env.observation_space.dtype == np.float32
    assert env.action_space.contains(env.action_space.sample())
    o1, i1 = env.reset(seed=11)
    o2, i2 = env.reset(seed=11)
    o3, i3 = env.reset(seed=12)
    assert np.array_equal(o1, o2) and i1["goal_position"] == i2["goal_position"] and i1["goal_position"] != i3["goal_position"]
    assert env.observation_space.contains(o1)
    env.close()


@pytest.mark.parametrize("dtype,tol,lanes", [("float64", 1e-6, 4), ("float64", 1e-6, 1), ("float32", 2e-3, 4), ("float32", 2e-3, 1)])
def test_rollover_regime_split_matches_oracle(dtype, tol, lanes):
    """Flat-floor model, tumbling robots mixed with upright ones in one batch: the fast kernel steps the upright environments, the
    general pass (NC = 4) the tilted ones (wheel caps, chassis plates on the floor).  Per step, every environment restarted from the
    oracle's state: contact counts exact (fp64), velocities within tolerance, unsupported = 0."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    from oracle.oracle import OracleSim
    M = _models()

    def quat(axis, deg):
        a = np.deg2rad(deg) / 2
        return np.array([np.cos(a), *(np.sin(a) * np.asarray(axis, float))])
    rng = np.random.default_rng(3)
    poses = [(quat([1, 0, 0], 180), 0.12), (quat([0, 1, 0], 80), 0.2), (quat([1, 0, 0], 90), 0.13), (quat([0, 1, 0], -100), 0.25),
             (quat([0, 0, 1], 30), 0.1), (quat([1, 0, 0], 10), 0.1)]
    for _ in range(4):
        q = rng.normal(size=4)
        poses.append((q / np.linalg.norm(q), 0.25))
    n = len(poses) * 4        # every pose four times: the warps mix both regimes
    env = BatchedAckermannEnv(n, dtype=dtype, auto_reset=False, lanes_per_env=lanes, solver_tolerance=1e-12 if dtype == "float64" else None)
    env.reset()
    sims = []
    for i in range(n):
        q, z = poses[i % len(poses)]
        o = OracleSim(M, tolerance=1e-12)
        o.reset()
        o.qpos[:3] = [0.3 * i, 0, z]; o.qpos[3:7] = q; o.qvel[3:6] = [1.0, -2.0, 0.5]
        sims.append(o)
    zero = torch.zeros((n, 2), device="cuda:0")
    flips, maxn = 0, 0
    for t in range(150):
        env.set_state(np.stack([o.qpos for o in sims]), np.stack([o.qvel for o in sims]), np.stack([o.qacc_warmstart for o in sims]))
        _, _, _, _, info = env.step(zero)
        qv = env.get_state()[1]
        ncon = info["ncon"].cpu().numpy()
        for i, o in enumerate(sims):
            o.ctrl[:] = 0      # zero action: bicycle controller gives zero controls
            o.step()
            flips += int(ncon[i] != o.ncon)
            maxn = max(maxn, o.ncon)
            if dtype == "float64":
                assert ncon[i] == o.ncon, f"step {t} env {i}: ncon {ncon[i]} vs oracle {o.ncon}"
            assert np.abs(qv[i] - o.qvel).max() < tol * max(1.0, np.abs(o.qvel).max()), f"step {t} env {i}"
    st = env.stats()
    print(f"roll-over {dtype} lanes={lanes}: contact-count flips {flips}/{150 * n}, max ncon {maxn}, unsupported {st['unsupported']}")
    assert st["unsupported"] == 0 and maxn >= 11
    assert flips == 0 if dtype == "float64" else flips <= 0.02 * 150 * n
    env.close()


def test_maze_wall_hits_use_plate_contacts():
    """maze:umaze, full throttle into the walls: the plates reach the blocks before the wheels do; the batch stays supported and the
    environments that drove into a wall stop there (no tunnelling through the block with the chassis)."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    env = BatchedAckermannEnv(512, model="maze:umaze", dtype="float32", seed=9, frame_skip=4, auto_reset=False)
    env.reset()
    act = torch.zeros((512, 2), device="cuda:0")
    act[:, 0] = 1.0
    maxn = 0
    for _ in range(400):
        _, _, _, _, info = env.step(act)
        maxn = max(maxn, int(info["ncon"].max().item()))
    st = env.stats()
    q = env.get_state()[0]
    assert st["unsupported"] == 0 and st["bad_state"] == 0
    assert maxn >= 9, "wheel contacts plus at least one plate-vs-block contact"
    assert np.isfinite(q).all() and (np.abs(q[:, 2] + 0.435) < 0.05).all(), "robots stay on the maze floor (z about -0.435)"
    env.close()


def _tumble_poses():
    def quat(axis, deg):
        a = np.deg2rad(deg) / 2
        return np.array([np.cos(a), *(np.sin(a) * np.asarray(axis, float))])
    return [(quat([1, 0, 0], 180), 0.12), (quat([0, 1, 0], 80), 0.2), (quat([1, 0, 0], 90), 0.13), (quat([0, 1, 0], -100), 0.25),
            (quat([0, 0, 1], 30), 0.1), (quat([1, 0, 0], 10), 0.1), (quat([1, 1, 0] / np.sqrt(2.0), 120), 0.2)]


@pytest.mark.parametrize("lanes", [1, 4])
def test_general_pass_list_longer_than_its_grid(lanes):
    """Regime split of the flat-floor model: the fast pass LISTS the tilted environments and the general pass is a fixed grid of one
    CTA per SM (32 environments each) walking that list.  8192 environments, five of seven poses tilted: the list (5852 entries) needs
    the grid-stride loop, its order is whatever the atomics gave, and the listed environments are scattered through the batch.
    Every copy of a pose must stay bit-identical to the first one (the same property a 70-environment batch, whose list fits one
    CTA pair, is checked for against the oracle in test_rollover_regime_split_matches_oracle), and the list must be empty again
    after every step (an upright batch stepped afterwards is not touched by stale entries)."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    M = _models()
    poses = _tumble_poses()
    n, P = 8192, len(poses)
    env = BatchedAckermannEnv(n, dtype="float32", auto_reset=False, lanes_per_env=lanes)
    small = BatchedAckermannEnv(P, dtype="float32", auto_reset=False, lanes_per_env=4)
    for e, m in ((env, n), (small, P)):
        e.reset()
        qpos, qvel = np.tile(M["qpos0"], (m, 1)), np.zeros((m, 12))
        for i in range(m):
            q, z = poses[i % P]
            qpos[i, 2] = z; qpos[i, 3:7] = q; qvel[i, 3:6] = [1.0, -2.0, 0.5]
        e.set_state(qpos, qvel, np.zeros((m, 12)))
    act = torch.zeros((n, 2), device="cuda:0")
    act[:, 0] = 0.3
    for t in range(12):
        env.step(act)
        small.step(act[:P].contiguous())
        q, v, _ = env.get_state()
        qs, vs, _ = small.get_state()
        for p in range(P):
            assert (q[p::P] == q[p]).all() and (v[p::P] == v[p]).all(), f"step {t}: copies of pose {p} differ"
        if lanes == 4:      # the small batch runs the 4-lane fast kernel: same arithmetic as the big one only in that layout
            assert np.array_equal(q[:P], qs) and np.array_equal(v[:P], vs), f"step {t}: 8192-env batch differs from the {P}-env batch"
    assert env.stats()["bad_state"] == 0
    # back to an upright batch: nothing may be left in the list
    qpos, qvel = np.tile(M["qpos0"], (n, 1)), np.zeros((n, 12))
    env.set_state(qpos, qvel, np.zeros((n, 12)))
    for _ in range(3):
        env.step(act)
    q, v, _ = env.get_state()
    assert (q == q[0]).all() and (v == v[0]).all()
    env.close(); small.close()


def test_step_replays_from_a_cuda_graph():
    """ackb_step is stream-capture safe (constants travel as a kernel parameter, the general pass is a programmatic dependent launch,
    its list counter is reset on the device): a step captured once and replayed equals the same steps launched one by one, bit for
    bit, including environments that go through the general pass and through the fused auto-reset."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    M = _models()
    poses = _tumble_poses()
    n, P = 512, len(poses)
    envs = [BatchedAckermannEnv(n, dtype="float32", auto_reset=True, seed=5, max_episode_steps=25, lanes_per_env=4) for _ in range(2)]
    qpos, qvel = np.tile(M["qpos0"], (n, 1)), np.zeros((n, 12))
    for i in range(0, n, 3):
        q, z = poses[i % P]
        qpos[i, 2] = z; qpos[i, 3:7] = q; qvel[i, 3:6] = [1.0, -2.0, 0.5]
    act = torch.zeros((n, 2), device="cuda:0")
    act[:, 0] = 0.5; act[:, 1] = 0.2
    for e in envs:
        e.reset()
        e.set_state(qpos, qvel, np.zeros((n, 12)))
        e.step(act)                      # warm-up: function attributes are set outside the capture
    eager, graphed = envs
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        graphed.step(act)
    # the captured launch did not run: both environments are one step in
    for t in range(60):
        eager.step(act)
        g.replay()
    torch.cuda.synchronize()
    qa, va, _ = eager.get_state()
    qb, vb, _ = graphed.get_state()
    assert np.array_equal(qa, qb) and np.array_equal(va, vb)
    assert torch.equal(eager.obs, graphed.obs) and torch.equal(eager.reward, graphed.reward)
    sa, sb = eager.stats(), graphed.stats()
    assert sa["episodes"] == sb["episodes"] and sa["episodes"] >= 2 * n and sa["unsupported"] == sb["unsupported"] == 0
    for e in envs:
        e.close()
