"""The CPU oracle (oracle/ackb_oracle.c) against analytic invariants and against the lidar known-answer that the reference's
checkpoints contain.  The reference has no tests or golden vectors for the physics (parity
unpinned), so these are the pins available."""
import numpy as np
import pytest

from mujoco_playground_b200.compiler.setconst import forward_kinematics, mass_matrix
from mujoco_playground_b200.models import load_model
from oracle.oracle import OracleSim

M = load_model("v2")


def test_free_fall_and_contact_onset():
    s = OracleSim(M)
    s.qpos[2] = 0.1
    z = []
    for i in range(60):
        s.step()
        z.append((s.qpos[2], s.ncon))
    # semi-implicit Euler free fall: z_k = z0 - g h^2 k(k+1)/2 while airborne
    for k in (1, 10, 30):
        assert abs(z[k - 1][0] - (0.1 - 9.81 * 0.002 ** 2 * k * (k + 1) / 2)) < 1e-12
    first = next(i for i, (_, n) in enumerate(z) if n > 0)
    assert first in (41, 42, 43), "35 mm drop: wheels reach the floor after ~0.084 s"
    assert z[first][1] == 8, "upright wheels: two rim contacts each"


def test_static_equilibrium_supports_weight():
    s = OracleSim(M, tolerance=1e-12)
    s.qpos[2] = 0.0655
    s.step(1500)
    assert s.ncon == 8 and s.nefc == 39
    f = s.efc("force")[-32:]
    # each pyramid row force acts along n +- mu t: the normal components sum to the weight
    assert abs(f.sum() - 10.3 * 9.81) < 1e-6
    assert np.abs(s.qvel).max() < 1e-8
    assert 0.0640 < s.qpos[2] < 0.0650


def test_mass_matrix_matches_jacobian_construction():
    rng = np.random.default_rng(0)
    s = OracleSim(M)
    for _ in range(5):
        q = rng.normal(size=4)
        s.qpos[:3] = rng.uniform(-2, 2, 3)
        s.qpos[3:7] = q / np.linalg.norm(q)
        s.qpos[7:] = rng.uniform(-1, 1, 6)
        s.qpos[2] += 3.0
        s.forward()
        Mm = mass_matrix(M, forward_kinematics(M, s.qpos.copy()))
        assert np.abs(s.qM() - Mm).max() < 1e-12
        w = np.linalg.eigvalsh(s.qM())
        assert w.min() > 0


def test_momentum_conserved_in_free_flight():
    """No contacts, no gravity-free trick needed: total linear momentum changes by m g h per step, and the
    chassis angular momentum about the COM is conserved when the hinges are frozen by zero relative velocity."""
    s = OracleSim(M)
    s.qpos[2] = 50.0
    s.qvel[:3] = [0.3, -0.2, 0.1]
    s.qvel[3:6] = [0.0, 0.0, 0.0]
    s.forward()
    p = []
    for _ in range(50):
        s.step()
        # momentum = M(q) qvel restricted to the translational rows (world frame)
        p.append((s.qM() @ s.qvel)[:3].copy())
    dp = np.diff(np.array(p), axis=0)
    assert np.abs(dp[:, :2]).max() < 1e-10
    assert np.abs(dp[:, 2] + 10.3 * 9.81 * 0.002).max() < 1e-9


def test_quaternion_stays_normalised_and_energy_bounded_on_ground():
    s = OracleSim(M)
    s.qpos[2] = 0.1
    s.ctrl[:] = [0.2, 20, 20]
    s.step(2000)
    assert abs(np.linalg.norm(s.qpos[3:7]) - 1) < 1e-12
    assert np.isfinite(s.qpos).all() and abs(s.qpos[2] - 0.0645) < 2e-3


def test_straight_line_drive_is_bounded_by_servo_target():
    """Velocity servos (kv=1, target 15 rad/s) on both rear wheels, zero steer.  The model's joint damping
    (0.15 / 0.12 N m s/rad on 32 mm wheels) is large: the rear wheels settle below the target (torque balance) and the
    undriven front wheels skid, so only bounds are asserted: forward motion, below the rolling speed of the target."""
    s = OracleSim(M)
    s.qpos[2] = 0.066
    s.ctrl[:] = [0.0, 15.0, 15.0]
    s.step(3000)
    v = s.qvel[0]
    assert 0.25 < v < 0.0325 * 15
    assert 10.0 < s.qvel[6] < 15.0 and 10.0 < s.qvel[7] < 15.0
    assert v < 0.0325 * s.qvel[6] * 1.001, "traction comes from the rear wheels: they cannot turn slower than rolling"
    assert s.qpos[0] > 1.5 and abs(s.qpos[1]) < 0.3


def _add_box(Mx, pos, half):
    """Append a static body with one box geom to a table model (all geom_* / body_* arrays grow by one row)."""
    nb, ng = Mx["nbody"], Mx["ngeom"]

    def app(key, row):
        a = Mx[key]
        Mx[key] = np.concatenate([a, np.asarray(row, dtype=a.dtype).reshape((1,) + a.shape[1:])])
    for key, row in (("body_parentid", 0), ("body_rootid", nb), ("body_weldid", 0), ("body_jntnum", 0), ("body_jntadr", -1),
                     ("body_dofnum", 0), ("body_dofadr", -1), ("body_mass", 0.0)):
        app(key, [row])
    for key, row in (("body_pos", [0, 0, 0]), ("body_ipos", [0, 0, 0]), ("body_inertia", [0, 0, 0]), ("body_quat", [1, 0, 0, 0]),
                     ("body_iquat", [1, 0, 0, 0]), ("body_invweight0", [0, 0])):
        app(key, row)
    for key, row in (("geom_type", 6), ("geom_bodyid", nb), ("geom_contype", 1), ("geom_conaffinity", 1), ("geom_condim", 3),
                     ("geom_priority", 0), ("geom_hulladr", -1), ("geom_hullnum", 0), ("geom_solmix", 1.0), ("geom_margin", 0.0),
                     ("geom_gap", 0.0), ("geom_alpha", 1.0)):
        app(key, [row])
    for key, row in (("geom_size", half), ("geom_pos", pos), ("geom_quat", [1, 0, 0, 0]), ("geom_friction", [1, 0.005, 0.0001]),
                     ("geom_solref", [0.02, 1]), ("geom_solimp", [0.9, 0.95, 0.001, 0.5, 2])):
        app(key, row)
    Mx["nbody"], Mx["ngeom"] = nb + 1, ng + 1


# beams 10..71 of rl_logs/ppo/ppo_model_10000_steps.zip -> data._last_obs (reference run in a
# PointMaze U-maze with 1 m cells, heading 9e-5 rad); reconstructing the hit points gives: right wall x=+0.6539, top wall
# y=+1.6781, bottom wall y=-1.3219, and the 1 m x 1 m centre block of the U on the left: x in [-1.3461, -0.3461], y in [-0.3219, 0.6781].
D1 = [0.98233, 1.10511, 1.27292, 1.51245, 1.75075, 1.70227, 1.66897, 1.64951, 1.64311, 1.64954, 1.66903, 1.70235, 1.75087, 1.81667,
      0.65718, 0.56840, 0.50344, 0.45447, 0.41682, 0.38753, 0.36467, 0.34691, 0.33334, 0.32334, 0.31647, 0.31246, 0.31114, 0.31246,
      0.31648, 0.32336, 0.33336, 0.34694, 0.36471, 0.38758, 0.41688, 1.83427, 1.69048, 1.57863, 1.49131, 1.42348, 1.37168, 1.33349,
      1.30726, 1.29193, 1.28689, 1.29195, 1.30730, 1.33355, 1.37177, 1.42360, 1.27253, 1.10483, 0.98212, 0.88962, 0.81849, 0.76317,
      0.71998, 0.68643, 0.66080, 0.64191, 0.62894, 0.62135]


def test_lidar_known_answer_from_reference_checkpoint():
    Mx = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in M.items()}
    t = 0.5
    _add_box(Mx, [0.6539 + t, 0.0, 0.5], [t, 3.0, 1.0])
    _add_box(Mx, [-0.3461 - t, 0.1781, 0.5], [t, 0.5, 1.0])
    _add_box(Mx, [0.0, 1.6781 + t, 0.5], [3.0, t, 1.0])
    _add_box(Mx, [0.0, -1.3219 - t, 0.5], [3.0, t, 1.0])
    s = OracleSim(Mx)
    yaw = 9e-5
    s.qpos[:3] = [0, 0, 0.065]
    s.qpos[3:7] = [np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
    s.forward()
    lidar = s.sensordata[5:77]
    np.testing.assert_allclose(lidar[10:72], D1, atol=2e-4)
    # the observation the reference would have built: slots 0..9 read beam 71 (checkpoint has obs[0:10] == 0.62135)
    from oracle.env_oracle import lidar_addrs_reference
    obs = np.array([s.sensordata[a] for a in lidar_addrs_reference(M)])
    assert np.allclose(obs[:10], 0.62135, atol=2e-4)


def test_scene_model_tables():
    S = load_model("scene")
    assert (S["nq"], S["nv"], S["nu"], S["nbody"], S["nsensordata"], S["neq"]) == (13, 12, 4, 49, 42, 0)
    occ = 0
    for g in range(S["ngeom"]):
        if S["geom_type"][g] == 6:
            x, y = S["geom_pos"][g][:2]
            occ |= 1 << (int(y + 4) * 8 + int(x + 4))
    assert occ == 0xff89a591c79199ff
    s = OracleSim(S)
    s.step(500)
    assert s.ncon == 8 and abs(s.qpos[2] - 0.0648) < 1e-3
    lidar = s.sensordata[6:42]
    assert (lidar > 0).all() and lidar.max() < 8.0, "36 beams at z=0.10 hit the 0.15 m high maze walls"


def test_rest_on_wheel_caps_and_on_chassis_plates_supports_weight():
    """Statics of the contacts beyond the eight rim points (no MuJoCo needed: at rest the normal components of the pyramid-row forces
    add up to the weight, whatever the impedance parameters).  On its side the robot lies on two wheel caps: rim points plus the two
    cap-down "triangle" points each.  Nose down it stands on the front edges of the chassis plates (convex hulls of the meshes: support
    vertex + hull-graph neighbours); that pose chatters between 4 and 6 contacts, so the force is averaged over time and the robot
    must neither sink through its plates nor fly off."""
    def quat(axis, deg):
        a = np.deg2rad(deg) / 2
        return np.array([np.cos(a), *(np.sin(a) * np.asarray(axis, float))])
    weight = 10.3 * 9.81

    def contact_force(s):
        ninc = sum(1 for c in s.contacts() if not c["exclude"])
        return s.efc("force")[-4 * ninc:].sum() if ninc else 0.0

    s = OracleSim(M, tolerance=1e-12)
    s.qpos[2] = 0.14; s.qpos[3:7] = quat([1, 0, 0], 90)
    s.step(6000)
    geoms = {(c["geom1"], c["geom2"]) for c in s.contacts()}
    assert s.ncon == 6 and len(geoms) == 2 and all(g1 == 0 and g2 >= 3 for g1, g2 in geoms), "two wheels, three points each"
    assert abs(contact_force(s) - weight) < 1e-6 and np.abs(s.qvel).max() < 1e-6

    s = OracleSim(M, tolerance=1e-12)
    s.qpos[2] = 0.22; s.qpos[3:7] = quat([0, 1, 0], 88)
    s.step(5000)
    forces, heights, counts = [], [], set()
    for _ in range(2000):
        s.step()
        forces.append(contact_force(s)); heights.append(s.qpos[2]); counts.add(s.ncon)
        assert {(c["geom1"], c["geom2"]) for c in s.contacts()} <= {(0, 1), (0, 2)}, "only the two plates touch the floor"
    assert abs(np.mean(forces) - weight) < 0.005 * weight
    assert max(heights) - min(heights) < 1e-4 and abs(np.mean(heights) - 0.1499) < 2e-4
    assert np.abs(s.qvel).max() < 2e-2 and counts <= {3, 4, 5, 6}
