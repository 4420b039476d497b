"""Physics parity against REAL MuJoCo golden vectors (north_star: "single-step qpos/qvel within 1e-5 relative in
fp64 mode and 1e-3 in fp32 ... integer contact counts must match exactly").

The golden file is produced by `tools/dump_golden.py` on a machine that has the third-party `mujoco` package (neither this build
image nor the GPU box has it) and committed as tests/golden/mujoco_*.npz.  Without such a file every test below that needs it
reports  SKIPPED (no oracle)  -- never "passed".  `test_golden_format_selfcheck` runs the same consumer code on a file written by
`tools/dump_golden.py --backend oracle` (this repo's own CPU oracle): it proves that a real file will be consumed correctly, and
nothing about MuJoCo.
"""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tools"))
from dump_mjmodel import load_table  # noqa: E402

GOLDEN_GLOB = os.path.join(ROOT, "tests", "golden", "mujoco_*.npz")
SKIP_MSG = ("SKIPPED (no oracle): no tests/golden/mujoco_*.npz -- run tools/dump_golden.py where the `mujoco` package is installed "
            "(it is not installable in this image) and commit the file")


def find_golden():
    for p in sorted(glob.glob(GOLDEN_GLOB)):
        with np.load(p, allow_pickle=False) as z:
            if str(z["backend"][0]) == "mujoco":
                return p
    return None


def scenario(G, name):
    """(table, arrays) of one scenario of an opened golden npz."""
    T = load_table(G, prefix=f"{name}/model/")
    A = {k[len(name) + 1:]: G[k] for k in G.files if k.startswith(name + "/") and not k.startswith(name + "/model/") and "/efc/" not in k}
    return T, A


def kind_of(A):
    return str(A["kind"][0])


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


# ---- consumers (shared by the real-golden tests and the format self-check) ---------------------------------------------------------
def check_oracle_single_steps(T, A, steps, tol=1e-5):
    """Oracle restarted from every golden state: one step each, qpos / qvel within `tol` relative, contact count exact."""
    from oracle.oracle import OracleSim
    o = OracleSim(T, tolerance=1e-10)
    worst, flips = 0.0, 0
    for t in range(steps):
        o.reset()
        o.qpos[:] = A["qpos"][t]; o.qvel[:] = A["qvel"][t]
        o.qacc_warmstart[:] = A["qacc_warmstart"][t - 1] if t > 0 else 0.0
        o.ctrl[:] = A["ctrl"][t]
        o.step()
        flips += int(o.ncon != int(A["ncon"][t]))
        worst = max(worst, rel_err(o.qpos, A["qpos"][t + 1]), rel_err(o.qvel, A["qvel"][t + 1]))
    return worst, flips


def check_oracle_horizon(T, A, horizon):
    """Free-running oracle from the golden initial state over a short horizon (<= 100 steps): worst relative qpos error."""
    from oracle.oracle import OracleSim
    o = OracleSim(T, tolerance=1e-10)
    o.reset()
    o.qpos[:] = A["qpos"][0]; o.qvel[:] = A["qvel"][0]
    worst = 0.0
    for t in range(horizon):
        o.ctrl[:] = A["ctrl"][t]
        o.step()
        worst = max(worst, rel_err(o.qpos, A["qpos"][t + 1]))
    return worst


def check_efc(G, name, T, A, steps):
    """Constraint rows of the first steps: row count, regulariser and reference acceleration (settles the pyramidal R / diagApprox items)."""
    from oracle.oracle import OracleSim
    o = OracleSim(T, tolerance=1e-10)
    worst = 0.0
    for t in range(steps):
        if f"{name}/efc/{t}/efc_D" not in G.files:
            break
        o.reset()
        o.qpos[:] = A["qpos"][t]; o.qvel[:] = A["qvel"][t]; o.ctrl[:] = A["ctrl"][t]
        o.qacc_warmstart[:] = A["qacc_warmstart"][t - 1] if t > 0 else 0.0
        o.forward()
        D, aref = G[f"{name}/efc/{t}/efc_D"], G[f"{name}/efc/{t}/efc_aref"]
        assert o.nefc == len(D), f"{name} step {t}: nefc {o.nefc} vs golden {len(D)}"
        if len(D):
            worst = max(worst, float(np.abs(o.efc("D") / D - 1).max()), rel_err(o.efc("aref"), aref))
    return worst


def check_model_constants(T_gold, kind):
    """Compiled table (this repo's MJCF compiler) vs the table dumped from MuJoCo: the version-dependent constants."""
    from mujoco_playground_b200.models import load_model
    M = load_model({"v2": "v2", "scene": "scene"}[kind])
    out = {}
    for k in ("body_mass", "body_ipos", "body_inertia", "body_invweight0", "dof_invweight0", "stat_meaninertia", "qM0", "qpos0", "geom_size",
              "geom_pos", "geom_friction", "geom_solref", "geom_solimp", "site_pos", "actuator_gainprm", "actuator_biasprm"):
        a, b = np.asarray(M[k], float), np.asarray(T_gold[k], float)
        out[k] = float("inf") if a.shape != b.shape else rel_err(a, b)
    out["sensor_names_equal"] = list(M["sensor_names"]) == list(T_gold["sensor_names"])
    out["hull_vertex_count_equal"] = len(M["hull_vert"]) == len(T_gold["hull_vert"])
    return out


def run_kernel_single_steps(T, A, kind, dtype, steps):
    """One CUDA env per golden step: state t and action t in, one env.step(), state t+1 out.  Returns (worst rel error, ncon flips)."""
    import torch
    from mujoco_playground_b200 import BatchedAckermannEnv
    n = steps
    has_box = bool((np.asarray(T["geom_type"]) == 6).any())
    env = BatchedAckermannEnv(n, device="cuda:0", dtype=dtype, auto_reset=False, model="scene" if (kind == "scene" or has_box) else "v2",
                              model_table=T, solver_tolerance=1e-10 if dtype == "float64" else None)
    env.reset()
    warm = np.zeros((n, 12))
    warm[1:] = A["qacc_warmstart"][: n - 1]
    env.set_state(A["qpos"][:n], A["qvel"][:n], warm)
    _, _, _, _, info = env.step(torch.from_numpy(np.ascontiguousarray(A["act"][:n])).cuda())
    qpos, qvel, _ = env.get_state()
    ncon = info["ncon"].cpu().numpy()
    env.close()
    worst = max(rel_err(qpos[i], A["qpos"][i + 1]) for i in range(n))
    worst = max(worst, max(rel_err(qvel[i], A["qvel"][i + 1]) for i in range(n)))
    return worst, int((ncon != A["ncon"][:n]).sum())


# ---- real golden vectors -------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def golden():
    p = find_golden()
    if p is None:
        pytest.skip(SKIP_MSG)
    return np.load(p, allow_pickle=False)


def test_mujoco_model_constants(golden):
    for name in ("cfg1", "scene"):
        T, A = scenario(golden, name)
        diff = check_model_constants(T, kind_of(A))
        bad = {k: v for k, v in diff.items() if (v is False) or (not isinstance(v, bool) and v > 1e-6)}
        assert not bad, f"{name}: compiled constants differ from MuJoCo {golden['mujoco_version'][0]}: {bad}"


@pytest.mark.parametrize("name", ["cfg1", "scene", "rollover", "nosedown", "onside", "wallhit"])
def test_mujoco_oracle_parity(golden, name):
    """The CPU oracle against mj_step on MuJoCo's own compiled constants: this is what turns the oracle from 'unpinned' to 'pinned'."""
    T, A = scenario(golden, name)
    steps = min(len(A["act"]), 300)
    worst, flips = check_oracle_single_steps(T, A, steps)
    assert flips == 0, f"{name}: {flips} of {steps} contact counts differ from MuJoCo"
    assert worst < 1e-5, f"{name}: single-step error {worst:.2e}"
    assert check_efc(golden, name, T, A, 50) < 1e-6
    assert check_oracle_horizon(T, A, min(100, steps)) < 1e-3, "100-step trajectory tolerance (chaotic divergence beyond is documented)"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1", "scene", "rollover", "nosedown", "onside", "wallhit"])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-5), ("float32", 1e-3)])
def test_mujoco_kernel_parity(golden, name, dtype, tol):
    """The CUDA kernels against mj_step: north_star tolerances, contact counts exact in fp64 (fp32 flip rate reported and bounded)."""
    T, A = scenario(golden, name)
    steps = min(len(A["act"]), 300)
    worst, flips = run_kernel_single_steps(T, A, kind_of(A), dtype, steps)
    print(f"{name} {dtype}: single-step error {worst:.2e}, contact-count flips {flips}/{steps}")
    assert worst < tol
    assert flips == 0 if dtype == "float64" else flips <= max(1, steps // 50)


def test_golden_presence_is_reported():
    """Makes the state of the pin visible in every test log."""
    p = find_golden()
    if p is None:
        pytest.skip(SKIP_MSG)
    with np.load(p) as z:
        print(f"MuJoCo golden vectors: {os.path.basename(p)} (mujoco {z['mujoco_version'][0]}), scenarios {list(z['scenarios'])}")


# ---- format self-check (no MuJoCo involved) -------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def self_golden(tmp_path_factory):
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("the reference checkout (MJCF sources) is not available on this machine")
    import dump_golden
    out = str(tmp_path_factory.mktemp("golden") / "self.npz")
    dump_golden.main(["--backend", "oracle", "--reference", ref, "--out", out, "--efc-steps", "5", "--only", "cfg1,scene,wallhit"])
    return np.load(out, allow_pickle=False)


def test_golden_format_selfcheck(self_golden):
    """tools/dump_golden.py --backend oracle -> the consumers above.  Agreement here is the oracle with itself (format check only)."""
    G = self_golden
    assert str(G["backend"][0]) == "oracle" and list(G["scenarios"]) == ["cfg1", "scene", "wallhit"]
    for name in G["scenarios"]:
        T, A = scenario(G, str(name))
        worst, flips = check_oracle_single_steps(T, A, 60)
        assert flips == 0 and worst < 1e-9
        assert check_efc(G, str(name), T, A, 5) < 1e-9
        assert check_oracle_horizon(T, A, 60) < 1e-9
    diff = check_model_constants(scenario(G, "cfg1")[0], "v2")
    assert all((v is True) or (not isinstance(v, bool) and v < 1e-12) for v in diff.values()), diff


@pytest.mark.gpu
def test_golden_format_selfcheck_kernel():
    """Kernel consumer on a self-made file committed as a fixture (tests/golden/self_cfg1.npz, oracle backend): fp64 1e-5, fp32 1e-3."""
    p = os.path.join(ROOT, "tests", "golden", "self_oracle.npz")
    if not os.path.exists(p):
        pytest.skip("tests/golden/self_oracle.npz missing")
    G = np.load(p, allow_pickle=False)
    for name in G["scenarios"]:
        T, A = scenario(G, str(name))
        for dtype, tol in (("float64", 1e-5), ("float32", 1e-3)):
            worst, flips = run_kernel_single_steps(T, A, kind_of(A), dtype, min(200, len(A["act"])))
            print(f"self-check {name} {dtype}: single-step error {worst:.2e}, contact-count flips {flips}")
            assert worst < tol, (name, dtype, worst)
            if dtype == "float64":
                assert flips == 0, (name, flips)
