"""The C-ABI library loads and exports every symbol include/ackb.h declares (no compute calls without a GPU); the
model compiler reproduces the committed tables; multi-rank host logic under gloo."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    lib = ctypes.CDLL(g.build_cuda())
    names = set()
    for h in ("ackb.h", "ackb_ppo.h"):
        hdr = open(os.path.join(ROOT, "include", h)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
        names |= set(re.findall(r"\b(ackb_[a-z_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/*.h but not exported"
    from mujoco_playground_b200 import _lib
    assert set(_lib.SYMBOLS) == names, "ctypes table and header disagree"
    assert lib.ackb_consts_len() > 0
    assert lib.ackb_ppo_num_params(79) == 18757, "parameter count of the reference's MlpPolicy (SURVEY 2.1)"


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mujoco_playground_b200 import BatchedAckermannEnv, _lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchedAckermannEnv(4)
    L = _lib.load()
    h = ctypes.c_void_p()
    blob = np.zeros(L.ackb_consts_len())
    rc = L.ackb_create(blob.ctypes.data_as(ctypes.c_void_p), len(blob), 4, 0, 0, 0, 4, ctypes.byref(h))
    assert rc == -3 and b"no CPU fallback" in L.ackb_last_error(None)


def test_constants_layout_matches_library():
    from mujoco_playground_b200 import _lib
    from mujoco_playground_b200.compiler.constants import build_consts, consts_layout
    from mujoco_playground_b200.models import load_model
    L = _lib.load()
    lay = consts_layout()
    assert sum(c for _, c in lay.values()) == L.ackb_consts_len()
    v2 = build_consts(load_model("v2"), model_kind=0)
    sc = build_consts(load_model("scene"), model_kind=1)

    def f(b, n):
        o, c = lay[n]
        return b[o:o + c]
    assert abs(f(v2, "mass")[0] - 10.3) < 1e-12 and abs(f(sc, "mass")[0] - 10.4) < 1e-12
    assert f(v2, "nbeam")[0] == 72 and f(sc, "nbeam")[0] == 36 and f(sc, "nbox")[0] == 38
    assert f(v2, "has_eq")[0] == 1 and f(sc, "has_eq")[0] == 0
    assert f(v2, "ctrl_kind")[0] == 0 and f(sc, "ctrl_kind")[0] == 1
    np.testing.assert_allclose(f(v2, "w_mu"), 1.4)
    np.testing.assert_allclose(f(sc, "w_mu"), [1.4, 1.4, 1.0, 1.0])
    np.testing.assert_allclose(f(v2, "st_hi"), np.deg2rad(35))
    np.testing.assert_allclose(f(v2, "spawn_qpos")[:7], [0, 0, 0.1, 1, 0, 0, 0])


@pytest.mark.skipif(not os.path.exists("/root/reference/models/ackermann_robot_v2.xml"), reason="reference checkout not present")
def test_committed_tables_match_a_fresh_compile():
    from mujoco_playground_b200.compiler.mjcf import compile_mjcf
    from mujoco_playground_b200.models import load_model
    for name, path in (("v2", "/root/reference/models/ackermann_robot_v2.xml"),
                       ("scene", "/root/reference/models/environments/ackermann_in_mushr_maze.xml")):
        fresh, stored = compile_mjcf(path), load_model(name)
        for k, v in fresh.items():
            if isinstance(v, np.ndarray) and v.dtype.kind == "f":
                np.testing.assert_allclose(stored[k], v, rtol=1e-12, atol=1e-14, err_msg=k)
            elif isinstance(v, np.ndarray):
                np.testing.assert_array_equal(stored[k], v, err_msg=k)
            else:
                assert stored[k] == v, k


def test_mesh_inertia_modes_survey_a3():
    """SURVEY.md Appendix A3 table (computed during the survey from the STL files)."""
    if not os.path.exists("/root/reference/CAD Models/Base.stl"):
        pytest.skip("reference checkout not present")
    from mujoco_playground_b200.compiler.mesh import process_mesh
    R = np.array([[0, 0, -1.0], [-1, 0, 0], [0, 1, 0]])       # geom euler (90, -90, 0)
    for mode, com, I in (("exact", (0.000921, 0.000665, 0.002), (0.008715, 0.029682, 0.038384)),
                         ("convex", (0.006198, 0, 0.002), (0.008833, 0.029175, 0.037994)),
                         ("legacy", (-0.00107, -0.00019, 0.002), (0.008072, 0.028981, 0.037042))):
        me = process_mesh("/root/reference/CAD Models/Base.stl", [1, 1, 1], mode)
        c = R @ me["com"]
        Ib = R @ (me["inertia_unit_density"] * 5.0 / me["volume"]) @ R.T
        np.testing.assert_allclose(c, com, atol=6e-6)   # SURVEY table is rounded to 5-6 decimals
        np.testing.assert_allclose(np.diag(Ib), I, atol=2e-6)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from mujoco_playground_b200.shard import max_over_ranks, reduce_stats, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(1001, rank, world)
    stats = dict(episodes=hi - lo, successes=rank, env_steps=10 * (hi - lo), collisions=1, unsupported=0, solver_iters=3,
                 obstacle_steps=2 * rank, contacts_sum=8 * (hi - lo), return_sum=-1.5 * (rank + 1), length_sum=100.0)
    red = reduce_stats(stats)
    t = max_over_ranks(1.0 + rank)
    dist.destroy_process_group()
    q.put((rank, lo, hi, red, t))


def test_sharding_and_stats_reduction_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    (r0, lo0, hi0, red0, t0), (r1, lo1, hi1, red1, t1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)
    assert red0 == red1 and red0["episodes"] == 1001 and red0["env_steps"] == 10010 and red0["successes"] == 1
    assert red0["obstacle_steps"] == 2 and red0["contacts_sum"] == 8008
    assert abs(red0["return_sum"] + 4.5) < 1e-12 and t0 == t1 == 2.0


def test_factory_signature_matches_the_reference_pyc(capsys):
    """make_ackermann_env / list_available_mazes: argument names, order and defaults recovered from the reference's
    src/rl/__pycache__/make_env.cpython-312.pyc (the source file is missing upstream)."""
    import inspect
    from mujoco_playground_b200 import list_available_mazes, make_ackermann_env
    sig = inspect.signature(make_ackermann_env)
    names = list(sig.parameters)
    assert names[:6] == ["env_type", "maze_id", "render_mode", "max_linear_velocity", "max_angular_velocity", "goal_distance_threshold"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["env_type"], d["maze_id"], d["render_mode"]) == ("maze", "PointMaze_UMaze-v3", None)
    assert (d["max_linear_velocity"], d["max_angular_velocity"], d["goal_distance_threshold"]) == (0.5, 1.0, 0.3)
    assert list_available_mazes() == ["PointMaze_UMaze-v3", "PointMaze-Open-v3", "PointMaze-Medium-v3", "PointMaze-Large-v3"]
    with pytest.raises(ValueError, match="Unknown environment type"):
        make_ackermann_env(env_type="nope")


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs on the host cores (the oracle port) and prints ONE JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0", "--cpu-steps", "8000"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]
