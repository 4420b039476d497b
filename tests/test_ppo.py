"""PPO learner: host-side logic on CPU (GAE, gradient all-reduce under gloo, SB3 parameter layout) and a short CUDA run."""
import os

import numpy as np
import pytest
import torch

from mujoco_playground_b200.ppo import ActorCritic, PPOConfig, allreduce_gradients_, compute_gae, ppo_update


def test_policy_layout_matches_reference_checkpoints():
    """rl_logs/ppo/*.zip -> policy.pth: 18 757 parameters, SB3 MlpPolicy key names."""
    p = ActorCritic(79)
    assert sum(x.numel() for x in p.parameters()) == 18757
    keys = set(p.state_dict())
    for k in ("log_std", "mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.2.bias", "mlp_extractor.value_net.0.weight",
              "mlp_extractor.value_net.2.weight", "action_net.weight", "value_net.bias"):
        assert k in keys
    assert p.state_dict()["action_net.weight"].shape == (2, 64) and p.state_dict()["value_net.weight"].shape == (1, 64)


def test_gae_matches_reference_recursion():
    rng = np.random.default_rng(0)
    T, N = 7, 5
    rew, val = rng.normal(size=(T, N)), rng.normal(size=(T, N))
    done = (rng.uniform(size=(T, N)) < 0.3).astype(np.float64)
    last = rng.normal(size=N)
    adv, ret = compute_gae(*(torch.tensor(x) for x in (rew, val, done)), torch.tensor(last), 0.99, 0.95)
    want = np.zeros((T, N))
    for n in range(N):
        gae = 0.0
        for t in reversed(range(T)):
            nv = last[n] if t == T - 1 else val[t + 1, n]
            nt = 1.0 - done[t, n]
            delta = rew[t, n] + 0.99 * nv * nt - val[t, n]
            gae = delta + 0.99 * 0.95 * nt * gae
            want[t, n] = gae
    np.testing.assert_allclose(adv.numpy(), want, atol=1e-12)
    np.testing.assert_allclose(ret.numpy(), want + val, atol=1e-12)


def test_ppo_update_improves_surrogate_on_cpu():
    torch.manual_seed(0)
    pol = ActorCritic(79)
    opt = torch.optim.Adam(pol.parameters(), lr=3e-3)
    n = 512
    obs = torch.randn(n, 79)
    with torch.no_grad():
        act, logp, val = pol.act(obs)
    adv = act[:, 0].clone()                   # reward pushing action[0] up
    batch = dict(obs=obs, act=act, logp=logp, adv=adv, ret=val + adv)
    before = pol.dist_params(obs)[0][:, 0].mean().item()
    st = ppo_update(pol, opt, batch, PPOConfig(n_epochs=5, minibatches=2))
    after = pol.dist_params(obs)[0][:, 0].mean().item()
    assert after > before and np.isfinite(st["pg_loss"]) and st["steps"] == 10


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    pol = ActorCritic(79)
    torch.manual_seed(100 + rank)
    x = torch.randn(64, 79)
    logp, ent, val = pol.evaluate(x, torch.randn(64, 2))
    loss = val.pow(2).mean() - logp.mean() + ent.mean()
    loss.backward()
    local = torch.cat([p.grad.reshape(-1) for p in pol.parameters()]).clone()
    nbytes = allreduce_gradients_(list(pol.parameters()), world)
    avg = torch.cat([p.grad.reshape(-1) for p in pol.parameters()])
    dist.destroy_process_group()
    q.put((rank, local.numpy(), avg.numpy(), nbytes))


def test_gradient_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 23456 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=60) for _ in ps), key=lambda r: r[0])
    for p in ps:
        p.join(timeout=60)
    (_, l0, a0, nb0), (_, l1, a1, nb1) = res
    np.testing.assert_allclose(a0, (l0 + l1) / 2, rtol=1e-6, atol=1e-9)
    np.testing.assert_array_equal(a0, a1)
    assert nb0 == nb1 == 18757 * 4          # one flattened 75 KB bucket


@pytest.mark.gpu
def test_ppo_short_run_on_cuda():
    from mujoco_playground_b200 import BatchedAckermannEnv
    from mujoco_playground_b200.ppo import PPOTrainer
    env = BatchedAckermannEnv(512, seed=1, max_episode_steps=8)   # every rollout of 8 steps ends all episodes by truncation
    tr = PPOTrainer(env, PPOConfig(n_steps=8, n_epochs=2, minibatches=2), seed=0)
    w0 = tr.policy.action_net.weight.detach().clone()
    out = tr.train(512 * 8 * 8, log=None)
    assert out["timesteps"] >= 512 * 64 and np.isfinite(out["pg_loss"]) and np.isfinite(out["v_loss"])
    assert out["episodes"] > 0 and np.isfinite(out["ep_rew_mean"])
    assert not torch.equal(w0, tr.policy.action_net.weight.detach())
    env.close()


@pytest.mark.skipif(not os.path.exists("/root/reference/rl_logs/ppo/ppo_model_10000_steps.zip"), reason="reference checkout not present")
def test_load_reference_sb3_checkpoints(tmp_path):
    """The three checkpoints the reference commits load into ActorCritic; the recorded _last_obs is the lidar known answer."""
    import base64
    import json
    from mujoco_playground_b200.sb3_io import load_sb3_policy, save_sb3_policy
    for name in ("ppo_model_10000_steps.zip", "ppo_model_20000_steps.zip", "ppo_model_30000_steps.zip"):
        pol, data = load_sb3_policy(f"/root/reference/rl_logs/ppo/{name}")
        assert sum(p.numel() for p in pol.parameters()) == 18757
        assert data.get("n_envs") == 1
        a, logp, v = pol.act(torch.zeros(3, 79))
        assert a.shape == (3, 2) and torch.isfinite(a).all() and torch.isfinite(v).all()
    out = tmp_path / "roundtrip.zip"
    save_sb3_policy(pol, str(out), num_timesteps=123)
    pol2, d2 = load_sb3_policy(str(out))
    assert d2["num_timesteps"] == 123
    for k, v in pol.state_dict().items():
        assert torch.equal(v, pol2.state_dict()[k])


def _sb3_style_load(path, mods):
    """What stable_baselines3.common.save_util.load_from_zip_file + json_to_data do (SB3 2.7.0), with the class references resolved
    through the modules in `mods` (the real ones when installed, stand-ins here)."""
    import base64
    import io
    import json
    import pickle
    import zipfile
    with zipfile.ZipFile(path) as z:
        names = z.namelist()
        data = json.loads(z.read("data").decode())
        out = {}
        for k, v in data.items():
            out[k] = pickle.loads(base64.b64decode(v[":serialized:"])) if isinstance(v, dict) and ":serialized:" in v else v
        params = {n[:-4]: torch.load(io.BytesIO(z.read(n)), map_location="cpu", weights_only=True) for n in names if n.endswith(".pth") and n != "pytorch_variables.pth"}
        pv = torch.load(io.BytesIO(z.read("pytorch_variables.pth")), map_location="cpu", weights_only=True)
        ver = z.read("_stable_baselines3_version").decode()
    return out, params, pv, ver


def test_saved_archive_has_the_layout_ppo_load_opens(tmp_path):
    """save_sb3_policy writes what PPO.load reads: every pickled field of `data` resolves to the SB3 / gymnasium
    class PPO.load expects and carries the state SB3 itself writes; policy.pth / policy.optimizer.pth have SB3's parameter order.
    SB3 cannot be installed here, so the class references are resolved through stand-in modules; where the reference checkout is
    present the pickle streams are also compared opcode by opcode with the reference's own archive."""
    import base64
    import json
    import pickletools
    import zipfile
    from mujoco_playground_b200.ppo import ActorCritic
    from mujoco_playground_b200.sb3_io import SB3_PARAM_ORDER, load_sb3_policy, save_sb3_policy, sb3_stub_modules
    torch.manual_seed(0)
    pol = ActorCritic(79)
    opt = torch.optim.Adam(pol.parameters(), lr=3e-4, eps=1e-5)
    a_, lp_, v_ = pol.act(torch.randn(8, 79))
    (lp_.sum() + v_.sum()).backward()
    opt.step()
    out = str(tmp_path / "ppo_final.zip")
    save_sb3_policy(pol, out, opt, num_timesteps=4096, total_timesteps=8192, hyper=dict(n_steps=16, batch_size=1024), last_obs=np.ones((4, 79), np.float32))
    with sb3_stub_modules() as mods:
        data, params, pv, ver = _sb3_style_load(out, mods)
        Box = mods["gymnasium.spaces.box"].Box
        U = mods["stable_baselines3.common.utils"]
        assert data["policy_class"] is mods["stable_baselines3.common.policies"].ActorCriticPolicy
        assert data["rollout_buffer_class"] is mods["stable_baselines3.common.buffers"].RolloutBuffer
        osp, asp = data["observation_space"], data["action_space"]
        assert isinstance(osp, Box) and osp._shape == (79,) and osp.dtype == np.float32 and np.all(np.isneginf(osp.low)) and not osp.bounded_above.any()
        assert isinstance(asp, Box) and asp._shape == (2,) and np.array_equal(asp.low, [-1, -1]) and np.array_equal(asp.high, [1, 1]) and asp.bounded_below.all()
        assert isinstance(data["clip_range"], U.FloatSchedule) and data["clip_range"].value_schedule.val == 0.2
        assert isinstance(data["lr_schedule"], U.FloatSchedule) and data["lr_schedule"].value_schedule.val == 3e-4
    assert data["_last_obs"].shape == (4, 79) and data["n_envs"] == 4 and data["num_timesteps"] == 4096 and data["n_steps"] == 16
    assert ver == "2.7.0" and pv == {}
    assert tuple(params["policy"].keys()) == SB3_PARAM_ORDER
    osd = params["policy.optimizer"]
    assert osd["param_groups"][0]["params"] == list(range(13)) and len(osd["state"]) == 13
    named = dict(pol.named_parameters())
    for i, name in enumerate(SB3_PARAM_ORDER):
        assert osd["state"][i]["exp_avg"].shape == named[name].shape
    pol2, _ = load_sb3_policy(out)
    for k, v in pol.state_dict().items():
        assert torch.equal(v, pol2.state_dict()[k])
    ref = "/root/reference/rl_logs/ppo/ppo_model_10000_steps.zip"
    if os.path.exists(ref):      # same pickle opcodes as SB3's own writer for the fields that name third-party classes
        with zipfile.ZipFile(ref) as z:
            rd = json.loads(z.read("data"))
        with zipfile.ZipFile(out) as z:
            md = json.loads(z.read("data"))
        assert set(rd.keys()) == set(md.keys()), set(rd.keys()) ^ set(md.keys())
        ops = lambda b: [(o.name, a) for o, a, _ in pickletools.genops(base64.b64decode(b)) if o.name not in ("FRAME", "MEMOIZE", "BINGET", "BINFLOAT")]
        for k in ("policy_class", "rollout_buffer_class", "observation_space", "action_space", "clip_range", "lr_schedule"):
            assert ops(rd[k][":serialized:"]) == ops(md[k][":serialized:"]) or k == "lr_schedule", k
            assert rd[k][":type:"] == md[k][":type:"], k


@pytest.mark.gpu
def test_cuda_graph_update_matches_eager_update():
    """GraphedMinibatchStep (forward/backward and clip+Adam captured in CUDA graphs) performs the same optimiser steps as the
    eager loop of ppo_update on the same batch and the same minibatch permutation."""
    import copy
    from mujoco_playground_b200.ppo import ActorCritic, GraphedMinibatchStep, ppo_update
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    n, D = 4096, 79
    batch = dict(obs=torch.randn(n, D, device=dev), act=torch.randn(n, 2, device=dev).clamp(-1, 1), logp=torch.randn(n, device=dev) * 0.1 - 2.0,
                 adv=torch.randn(n, device=dev), ret=torch.randn(n, device=dev))
    cfg = PPOConfig(n_steps=1, n_epochs=3, minibatches=4)
    pol_a = ActorCritic(D).to(dev)
    pol_b = copy.deepcopy(pol_a)
    opt_a = torch.optim.Adam(pol_a.parameters(), lr=cfg.learning_rate, eps=cfg.adam_eps)
    opt_b = torch.optim.Adam(pol_b.parameters(), lr=cfg.learning_rate, eps=cfg.adam_eps, capturable=True)
    g = GraphedMinibatchStep(pol_b, opt_b, cfg, n // cfg.minibatches, D, dev)
    for pa, pb in zip(pol_a.parameters(), pol_b.parameters()):
        assert torch.equal(pa, pb), "capture must leave the weights untouched"
    ga, gb = torch.Generator(device=dev), torch.Generator(device=dev)
    ga.manual_seed(11); gb.manual_seed(11)
    sa = ppo_update(pol_a, opt_a, batch, cfg, generator=ga)
    sb = ppo_update(pol_b, opt_b, batch, cfg, generator=gb, graphed=g)
    assert sa["steps"] == sb["steps"] == 12
    for k in ("pg_loss", "v_loss", "entropy", "approx_kl"):
        assert abs(sa[k] - sb[k]) < 1e-4 * max(1.0, abs(sa[k])), k
    for pa, pb in zip(pol_a.parameters(), pol_b.parameters()):
        assert torch.allclose(pa, pb, atol=2e-5, rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("obs_dim,n,tc,tol", [(79, 4096, 0, 2e-4), (43, 1000, 0, 2e-4), (79, 4096, 1, 4e-3), (43, 1000, 1, 4e-3),
                                              (79, 4096, 2, 4e-3), (43, 1000, 2, 4e-3), (79, 131, 2, 4e-3), (79, 40000, 1, 1e-2), (79, 40000, 2, 1e-2)])
def test_fused_gradient_kernel_matches_autograd(obs_dim, n, tc, tol):
    """csrc/ackb_ppo.cu against torch autograd on the same minibatch: every parameter gradient and the loss diagnostics
    (ragged last tile, index gather, both observation widths)."""
    import copy
    from mujoco_playground_b200.ppo import ActorCritic, FusedMinibatchStep
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    cfg = PPOConfig(n_steps=1, n_epochs=1, minibatches=1)
    N = 3 * n
    batch = dict(obs=torch.randn(N, obs_dim, device=dev), act=torch.randn(N, 2, device=dev).clamp(-1, 1), logp=torch.randn(N, device=dev) * 0.3 - 2.0,
                 adv=torch.randn(N, device=dev) * 3 + 1, ret=torch.randn(N, device=dev) * 5)
    idx = torch.randperm(N, device=dev)[:n]
    pol = ActorCritic(obs_dim).to(dev)
    with torch.no_grad():
        pol.log_std.copy_(torch.tensor([-0.3, 0.2], device=dev))
        pol.action_net.weight.mul_(30.0)           # away from the tiny initial gain so that ratios leave the clip range
    ref = copy.deepcopy(pol)
    # torch reference
    obs, act, old_logp, adv, ret = (batch[k][idx] for k in ("obs", "act", "logp", "adv", "ret"))
    advn = (adv - adv.mean()) / (adv.std() + 1e-8)
    logp, ent, val = ref.evaluate(obs, act)
    ratio = torch.exp(logp - old_logp)
    pg = -torch.min(advn * ratio, advn * torch.clamp(ratio, 1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
    vl = torch.nn.functional.mse_loss(val, ret)
    (pg + cfg.vf_coef * vl - cfg.ent_coef * ent.mean()).backward()
    clip_frac = ((ratio - 1).abs() > cfg.clip_range).float().mean().item()
    assert 0.05 < clip_frac < 0.95, "the sample must exercise both branches of the clipped surrogate"
    # fused kernel (lr = 0: the optimiser step inside run() must not move the weights)
    opt = torch.optim.SGD(pol.parameters(), lr=0.0)
    f = FusedMinibatchStep(pol, opt, cfg, obs_dim, dev)
    f.mode = tc                      # 0: fp32 CUDA cores, 1: TF32 mma.sync, 2: TF32 tcgen05 / TMEM (fp32 accumulation)
    f.cfg = PPOConfig(max_grad_norm=1e30)          # no clipping: compare raw gradients
    f.run(batch, idx, world=1)
    torch.cuda.synchronize()
    for (name, pr), pf in zip(ref.named_parameters(), pol.parameters()):
        scale = max(1e-6, pr.grad.abs().max().item())
        assert (pr.grad - pf.grad).abs().max().item() < tol * scale + 1e-7, name
    d = f.diag.cpu().numpy()
    dt = 1e-4 if tc == 0 else 5e-3
    assert abs(d[0] - pg.item()) < dt * max(1, abs(pg.item())) and abs(d[1] - vl.item()) < dt * vl.item()
    assert abs(d[2] - ent.mean().item()) < 1e-5 and abs(d[4] - clip_frac) < (1e-6 if tc == 0 else 5e-3)
    assert abs(d[3] - ((ratio - 1) - (logp - old_logp)).mean().item()) < 10 * dt * max(1.0, abs(d[3]))
    f.mode = 1


@pytest.mark.gpu
def test_fused_learner_update_matches_eager_update():
    """Whole ppo_update (3 epochs x 4 minibatches, clipping + Adam) with the fused gradient kernel vs the eager torch loop."""
    import copy
    from mujoco_playground_b200.ppo import ActorCritic, FusedMinibatchStep, ppo_update
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    n, D = 4096, 79
    batch = dict(obs=torch.randn(n, D, device=dev), act=torch.randn(n, 2, device=dev).clamp(-1, 1), logp=torch.randn(n, device=dev) * 0.1 - 2.0,
                 adv=torch.randn(n, device=dev), ret=torch.randn(n, device=dev))
    cfg = PPOConfig(n_steps=1, n_epochs=3, minibatches=4)
    pol_a = ActorCritic(D).to(dev)
    pol_b = copy.deepcopy(pol_a)
    opt_a = torch.optim.Adam(pol_a.parameters(), lr=cfg.learning_rate, eps=cfg.adam_eps)
    opt_b = torch.optim.Adam(pol_b.parameters(), lr=cfg.learning_rate, eps=cfg.adam_eps)
    f = FusedMinibatchStep(pol_b, opt_b, cfg, D, dev)
    f.mode = 0                       # fp32 CUDA-core arithmetic for the tight comparison with the eager loop
    # third copy: the same fused step replayed from CUDA graphs (capturable Adam)
    pol_c = copy.deepcopy(pol_a)
    opt_c = torch.optim.Adam(pol_c.parameters(), lr=cfg.learning_rate, eps=cfg.adam_eps, capturable=True)
    fc = FusedMinibatchStep(pol_c, opt_c, cfg, D, dev, mode="fp32")
    fc.capture(fc.shuffle_epoch(batch, torch.arange(n, device=dev)), [(i * (n // 4), n // 4) for i in range(4)])
    for pa, pc in zip(pol_a.parameters(), pol_c.parameters()):
        assert torch.equal(pa, pc), "capture must leave the weights untouched"
    gens = [torch.Generator(device=dev) for _ in range(3)]
    for g in gens:
        g.manual_seed(11)
    sa = ppo_update(pol_a, opt_a, batch, cfg, generator=gens[0])
    sb = ppo_update(pol_b, opt_b, batch, cfg, generator=gens[1], graphed=f)
    sc = ppo_update(pol_c, opt_c, batch, cfg, generator=gens[2], graphed=fc)
    assert sa["steps"] == sb["steps"] == sc["steps"] == 12
    for k in ("pg_loss", "v_loss", "entropy", "approx_kl"):
        assert abs(sa[k] - sb[k]) < 1e-4 * max(1.0, abs(sa[k])), k
        assert abs(sa[k] - sc[k]) < 1e-4 * max(1.0, abs(sa[k])), k
    for pa, pb, pc in zip(pol_a.parameters(), pol_b.parameters(), pol_c.parameters()):
        assert torch.allclose(pa, pb, atol=5e-5, rtol=1e-3)
        assert torch.allclose(pa, pc, atol=5e-5, rtol=1e-3)
    f.mode = 1


@pytest.mark.gpu
@pytest.mark.parametrize("obs_dim,n,pitch", [(79, 5000, 0), (43, 777, 0), (79, 5000, 80), (79, 131, 80), (79, 70000, 80), (43, 777, 44)])
def test_fused_act_kernel_matches_policy_forward(obs_dim, n, pitch):
    """Rollout forward (ppo_act_kernel on mma.sync TF32 tiles for dense rows; ppo_act_kernel_tcgen05 for rows pitched to a multiple of
    4 floats, what PPOTrainer feeds it): mean and value against the torch policy, the sampled action is mean + exp(log_std) * eps
    with eps ~ N(0, 1) (moments checked), its log-probability matches the Gaussian density, the noise stream is reproducible per
    (seed, step) and differs between steps; value-only mode gives the same values; both kernels draw the same noise."""
    import ctypes
    from mujoco_playground_b200.ppo import ActorCritic, FusedMinibatchStep
    dev = torch.device("cuda:0")
    torch.manual_seed(9)
    pol = ActorCritic(obs_dim).to(dev)
    with torch.no_grad():
        pol.log_std.copy_(torch.tensor([-0.4, 0.3], device=dev))
        pol.action_net.weight.mul_(20.0)
    dense = torch.randn(n, obs_dim, device=dev) * 2
    with torch.no_grad():
        mean_t, _ = pol.dist_params(dense)
        val_t = pol.value(dense)
    if pitch:
        buf = torch.full((n, pitch), 55.0, device=dev)     # padding columns hold junk: never used
        buf[:, :obs_dim] = dense
        obs = buf[:, :obs_dim]                             # row stride = pitch
    else:
        obs = dense
    f = FusedMinibatchStep(pol, torch.optim.SGD(pol.parameters(), lr=0.0), PPOConfig(), obs_dim, dev)
    act, logp, val = torch.empty(n, 2, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev)
    f.act(obs, act, logp, val, seed=77, step=3)
    mean = torch.empty(n, 2, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert f.L.ackb_ppo_act_pitched(p(obs), int(obs.stride(0)), n, obs_dim, p(f.flat_p), p(mean), p(val), None, None, 0, 0, 0, None) == 0
    torch.cuda.synchronize()
    if pitch:      # same random stream as the dense-row kernel
        act_d, logp_d, val_d = torch.empty_like(act), torch.empty_like(logp), torch.empty_like(val)
        f.act(dense, act_d, logp_d, val_d, seed=77, step=3)
        torch.cuda.synchronize()
        assert (act - act_d).abs().max().item() < 1e-2 * max(1.0, mean_t.abs().max().item()) and (logp - logp_d).abs().max().item() < 1e-5
    assert (mean - mean_t).abs().max().item() < 5e-3 * max(1.0, mean_t.abs().max().item())
    assert (val - val_t).abs().max().item() < 5e-3 * max(1.0, val_t.abs().max().item())
    eps = (act - mean) / torch.exp(pol.log_std.detach())
    if n >= 700:        # sample moments need samples
        assert abs(eps.mean().item()) < 0.06 and abs(eps.var().item() - 1.0) < 0.08
        assert abs((eps[:, 0] * eps[:, 1]).mean().item()) < 0.06, "the two action noises are independent"
    want_logp = ActorCritic.log_prob(mean, pol.log_std.detach(), act)
    assert (logp - want_logp).abs().max().item() < 1e-3
    act2, logp2, val2 = torch.empty_like(act), torch.empty_like(logp), torch.empty_like(val)
    f.act(obs, act2, logp2, val2, seed=77, step=3)
    assert torch.equal(act, act2) and torch.equal(logp, logp2)
    f.act(obs, act2, logp2, val2, seed=77, step=4)
    assert not torch.equal(act, act2)
    v_only = torch.empty(n, device=dev)
    f.value(obs, v_only)
    torch.cuda.synchronize()
    assert torch.allclose(v_only, val, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("T,n", [(16, 1000), (1, 7), (5, 4097)])
def test_fused_gae_kernel_matches_compute_gae(T, n):
    """ackb_ppo_gae against the torch loop (SB3 compute_returns_and_advantage) incl. episode ends inside the rollout."""
    import ctypes
    from mujoco_playground_b200 import _lib
    L, dev = _lib.load(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(T * 131 + n)
    rew, val = torch.randn(T, n, generator=g).to(dev), torch.randn(T, n, generator=g).to(dev)
    done = (torch.rand(T, n, generator=g) < 0.2).float().to(dev)
    last = torch.randn(n, generator=g).to(dev)
    adv, ret = torch.empty(T, n, device=dev), torch.empty(T, n, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert L.ackb_ppo_gae(p(rew), p(val), p(done), p(last), T, n, 0.99, 0.95, p(adv), p(ret), None) == 0
    torch.cuda.synchronize()
    want_adv, want_ret = compute_gae(rew.double(), val.double(), done.double(), last.double(), 0.99, 0.95)
    assert (adv.double() - want_adv).abs().max().item() < 1e-5
    assert (ret.double() - want_ret).abs().max().item() < 1e-5
    assert L.ackb_ppo_gae(None, p(val), p(done), p(last), T, n, 0.99, 0.95, p(adv), p(ret), None) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["index", "copy"])
def test_fused_update_minibatch_modes_agree_with_eager(mode, monkeypatch):
    """Index mode (kernel gathers rows through the epoch's permutation) and copy mode (permuted copy, contiguous minibatches)
    run the same optimiser steps as the eager torch loop for the same permutations."""
    from mujoco_playground_b200.ppo import FusedMinibatchStep
    monkeypatch.setenv("ACKB_PPO_SHUFFLE", mode)
    dev, obs_dim, n = torch.device("cuda:0"), 79, 2048
    cfg = PPOConfig(n_epochs=2, minibatches=4)
    g = torch.Generator().manual_seed(5)
    batch = dict(obs=torch.randn(n, obs_dim, generator=g), act=torch.randn(n, 2, generator=g) * 0.5, logp=-torch.rand(n, generator=g) - 1.0,
                 adv=torch.randn(n, generator=g), ret=torch.randn(n, generator=g))
    batch = {k: v.to(dev) for k, v in batch.items()}
    torch.manual_seed(3)
    pa = ActorCritic(obs_dim).to(dev)
    pb = ActorCritic(obs_dim).to(dev)
    pb.load_state_dict(pa.state_dict())
    oa, ob = torch.optim.Adam(pa.parameters(), lr=3e-4, eps=1e-5), torch.optim.Adam(pb.parameters(), lr=3e-4, eps=1e-5)
    fused = FusedMinibatchStep(pb, ob, cfg, obs_dim, dev)
    assert fused.index_mode == (mode == "index")
    fused.mode = 0                   # fp32 arithmetic for the tight comparison
    try:
        ppo_update(pa, oa, batch, cfg, generator=torch.Generator(device=dev).manual_seed(9))
        ppo_update(pb, ob, batch, cfg, generator=torch.Generator(device=dev).manual_seed(9), graphed=fused)
    finally:
        fused.mode = 1
    for (name, x), y in zip(pa.state_dict().items(), pb.state_dict().values()):
        assert torch.allclose(x, y, atol=5e-5, rtol=1e-3), name


@pytest.mark.gpu
def test_native_adv_stats_and_clip_adam_match_torch():
    """ackb_ppo_adv_stats vs mean()/std(), ackb_ppo_clip_adam vs clip_grad_norm_ + torch.optim.Adam over several steps
    (clipping active and inactive)."""
    import ctypes
    from mujoco_playground_b200 import _lib
    L, dev = _lib.load(), torch.device("cuda:0")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    g = torch.Generator().manual_seed(21)
    adv = (torch.randn(100000, generator=g) * 3 + 0.7).to(dev)
    idx = torch.randperm(100000, generator=g)[:4097].to(dev)
    out = torch.zeros(2, device=dev)
    assert L.ackb_ppo_adv_stats(p(adv), p(idx), 4097, p(out), None) == 0
    sel = adv[idx]
    assert torch.allclose(out, torch.stack([sel.mean(), sel.std()]), rtol=1e-5, atol=1e-6)
    assert L.ackb_ppo_adv_stats(p(adv), None, 1000, p(out), None) == 0
    assert torch.allclose(out, torch.stack([adv[:1000].mean(), adv[:1000].std()]), rtol=1e-5, atol=1e-6)

    n = 18757
    w0 = torch.randn(n, generator=g).to(dev)
    ref = torch.nn.Parameter(w0.clone())
    opt = torch.optim.Adam([ref], lr=3e-4, eps=1e-5)
    w, m, v, step = w0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros((), device=dev)
    for it in range(6):
        grad = (torch.randn(n, generator=g) * (0.001 if it % 2 else 0.05)).to(dev)      # norm 0.14 (no clip) / 6.8 (clipped to 0.5)
        ref.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([ref], 0.5)
        opt.step()
        assert L.ackb_ppo_clip_adam(p(w), p(grad), p(m), p(v), p(step), n, 0.5, 3e-4, 0.9, 0.999, 1e-5, None) == 0
    torch.cuda.synchronize()
    assert step.item() == 6.0
    assert torch.allclose(w, ref.detach(), atol=1e-6, rtol=1e-5)
    st = opt.state[ref]
    assert torch.allclose(m, st["exp_avg"], atol=1e-8, rtol=1e-4) and torch.allclose(v, st["exp_avg_sq"], atol=1e-12, rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("n,p_trunc", [(5000, 0.01), (33, 0.5), (4096, 0.0)])
def test_bootstrap_kernel_matches_torch(n, p_trunc):
    """ackb_ppo_bootstrap: reward + gamma V(terminal obs) only where truncated and not terminated; done = term | trunc.  Tiles
    without a truncated row take the shortcut that skips the value network."""
    from mujoco_playground_b200.ppo import ActorCritic, FusedMinibatchStep
    dev, D = torch.device("cuda:0"), 79
    torch.manual_seed(4)
    pol = ActorCritic(D).to(dev)
    f = FusedMinibatchStep(pol, torch.optim.SGD(pol.parameters(), lr=0.0), PPOConfig(), D, dev)
    g = torch.Generator().manual_seed(n)
    tobs = (torch.randn(n, D, generator=g) * 2).to(dev)
    term = (torch.rand(n, generator=g) < 0.05).to(torch.uint8).to(dev)
    trunc = (torch.rand(n, generator=g) < p_trunc).to(torch.uint8).to(dev)
    rew = torch.randn(n, generator=g).to(dev)
    rew_out, done_out = torch.full((n,), 7.0, device=dev), torch.full((n,), 7.0, device=dev)
    f.bootstrap(tobs, term, trunc, rew, rew_out, done_out)
    torch.cuda.synchronize()
    with torch.no_grad():
        v = pol.value(tobs)
    only = ((trunc != 0) & (term == 0)).float()
    want = rew + 0.99 * v * only
    assert torch.equal(done_out, ((term != 0) | (trunc != 0)).float())
    assert torch.equal(rew_out[only == 0], rew[only == 0]), "rows without a time-limit truncation keep their reward bit for bit"
    assert (rew_out - want).abs().max().item() < 5e-3 * max(1.0, v.abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 5, 1000, 4096, 65537, 1048576])
def test_native_permutation_is_a_permutation(n):
    """ackb_ppo_permutation: a bijection of 0..n-1 for any n, reproducible per (seed, stream), different across streams, and
    not close to the identity (mean displacement ~ n/3 for a random permutation)."""
    import ctypes
    from mujoco_playground_b200 import _lib
    L, dev = _lib.load(), torch.device("cuda:0")
    a, b, c = (torch.full((n,), -1, dtype=torch.int64, device=dev) for _ in range(3))
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert L.ackb_ppo_permutation(p(a), n, 1234567, 1, None) == 0
    assert L.ackb_ppo_permutation(p(b), n, 1234567, 1, None) == 0
    assert L.ackb_ppo_permutation(p(c), n, 1234567, 2, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(torch.sort(a).values, torch.arange(n, device=dev))
    assert torch.equal(torch.sort(c).values, torch.arange(n, device=dev))
    assert torch.equal(a, b)
    if n >= 1000:
        assert not torch.equal(a, c)
        disp = (a - torch.arange(n, device=dev)).abs().double().mean().item() / n
        assert 0.30 < disp < 0.37, disp
        # minibatch quarters draw evenly from the whole range
        q = a[: n // 4].double().mean().item() / n
        assert 0.45 < q < 0.55, q
    assert L.ackb_ppo_permutation(None, n, 0, 0, None) != 0


@pytest.mark.gpu
def test_rollout_direct_observation_writes_match_copy_path():
    """collect() with the step kernel writing observations straight into the rollout buffer (obs_out) fills the same buffers,
    bit for bit, as the path that copies the environment's observation tensor."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    from mujoco_playground_b200.ppo import PPOTrainer
    bufs = []
    for direct in (True, False):
        env = BatchedAckermannEnv(300, seed=5, max_episode_steps=6)
        tr = PPOTrainer(env, PPOConfig(n_steps=8, n_epochs=1, minibatches=2), seed=2)
        assert tr._env_takes_obs_out
        tr._env_takes_obs_out = direct
        tr.collect()
        tr.collect()
        bufs.append({k: v.clone() for k, v in tr.buf.items()} | {"last_obs": tr.obs.clone()})
        env.close()
    for k in bufs[0]:
        assert torch.equal(bufs[0][k], bufs[1][k]), k
    with pytest.raises(ValueError):
        env = BatchedAckermannEnv(4, seed=1)
        try:
            env.step(None, obs_out=torch.zeros(3, env.obs_dim, device=env.device))
        finally:
            env.close()


@pytest.mark.gpu
def test_pitched_observation_rows():
    """Rows padded to 80 floats (ackb_set_obs_pitch + the *_pitched learner entry points): the env writes the same observations into
    a padded array, terminal observations follow the pitch, and the tcgen05 gradient kernel (16-byte vector gather) returns the same
    gradients as on dense rows."""
    from mujoco_playground_b200 import BatchedAckermannEnv
    from mujoco_playground_b200.ppo import ActorCritic, FusedMinibatchStep
    dev = torch.device("cuda:0")
    n = 300
    a, b = BatchedAckermannEnv(n, seed=4, max_episode_steps=6), BatchedAckermannEnv(n, seed=4, max_episode_steps=6)
    pad = torch.full((n, 80), 7.0, device=dev)
    oa = a.reset()
    ob = b.reset(obs_out=pad)
    assert torch.equal(oa, ob[:, :79]) and (ob[:, 79] == 7.0).all()
    for t in range(6):                      # the sixth step ends every episode (time limit)
        oa, ra, ta, tra, ia = a.step(None)
        ob, rb, tb, trb, ib = b.step(None, obs_out=pad)
        assert torch.equal(oa, ob[:, :79]) and torch.equal(ra, rb) and (ob[:, 79] == 7.0).all()
    assert trb.any() and ib["terminal_observation"].shape == (n, 80)
    done = trb.bool() | tb.bool()
    assert torch.equal(ia["terminal_observation"][done], ib["terminal_observation"][done][:, :79])
    oa2 = a.step(None)[0]                   # back to the dense buffer after pitched calls on the other handle
    ob2 = b.step(None)[0]
    assert torch.equal(oa2, ob2) and ob2.shape == (n, 79)
    a.close(); b.close()
    # learner: dense vs padded rollout rows
    torch.manual_seed(5)
    m, D = 3000, 79
    obs = torch.randn(m, D, device=dev)
    obs_p = torch.zeros(m, 80, device=dev)
    obs_p[:, :D] = obs
    obs_p[:, 79] = 123.0                     # never read
    rest = dict(act=torch.randn(m, 2, device=dev).clamp(-1, 1), logp=torch.randn(m, device=dev) * 0.3 - 2.0, adv=torch.randn(m, device=dev),
                ret=torch.randn(m, device=dev))
    cfg = PPOConfig(max_grad_norm=1e30)
    pol = ActorCritic(D).to(dev)
    f = FusedMinibatchStep(pol, torch.optim.SGD(pol.parameters(), lr=0.0), cfg, D, dev, mode="tcgen05")
    idx = torch.randperm(m, device=dev)[: m - 5]
    out = []
    for o in (obs, obs_p):
        f.run(dict(obs=o, **rest), idx, world=1)
        torch.cuda.synchronize()
        out.append((f.flat_g.clone(), f.diag.clone()))
    sc = out[0][0].abs().max().item()
    assert (out[0][0] - out[1][0]).abs().max().item() < 1e-5 * sc, "same TF32 operands, different gather path: only the summation order of atomics differs"
    assert torch.allclose(out[0][1], out[1][1], rtol=1e-5, atol=1e-6)
    for mode in ("tf32", "fp32"):            # the older kernels take the pitch too
        f.mode = f.MODES[mode]
        f.run(dict(obs=obs_p, **rest), idx, world=1)
        torch.cuda.synchronize()
        assert (f.flat_g - out[0][0]).abs().max().item() < 1e-2 * sc


@pytest.mark.gpu
def test_peer_memory_allreduce_matches_nccl_on_two_gpus():
    """ackb_ppo_clip_adam_allreduce (gradient all-reduce over NVLink peer memory inside the optimiser-step kernel) against the NCCL
    path: same parameters after four PPO iterations, bit-identical on every rank.  Needs two GPUs (tools/gpu/ppo_peer_check.py under
    torchrun); skipped on a single-GPU box."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "gpu", "ppo_peer_check.py"), "4096"],
                       capture_output=True, text=True, timeout=600)
    assert "PEER CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["tcgen05", "tf32", "fp32"])
def test_diag_accumulate_flag_sums_the_minibatch_diagnostics(mode):
    """ACKB_PPO_DIAG_ACCUMULATE: the gradient call adds its five diagnostics to `diag` instead of overwriting it (ppo_update then
    launches nothing per optimiser step to accumulate them); the gradient itself is unchanged."""
    from mujoco_playground_b200.ppo import FusedMinibatchStep
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    m, D = 4096, 79
    obs = torch.zeros(m, 80, device=dev)
    obs[:, :D] = torch.randn(m, D, device=dev)
    batch = dict(obs=obs, act=torch.randn(m, 2, device=dev).clamp(-1, 1), logp=torch.randn(m, device=dev) * 0.3 - 2.0,
                 adv=torch.randn(m, device=dev), ret=torch.randn(m, device=dev))
    pol = ActorCritic(D).to(dev)
    f = FusedMinibatchStep(pol, torch.optim.SGD(pol.parameters(), lr=0.0), PPOConfig(max_grad_norm=1e30), D, dev, mode=mode)
    parts = [torch.randperm(m, device=dev)[:1500 + 300 * i] for i in range(3)]
    single = []
    for ix in parts:
        f.run(batch, ix, world=1)
        torch.cuda.synchronize()
        single.append((f.diag.clone(), f.flat_g.clone()))
    f.accumulate_diag = True
    f.diag.zero_()
    for i, ix in enumerate(parts):
        f.run(batch, ix, world=1)
        torch.cuda.synchronize()
        sc = single[i][1].abs().max().item()
        assert (f.flat_g - single[i][1]).abs().max().item() < 1e-5 * sc
    want = sum(d for d, _ in single)
    assert torch.allclose(f.diag, want, rtol=1e-5, atol=1e-6), (f.diag, want)


def test_ppo_update_host_logic_with_in_kernel_diagnostics():
    """ppo_update around a FusedMinibatchStep whose gradient calls accumulate the diagnostics themselves (accumulate_diag): the
    accumulator is zeroed once per update, every epoch goes through run_epoch with the (start, count) ranges of its minibatches, no
    per-step accumulation happens on the host side, and the reported losses are the accumulated sums divided by the number of
    optimiser steps.  The learner is a stand-in without CUDA: this is the host logic only (the kernels are covered by the gpu tests)."""
    from mujoco_playground_b200.ppo import FusedMinibatchStep
    import ctypes

    def header_flag():
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for line in open(os.path.join(root, "include", "ackb_ppo.h")):
            if line.strip().startswith("ACKB_PPO_DIAG_ACCUMULATE"):
                return int(line.split("=")[1].strip().rstrip(",").split()[0], 0)
        raise AssertionError("ACKB_PPO_DIAG_ACCUMULATE not declared in include/ackb_ppo.h")
    assert FusedMinibatchStep.DIAG_ACCUMULATE == header_flag()

    class Stub(FusedMinibatchStep):
        def __init__(self):      # no library, no device
            self.accumulate_diag, self.index_mode, self.mode = True, True, 2
            self.diag = torch.full((5,), 123.0)      # stale values of an earlier update: must be zeroed
            self.calls, self.epochs = [], []
            self.ct = ctypes

        def new_epoch(self, batch, seed, epoch):
            self.epochs.append((seed, epoch))
            return batch

        def run_epoch(self, batch, slots, world):
            self.calls.append(tuple(slots))
            for _ in slots:
                self.diag += torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0])
            return 10 * len(slots)

        def run(self, *a, **k):
            raise AssertionError("per-step path taken although the diagnostics accumulate in the kernel")

    n, cfg = 1000, PPOConfig(n_epochs=3, minibatches=4)
    batch = dict(obs=torch.zeros(n, 79), act=torch.zeros(n, 2), logp=torch.zeros(n), adv=torch.zeros(n), ret=torch.zeros(n))
    pol = ActorCritic(79)
    f = Stub()
    st = ppo_update(pol, torch.optim.SGD(pol.parameters(), lr=0.0), batch, cfg, world=1, graphed=f, perm_seed=77)
    assert f.calls == [((0, 250), (250, 250), (500, 250), (750, 250))] * 3
    assert [s for s, _ in f.epochs] == [77] * 3 and len({e for _, e in f.epochs}) == 3      # a fresh permutation per epoch
    assert st["steps"] == 12 and st["allreduce_bytes"] == 120
    assert (st["pg_loss"], st["v_loss"], st["entropy"], st["approx_kl"], st["clip_frac"]) == (1.0, 2.0, 3.0, 4.0, 5.0)
    # the flag travels on the per-call mode; "default" (-1) is resolved to a concrete arithmetic first
    assert f._call_mode() == (2 | FusedMinibatchStep.DIAG_ACCUMULATE)
    f.mode = -1
    assert f._call_mode() in (0 | FusedMinibatchStep.DIAG_ACCUMULATE, 1 | FusedMinibatchStep.DIAG_ACCUMULATE)
    f.accumulate_diag = False
    assert f._call_mode() == -1


def test_run_epoch_graph_selection_logic():
    """FusedMinibatchStep.run_epoch replays the epoch graph only for exactly the captured slots and arrays, on a single rank or with the
    peer-memory all-reduce, and (peer path) only when the gradient-buffer parity is back at zero; otherwise it steps one by one."""
    from mujoco_playground_b200.ppo import FusedMinibatchStep

    class FakeGraph:
        def __init__(self):
            self.n = 0

        def replay(self):
            self.n += 1

    class Stub(FusedMinibatchStep):
        def __init__(self):
            self.g_epoch, self._epoch_slots, self._mbc, self.peer = FakeGraph(), ((0, 4), (4, 4)), 0, None
            self.flat_g = torch.zeros(10)
            self.same, self.steps = True, []

        def _same_arrays(self, batch):
            return self.same

        def run(self, batch, idx, world):
            self.steps.append(idx)
            self._mbc += 1
            return 7

    f, slots = Stub(), [(0, 4), (4, 4)]
    assert f.run_epoch({}, slots, 1) == 0 and f.g_epoch.n == 1 and f._mbc == 2 and f.steps == []
    assert f.run_epoch({}, [(0, 4), (4, 3)], 1) == 14 and f.g_epoch.n == 1 and f.steps == [(0, 4), (4, 3)]       # other slots
    f.same = False
    assert f.run_epoch({}, slots, 1) == 14 and f.g_epoch.n == 1                                                 # other arrays
    f.same = True
    assert f.run_epoch({}, slots, 2) == 14 and f.g_epoch.n == 1        # two ranks without the peer path: NCCL sits between the launches
    f.peer = {"any": 1}
    assert f.run_epoch({}, slots, 2) == 2 * 40 and f.g_epoch.n == 2    # peer all-reduce: one launch, bytes moved by the kernel
    f._mbc += 1
    assert f.run_epoch({}, slots, 2) == 14 and f.g_epoch.n == 2        # odd parity: the captured buffers would not alternate
    f.g_epoch = None
    assert f.run_epoch({}, slots, 1) == 14
