"""Phase timing of one PPO iteration (GPU box)."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from mujoco_playground_b200 import BatchedAckermannEnv
from mujoco_playground_b200.ppo import PPOConfig, PPOTrainer, compute_gae, ppo_update
torch.backends.cuda.matmul.allow_tf32 = True
env = BatchedAckermannEnv(65536, seed=1)
tr = PPOTrainer(env, PPOConfig(n_steps=16), seed=0)
for _ in range(3):
    tr.collect(); tr.update()
def timed(fn, n=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
res = {}
res["collect_ms"], _ = timed(tr.collect)
cfg, b = tr.cfg, tr.buf
def gae():
    with torch.no_grad():
        tr.graphed.value(tr.obs, tr._tv)
        tr.graphed.gae(b["rew"], b["val"], b["done"], tr._tv, tr._adv, tr._ret)
        return tr._adv, tr._ret
res["gae_ms"], (adv, ret) = timed(gae)
flat = dict(obs=b["obs"].flatten(0, 1), act=b["act"].flatten(0, 1), logp=b["logp"].flatten(), adv=adv.flatten(), ret=ret.flatten())
n = flat["obs"].shape[0]
res["randperm_ms"], perm = timed(lambda: torch.randperm(n, device=env.device))
res["shuffle_ms"], sh = timed(lambda: tr.graphed.shuffle_epoch(flat, perm))
mb = n // cfg.minibatches
res["step_ms"], _ = timed(lambda: tr.graphed.run(sh, (0, mb), 1), 20)
res["ppo_update_ms"], _ = timed(lambda: ppo_update(tr.policy, tr.opt, flat, cfg, 1, graphed=tr.graphed))
res["update_ms"], _ = timed(tr.update)
# env step alone and pieces of collect
res["env_step_ms"], _ = timed(lambda: env.step(torch.clamp(b["act"][0], -1, 1)), 20)
res["act_kernel_ms"], _ = timed(lambda: tr.graphed.act(b["obs"][0], b["act"][0], b["logp"][0], b["val"][0], 1, 1), 20)
res["value_kernel_ms"], _ = timed(lambda: tr.graphed.value(b["obs"][0], tr._tv), 20)
print(json.dumps({k: round(v, 3) for k, v in res.items()}))
# rollout: host time to ISSUE one collect() (no sync inside) against the device time of the same work
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import mujoco_playground_b200.ppo as P
_sync = torch.cuda.synchronize
t0 = time.perf_counter(); e0.record()
torch.cuda.synchronize = lambda *a, **k: None      # collect() ends with a synchronize: measure the issue time alone
try:
    tr.collect()
finally:
    torch.cuda.synchronize = _sync
t_issue = (time.perf_counter() - t0) * 1e3
e1.record(); torch.cuda.synchronize()
print(json.dumps({"collect_host_issue_ms": round(t_issue, 3), "collect_device_ms": round(e0.elapsed_time(e1), 3),
                  "lanes": int(getattr(env, "lanes_per_env", -1)) if hasattr(env, "lanes_per_env") else None}))
