"""Times the fused PPO gradient kernel alone (GPU box)."""
import ctypes, sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from mujoco_playground_b200 import _lib
from mujoco_playground_b200.ppo import ActorCritic, FusedMinibatchStep, PPOConfig
dev = torch.device("cuda:0")
N, D, mb = 1048576, 79, 262144
batch = dict(obs=torch.randn(N, D, device=dev), act=torch.randn(N, 2, device=dev).clamp(-1, 1), logp=torch.randn(N, device=dev) * 0.3 - 2.0,
             adv=torch.randn(N, device=dev), ret=torch.randn(N, device=dev))
pol = ActorCritic(D).to(dev)
opt = torch.optim.Adam(pol.parameters(), lr=3e-4)
f = FusedMinibatchStep(pol, opt, PPOConfig(), D, dev)
L = _lib.load()
idx = torch.randperm(N, device=dev)[:mb].contiguous()
p = lambda t: ctypes.c_void_p(t.data_ptr())
def kern(ix):
    L.ackb_ppo_minibatch_grad(p(batch["obs"]), p(batch["act"]), p(batch["logp"]), p(batch["adv"]), p(batch["ret"]), p(ix) if ix is not None else None, mb, D,
                              p(f.adv_stats), p(f.flat_p), p(f.flat_g), p(f.diag), 0.2, 0.5, 0.01, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
f.adv_stats[0] = 0; f.adv_stats[1] = 1
res = {}
for tag, ix in (("gather", idx), ("contiguous", None)):
    for _ in range(3): kern(ix)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): kern(ix)
    e1.record(); torch.cuda.synchronize()
    res[tag + "_ms"] = e0.elapsed_time(e1) / 20
# full step
for _ in range(3): f.run(batch, idx, 1)
torch.cuda.synchronize(); e0.record()
for _ in range(20): f.run(batch, idx, 1)
e1.record(); torch.cuda.synchronize()
res["full_step_ms"] = e0.elapsed_time(e1) / 20
res["gflop_per_step"] = mb * 111e3 / 1e9
res["tflops_kernel"] = res["gflop_per_step"] / res["gather_ms"]
print(json.dumps(res))
