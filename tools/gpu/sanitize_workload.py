#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel layout of the step path on a ragged batch.

    compute-sanitizer --tool racecheck python tools/gpu/sanitize_workload.py

The 1-lane layout aliases the per-warp observation tile onto the wheel records in shared memory (ackb_kernels.cu) -- the code
racecheck exists for.  tests/test_gpu_sanitizer.py runs this under the three tools and requires 0 errors.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from mujoco_playground_b200 import BatchedAckermannEnv  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    cases = [("v2", 1, "float32"), ("v2", 4, "float32"), ("v2", 8, "float32"), ("scene", 4, "float32"), ("scene", 1, "float32"),
             ("maze:umaze", 4, "float32"), ("v2", 4, "float64"), ("v2", 1, "float64")]
    if which == "fast":
        cases = cases[:2] + cases[3:4]
    for model, lanes, dtype in cases:
        n = 77      # ragged: partial warps and partial CTAs
        env = BatchedAckermannEnv(n, model=model, lanes_per_env=lanes, dtype=dtype, seed=1, frame_skip=2, max_episode_steps=3)
        env.reset()
        for _ in range(4):          # includes auto-resets (3-step episodes)
            obs, rew, term, trunc, info = env.step(None)
        torch.cuda.synchronize()
        assert torch.isfinite(obs).all()
        h = [torch.zeros((n, 2)).pin_memory(), torch.zeros((n, env.obs_dim)).pin_memory(), torch.zeros(n).pin_memory(),
             torch.zeros(n, dtype=torch.uint8).pin_memory(), torch.zeros(n, dtype=torch.uint8).pin_memory()]
        env.step_host(*h)
        env.close()
        print("ok", model, lanes, dtype, flush=True)


if __name__ == "__main__":
    main()
