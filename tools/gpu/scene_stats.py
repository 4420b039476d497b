import sys, numpy as np, torch
sys.path.insert(0, '.')
from mujoco_playground_b200 import BatchedAckermannEnv
for jit in (0.1, 0.12, 0.15):
    n = 65536
    env = BatchedAckermannEnv(n, model="scene", dtype="float32", seed=5, frame_skip=4, spawn_yaw_range=np.pi, spawn_xy_jitter=jit)
    env.reset()
    wb = 0; mx = 0
    for _ in range(250):
        obs, rew, term, trunc, info = env.step(None)
        nc = info["ncon"]; wb += int((nc > 8).sum().item()); mx = max(mx, int(nc.max().item()))
    q, v, _ = env.get_state()
    st = env.stats()
    print(f"jitter {jit}: z [{q[:,2].min():.3f}, {q[:,2].max():.3f}] frac z<0.08 {(q[:,2]<0.08).mean():.4f} frac z<0.05 {(q[:,2]<0.05).mean():.5f} box-contact frac {wb/(250*n):.5f} max ncon {mx} unsupported {st['unsupported']/st['env_steps']:.5f} |x|max {np.abs(q[:,0]).max():.2f} |y|max {np.abs(q[:,1]).max():.2f} episodes {st['episodes']} obstacle_steps frac {st['obstacle_steps']/st['env_steps']:.5f} mean ncon {st['contacts_sum']/st['env_steps']:.3f}")
    env.close()
