# per-line profile of the 1-lane step kernel at frame_skip 1 (the PPO rollout's shape: 65536 envs): what the env step costs besides
# the physics substep
cd $GRAFT_REPO_ROOT
set -x
timeout 120 python tools/gpu/time_step.py --envs 65536 --fs 1 --iters 20 > gpurun_out/r02q_plain.log 2>&1 &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 900 -c 1 -f -o gpurun_out/prof_r02q_step_l1_65536_fs1 python tools/gpu/time_step.py --envs 65536 --fs 1 --iters 20 > gpurun_out/ncu_r02q.log 2>&1
tail -2 gpurun_out/ncu_r02q.log
ls -la gpurun_out | tail -5
