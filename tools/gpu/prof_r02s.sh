# per-line profile of the scene kernel (configs[2], 65536 envs, frame_skip 1): share and divergence of the lidar cell walk
cd $GRAFT_REPO_ROOT
set -x
timeout 200 python bench.py --workload scene --frame-skip 1 --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/r02s_scene_plain.json 2> gpurun_out/r02s_scene_plain.err &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 405 -c 1 -f -o gpurun_out/prof_r02s_step_scene_fs1 python bench.py --workload scene --frame-skip 1 --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_r02s_scene.log 2>&1
tail -2 gpurun_out/ncu_r02s_scene.log
