cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 402 -c 1 -f -o gpurun_out/prof_r01o_scene_l4_fs1 python tools/gpu/time_step.py --envs 65536 --model scene --fs 1 --lanes 4 --iters 2 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
