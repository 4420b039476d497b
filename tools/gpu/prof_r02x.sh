# final capture of the configs[1] step kernel (4096 envs, 4 lanes per env) with the list-walk general pass
cd $GRAFT_REPO_ROOT
set -x
timeout 200 python bench.py --steps 5 --warmup 5 --no-sub --no-cpu-baseline --envs 4096 > gpurun_out/r02x_b4096_plain.json 2> gpurun_out/r02x_b4096_plain.err &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 210 -c 2 -f -o gpurun_out/prof_r02x_step_l4_4096 python bench.py --steps 5 --warmup 5 --no-sub --no-cpu-baseline --envs 4096 > gpurun_out/ncu_r02x_step4096.log 2>&1
tail -2 gpurun_out/ncu_r02x_step4096.log
