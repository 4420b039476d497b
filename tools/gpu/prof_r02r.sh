# final captures of round 2 (one gpurun call): plain runs first, ncu only after they exited 0
cd $GRAFT_REPO_ROOT
set -x
timeout 600 python bench.py > gpurun_out/r02r_bench_default.json 2> gpurun_out/r02r_bench_default.err || exit 1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_bench_short.json 2> gpurun_out/r02r_bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02r_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02r_launches.log 2>&1
tail -2 gpurun_out/ncu_r02r_launches.log
timeout 120 python tools/gpu/ppo_t5_launch.py > gpurun_out/r02r_ppo_plain.log 2>&1 &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:ppo_grad_kernel_tcgen05 -s 3 -c 1 -f -o gpurun_out/prof_r02r_ppo_tcgen05 python tools/gpu/ppo_t5_launch.py > gpurun_out/ncu_r02r_ppo.log 2>&1
tail -2 gpurun_out/ncu_r02r_ppo.log
timeout 200 python bench.py --steps 5 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/r02r_b131k_plain.json 2> gpurun_out/r02r_b131k_plain.err &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 210 -c 2 -f -o gpurun_out/prof_r02r_step_l1_131072 python bench.py --steps 5 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/ncu_r02r_step131k.log 2>&1
tail -2 gpurun_out/ncu_r02r_step131k.log
ls -la gpurun_out | tail -12
