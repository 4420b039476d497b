# round-2 ncu captures (one gpurun call): plain run first, ncu only if it exited 0
cd $GRAFT_REPO_ROOT
set -x
timeout 120 python tools/gpu/ppo_t5_launch.py > gpurun_out/r02_ppo_plain.log 2>&1 &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:ppo_grad_kernel_tcgen05 -s 3 -c 1 -f -o gpurun_out/prof_r02_ppo_tcgen05 python tools/gpu/ppo_t5_launch.py > gpurun_out/ncu_r02_ppo.log 2>&1
tail -2 gpurun_out/ncu_r02_ppo.log
timeout 200 python bench.py --steps 5 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/r02_b131k_plain.json 2> gpurun_out/r02_b131k_plain.err &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 210 -c 2 -f -o gpurun_out/prof_r02_step_l1_131072 python bench.py --steps 5 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/ncu_r02_step131k.log 2>&1
tail -2 gpurun_out/ncu_r02_step131k.log
timeout 200 python bench.py --steps 5 --warmup 5 --no-sub --no-cpu-baseline --envs 4096 > gpurun_out/r02_b4096_plain.json 2> gpurun_out/r02_b4096_plain.err &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 210 -c 2 -f -o gpurun_out/prof_r02_step_l4_4096 python bench.py --steps 5 --warmup 5 --no-sub --no-cpu-baseline --envs 4096 > gpurun_out/ncu_r02_step4096.log 2>&1
tail -2 gpurun_out/ncu_r02_step4096.log
timeout 200 python bench.py --workload scene --frame-skip 1 --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scene_plain.json 2> gpurun_out/r02_scene_plain.err &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 405 -c 1 -f -o gpurun_out/prof_r02_step_scene_fs1 python bench.py --workload scene --frame-skip 1 --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_r02_scene.log 2>&1
tail -2 gpurun_out/ncu_r02_scene.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02_launches.log 2>&1
tail -2 gpurun_out/ncu_r02_launches.log
ls -la gpurun_out | tail -20
