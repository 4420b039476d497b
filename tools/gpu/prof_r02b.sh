# round-2 ncu captures, compact: every .ncu-rep is turned into its raw-counter CSV (and the source-page CSV for the PPO kernel) on the
# box; only the small PPO report travels back.  The same command lines exited 0 without ncu in the previous call (prof_r02.sh).
cd $GRAFT_REPO_ROOT
set -x
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:$rx --launch-skip $skip -c $cnt -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  tail -1 gpurun_out/ncu_$name.log
}
cap prof_r02_ppo_tcgen05 ppo_grad_kernel_tcgen05 3 1 python tools/gpu/ppo_t5_launch.py
ncu -i /tmp/prof_r02_ppo_tcgen05.ncu-rep --page source --csv > gpurun_out/prof_r02_ppo_tcgen05.source.csv 2>/dev/null
cp /tmp/prof_r02_ppo_tcgen05.ncu-rep gpurun_out/
cap prof_r02_ppo_mma_sync ppo_grad_kernel_tc 3 1 python tools/gpu/ppo_t5_launch.py tf32
cap prof_r02_step_l1_131072 step_kernel 210 2 python bench.py --steps 5 --warmup 5 --no-sub --no-cpu-baseline
cap prof_r02_step_l4_4096 step_kernel 210 2 python bench.py --steps 5 --warmup 5 --no-sub --no-cpu-baseline --envs 4096
cap prof_r02_step_scene_fs1 step_kernel 405 1 python bench.py --workload scene --frame-skip 1 --steps 5 --warmup 5 --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02_launches.log 2>&1
ls -la gpurun_out | tail -20
