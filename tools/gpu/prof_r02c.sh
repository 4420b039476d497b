# round-2 final captures (same recipe as prof_r02b.sh): new PPO gradient kernel and the scene kernel with the extra CTA barrier
cd $GRAFT_REPO_ROOT
set -x
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:$rx --launch-skip $skip -c $cnt -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  tail -1 gpurun_out/ncu_$name.log
}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; tail -3 gpurun_out/r02h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -c 600 gpurun_out/r02h_bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02h_ref.json 2>/dev/null; tail -c 400 gpurun_out/r02h_ref.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; tail -2 gpurun_out/r02h_smoke.log
cap prof_r02e_ppo_tcgen05 ppo_grad_kernel_tcgen05 3 1 python tools/gpu/ppo_t5_launch.py
cp /tmp/prof_r02e_ppo_tcgen05.ncu-rep gpurun_out/
cap prof_r02e_step_scene_fs1 step_kernel 405 1 python bench.py --workload scene --frame-skip 1 --steps 5 --warmup 5 --no-cpu-baseline
ls -la gpurun_out | tail -12
