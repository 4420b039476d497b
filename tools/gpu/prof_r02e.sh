# final evidence of the round: tests, default bench, reference arm, smoke, ncu of the two tcgen05 learner kernels, launch list
cd $GRAFT_REPO_ROOT
set -x
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:$rx --launch-skip $skip -c $cnt -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  tail -1 gpurun_out/ncu_$name.log
}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; tail -3 gpurun_out/r02m_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; tail -c 300 gpurun_out/r02m_bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02m_ref.json 2>/dev/null; tail -c 200 gpurun_out/r02m_ref.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02m_smoke.log 2>&1; tail -2 gpurun_out/r02m_smoke.log
cap prof_r02m_ppo_grad ppo_grad_kernel_tcgen05 3 1 python tools/gpu/ppo_t5_launch.py
cap prof_r02m_ppo_act ppo_act_kernel_tcgen05 8 1 python bench.py --workload ppo --steps 2 --warmup 1 --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02m_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02m_launches.log 2>&1
ls -la gpurun_out | tail -8
