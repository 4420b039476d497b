# final run of the round: default bench, reference arm, smoke, launch list of the default bench command
cd $GRAFT_REPO_ROOT
set -x
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; tail -c 300 gpurun_out/r02j_bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02j_ref.json 2>/dev/null; tail -c 200 gpurun_out/r02j_ref.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_smoke.log 2>&1; tail -2 gpurun_out/r02j_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02j_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02j_launches.log 2>&1
ls -la gpurun_out | tail -6
