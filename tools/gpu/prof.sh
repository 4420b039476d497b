set -x
cd $GRAFT_REPO_ROOT
python bench.py --envs 131072 --steps 20 --warmup 5 --no-aux --no-cpu-baseline > gpurun_out/b131k.json 2> gpurun_out/b131k.err
cat gpurun_out/b131k.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 108 -c 1 -f -o gpurun_out/prof_r01p_l1_131072 python bench.py --envs 131072 --steps 5 --warmup 5 --no-aux --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 108 -c 1 -f -o gpurun_out/prof_r01p_l4_4096 python bench.py --steps 5 --warmup 5 --no-aux --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out
