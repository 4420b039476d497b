cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/gpu/time_step.py --envs 131072 --tag main
python tools/gpu/time_step.py --envs 131072 --fs 1 --tag main
python tools/gpu/time_step.py --envs 4096 --iters 400 --tag main
python tools/gpu/time_step.py --envs 131072 --dtype float64 --tag f64
bash tools/gpu/prof.sh > /dev/null 2>&1
ls gpurun_out/*.ncu-rep
