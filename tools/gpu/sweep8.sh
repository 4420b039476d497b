cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 compute-sanitizer --tool memcheck python tools/gpu/time_step.py --envs 2048 --model scene --fs 2 --lanes 1 --iters 2 --tag sanit 2>&1 | tail -8
python tools/gpu/time_step.py --envs 65536 --model scene --fs 1 --lanes 1 --tag scene
python tools/gpu/time_step.py --envs 65536 --model scene --fs 4 --lanes 1 --tag scene
python tools/gpu/time_step.py --envs 131072 --dtype float64 --tag f64
