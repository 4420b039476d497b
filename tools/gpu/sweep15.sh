cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/gpu/time_step.py --envs 131072 --tag main
python tools/gpu/time_step.py --envs 131072 --fs 1 --tag main
python tools/gpu/time_step.py --envs 4096 --iters 400 --tag main
python tools/gpu/time_step.py --envs 16384 --iters 200 --tag main
python tools/gpu/time_step.py --envs 65536 --model scene --fs 4 --lanes 1 --tag scene
python tools/gpu/time_step.py --envs 65536 --model scene --fs 1 --lanes 1 --tag scene
python tools/gpu/time_step.py --envs 65536 --model scene --fs 4 --lanes 4 --tag scene
python tools/gpu/time_step.py --envs 65536 --model scene --fs 1 --lanes 4 --tag scene
python tools/gpu/time_step.py --envs 131072 --dtype float64 --tag f64
