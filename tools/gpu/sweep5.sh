cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in opt4 opt4ftz opt4pd opt4fm; do
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 131072 --lanes 1 --tag $v
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 4096 --lanes 4 --tag $v --iters 400
done
python tools/gpu/time_step.py --envs 131072 --dtype float64 --tag f64
python tools/gpu/time_step.py --envs 65536 --model scene --fs 1 --tag scene
python tools/gpu/time_step.py --envs 65536 --model scene --fs 4 --tag scene
python tools/gpu/time_step.py --envs 4096 --lanes 8 --iters 400 --tag l8
