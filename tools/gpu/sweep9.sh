cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in tail0 tail1; do
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 131072 --lanes 1 --tag $v
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 131072 --lanes 1 --fs 1 --tag $v
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 32768 --lanes 1 --tag $v
done
python tools/gpu/time_step.py --envs 65536 --model scene --fs 4 --lanes 1 --tag scene
python tools/gpu/time_step.py --envs 65536 --model scene --fs 1 --lanes 1 --tag scene
python tools/gpu/time_step.py --envs 131072 --dtype float64 --tag f64
