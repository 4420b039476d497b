cd $GRAFT_REPO_ROOT
for v in base opt1 opt1cu2 opt1fm opt1b32; do
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 131072 --lanes 1 --tag $v
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 4096 --lanes 4 --tag $v --iters 400
done
