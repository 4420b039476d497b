cd $GRAFT_REPO_ROOT
for v in opt1 opt2 opt3 opt3cu2 opt3fm opt3b64; do
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 131072 --lanes 1 --tag $v
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 4096 --lanes 4 --tag $v --iters 400
done
for n in 8192 16384 32768; do for l in 4 1; do ACKB_LIB=build/variants/opt3.so python tools/gpu/time_step.py --envs $n --lanes $l --iters 100 --tag size; done; done
