cd $GRAFT_REPO_ROOT
for v in opt3 opt4 opt4fm; do
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 131072 --lanes 1 --tag $v
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 4096 --lanes 4 --tag $v --iters 400
done
for n in 4096 8192 16384 32768 65536; do for sy in 0 1; do ACKB_CTA_SYNC=$sy ACKB_LIB=build/variants/opt4.so python tools/gpu/time_step.py --envs $n --lanes 4 --iters 200 --tag sync$sy; done; done
ACKB_CTA_SYNC=0 ACKB_LIB=build/variants/opt4.so python tools/gpu/time_step.py --envs 4096 --lanes 8 --iters 200 --tag l8sync0
