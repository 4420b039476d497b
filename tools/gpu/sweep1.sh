cd $GRAFT_REPO_ROOT
for v in base b64m5 b64m4 b96m3 b32m10 b32m8; do
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 131072 --lanes 1 --tag $v
done
for s in "ls_fast_iters=2 ls_mid_iters=4" "ls_fast_iters=0 ls_mid_iters=0" "ls_fast_iters=2 ls_mid_iters=6" "ls_fast_iters=3 ls_mid_iters=5" "ls_fast_iters=4 ls_mid_iters=6"; do
  ACKB_LIB=build/variants/base.so python tools/gpu/time_step.py --envs 131072 --lanes 1 --tag ls --set $s
  ACKB_LIB=build/variants/base.so python tools/gpu/time_step.py --envs 4096 --lanes 4 --tag ls --set $s --iters 400
done
for l in 4 8 1; do python tools/gpu/time_step.py --envs 4096 --lanes $l --iters 400 --tag lanes; done
for n in 8192 16384 32768 65536; do for l in 4 1; do python tools/gpu/time_step.py --envs $n --lanes $l --iters 100 --tag size; done; done
