cd $GRAFT_REPO_ROOT
for v in tail0 tailK0 tailK3 tailK5 tailK6 tail1; do
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 131072 --lanes 1 --tag $v
  ACKB_LIB=build/variants/$v.so python tools/gpu/time_step.py --envs 131072 --lanes 1 --fs 1 --tag $v
done
