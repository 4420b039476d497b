cd $GRAFT_REPO_ROOT
for n in 4096 16384 65536 131072; do for z in 0 1; do
ACKB_ZERO_COPY=$z python bench.py --envs $n --steps 200 --warmup 20 --no-aux --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('n', $n, 'zero_copy', $z, 'value %.1fM' % (d['value']/1e6), 'e2e %.1fM' % (d['e2e']['value']/1e6))"
done; done
