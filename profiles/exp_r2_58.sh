# k_solve_big CTAs per SM under a steady stream of auto-resets (c3-resets: 26.8 k big islands per step, k_solve_big ends 2.6 ms after k_pre)
for B in 1 2 3; do
for c in c3-resets c3; do
MRP_BIG_CTAS=$B python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('big_ctas $B', '$c', 'value %.4e ms %.3f e2e %.4e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done; done
