#!/bin/bash
# usage: tools/lanes_quick.sh <workload> <lanes...>  -> env-steps/s of the fused step per lane count (one row each)
w=$1; shift
for l in "$@"; do
python bench.py --workload "$w" --lanes "$l" --steps 100 --warmup 10 --no-cpu --kernel-only 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); c = j['config']
print('$w lanes=$l', c['launch'], '%.4e env-steps/s' % j['value'], '%.4f ms' % j['ms_per_step'], 'it %.2f' % j['mean_iterations'])"
done
