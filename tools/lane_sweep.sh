#!/bin/bash
# usage: tools/lane_sweep.sh <workload> <lanes...>   -> one line per lane count (device-resident value only)
w=$1; shift
for l in "$@"; do
  python bench.py --workload $w --lanes $l --steps 100 --warmup 20 --no-cpu 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    try: j=json.loads(ln)
    except Exception: continue
    li=j['config']['launch']
    print('$w lanes=%d thr=%d grid=%d smem=%d | %.4e env-steps/s %.3f ms | e2e %.4e | it %.2f conv %.4f' % (li['lanes'],li['threads'],li['grid'],li['smem_bytes'],j['value'],j['ms_per_step'],j['e2e']['value'],j['mean_iterations'],j['converged_frac']))
"
done
