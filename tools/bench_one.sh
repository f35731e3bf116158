#!/bin/bash
# usage: tools/bench_one.sh <workload> <instances> [steps]   -> one summary row (launch plan, env-steps/s, ms per step)
python bench.py --workload "$1" --envs "$2" --steps "${3:-100}" --warmup 10 --no-cpu 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); c = j['config']
print('$1', c['instances_per_gpu'], c['launch'], '%.4e env-steps/s' % j['value'], '%.4f ms' % j['ms_per_step'])"
