#!/bin/bash
# usage: tools/lanes_one.sh <workload> <lanes> [instances] [steps]
python bench.py --workload "$1" --lanes "$2" ${3:+--envs $3} --steps "${4:-100}" --warmup 10 --no-cpu 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); c = j['config']
print('$1 lanes=$2', c['instances_per_gpu'], c['launch'], '%.4e env-steps/s' % j['value'], '%.4f ms' % j['ms_per_step'])"
