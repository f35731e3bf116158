for w in ieee13 ieee13_newton ieee34 ieee123; do for e in 4096 16384; do python bench.py --workload $w --envs $e --steps 200 --warmup 20 --no-cpu 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=j['config']
print('$w', c['instances_per_gpu'], c['launch'], '%.4e'%j['value'], '%.4f ms'%j['ms_per_step'])"; done; done
