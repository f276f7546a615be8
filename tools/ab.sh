run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ka > gpurun_out/ab_$name.json 2>gpurun_out/ab_$name.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/ab_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), round(d['e2e']['value'],1))
"; }
