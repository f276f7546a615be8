#!/bin/bash
# A/B of one environment switch on one box:  bash tools/ab_env.sh PD_STREAMK_CPS 36 37 36 37
# prints "<VAR>=<value> <device sample-steps/s> <e2e sample-steps/s>" per run (short bench, no extra legs)
VAR=$1; shift
for v in "$@"; do
    env $VAR=$v python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ka --no-tf32 --no-extras 2>/dev/null > /tmp/ab_line.json
    python - "$VAR=$v" <<'PY'
import json, sys
d = json.loads(open("/tmp/ab_line.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"], 1), round(d["e2e"]["value"], 1))
PY
done
