#!/bin/bash
# tools/sweep.sh VAR v1 v2 ...: kernel-only bench line per value of an environment variable (run under gpurun)
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$var=$v', round(d['ms_per_step'],4), {k[:8]: round(x,4) for k,x in d['roofline']['ms_per_step_by_kernel'].items()})"
done
