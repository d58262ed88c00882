#!/bin/bash
# tools/sweep_env.sh VAR v1 v2 ...: kernel times of one device batch per value of an environment knob (run under gpurun)
var=$1; shift
for v in "$@"; do
  echo -n "$var=$v: "; env $var=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-e2e-cli 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['roofline']['ms_per_batch_by_kernel'])"
done
