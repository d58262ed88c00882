# tools/fmt_variants.sh: rebuild the library with extra nvcc flags on the GPU box and time the format kernel (experiments)
run() { echo -n "$1: "; python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-e2e-cli 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['roofline']['ms_per_batch_by_kernel'])"; }
for v in "$@"; do
  NVCC_EXTRA="$v" python -c "from dwgsim_b200 import build; build.build(force=True)" && run "$v"
done
