# tools/fmt_ab.sh: parity tests, then the format kernel's time for several DWGSIM_FMT_RUN settings (run under gpurun)
run() { echo -n "run=$1: "; DWGSIM_FMT_RUN=$1 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-e2e-cli 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['roofline']['ms_per_batch_by_kernel'])"; }
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
DWGSIM_FMT_RUN=3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for r in "$@"; do run $r; done
