# tools/fmt_sweep.sh: format-kernel time over warps per CTA x pairs per mini-tile (DWGSIM_FMT_WARPS, DWGSIM_TILE_PAIRS)
run() { echo -n "warps=$1 WP=$2: "; DWGSIM_FMT_WARPS=$1 DWGSIM_TILE_PAIRS=$2 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-e2e-cli 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['roofline']['ms_per_batch_by_kernel']['format_fastq_kernel'])"; }
run 24 4
run 24 3
run 22 4
run 20 5
run 18 5
run 16 6
run 24 2
