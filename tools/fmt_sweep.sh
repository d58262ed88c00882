# tools/fmt_sweep.sh: format-kernel time over kernel variants x generation (DWGSIM_FORMAT) x warps x pairs per tile
run() { echo -n "lib=$1 format=$2 warps=$3 TP=$4: "; DWGSIM_LIB=$PWD/variants/$1 DWGSIM_FORMAT=$2 DWGSIM_FMT_WARPS=$3 DWGSIM_TILE_PAIRS=$4 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), d['roofline']['ms_per_step_by_kernel']['format_fastq_kernel'])"; }
run lib_pf1.so 3 10 48
run lib_pf0.so 3 10 48
run lib_pf1.so 3 12 48
run lib_pf0.so 3 12 48
run lib_pf1.so 3 12 36
run lib_pf1.so 2 24 4
