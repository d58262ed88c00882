python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python tools/fuzz_gpu_vs_oracle.py 501 400 2>&1 | tail -1
python tools/fuzz_gpu_vs_oracle.py 502 150 random-fasta 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -2 gpurun_out/r02b_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-e2e --no-cpu-baseline --no-configs --no-e2e-cli > /dev/null 2>&1
SKIP=63 COUNT=9 bash tools/ncu_capture.sh r02b_step '.*'
