#!/usr/bin/env python
"""tools/cli_multi_gpu.py [gpus ...]: wall time of the drop-in binary on the 3.1 Gbp synthetic reference (-C 30, streams to
/dev/null) for each GPU count given (run under gpurun --gpus N)."""
import os, re, subprocess, sys, tempfile, time, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dwgsim_b200 import build
exe = build.build_cli()
wd = tempfile.mkdtemp(prefix="dwgsim_cli_", dir="/dev/shm")
try:
    fa = os.path.join(wd, "genome.fa")
    bases = bench.write_genome_fasta(fa, float(os.environ.get("GENOME_SCALE", "1.0")))
    for g in [int(a) for a in sys.argv[1:]] or [1]:
        prefix = os.path.join(wd, "out%d" % g)
        for f in ("bwa.read1.fastq.gz", "bwa.read2.fastq.gz", "bfast.fastq.gz"):
            os.symlink("/dev/null", prefix + "." + f)
        t0 = time.perf_counter()
        r = subprocess.run([exe] + bench.REF_ARGV + ["-C", "30", "-z", "1", "--gpus", str(g)] + os.environ.get("CLI_EXTRA", "").split() + [fa, prefix], stdout=subprocess.DEVNULL,
                           stderr=subprocess.PIPE, text=True, env=dict(os.environ, DWGSIM_STATS="1"))
        wall = time.perf_counter() - t0
        stats = [l for l in r.stderr.splitlines() if "[dwgsim_b200]" in l]
        print("gpus %d rc %d wall %.2f s | %s" % (g, r.returncode, wall, stats[-1] if stats else r.stderr[-300:]), flush=True)
        runs = [l for l in r.stderr.splitlines() if "[dwgsim_gpu_run]" in l]           # DWGSIM_RUN_TIMING=1
        if runs:
            tot = {}
            for l in runs:
                if "rank 0" not in l:
                    continue
                for key, val in re.findall(r"(waits copy|turn|sink|collect|gz \(launch|kernels|gz kernels|batches,) ([\d.]+)", l):
                    tot[key] = tot.get(key, 0.0) + float(val)
            print("   rank 0 run() totals over %d runs (ms): %s" % (len([l for l in runs if "rank 0" in l]), tot))
            for l in runs[:1] + runs[4:6]:
                print("   " + l.lstrip("\r0123456789"))
finally:
    shutil.rmtree(wd, ignore_errors=True)
