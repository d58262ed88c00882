"""Host shell on a multi-contig FASTA with and without the producer thread (DWGSIM_PIPELINE): does the GPU read loop hide
behind the next contig's prologue?   python tools/cli_pipeline_bench.py [Mbp] [contigs] [coverage]   (run under gpurun)"""
import os, shutil, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 200.0
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cov = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
wd = "/dev/shm/dwgsim_cli_pipe"; shutil.rmtree(wd, ignore_errors=True); os.makedirs(wd)
fa = os.path.join(wd, "g.fa")
rng = np.random.default_rng(5)
with open(fa, "wb") as f:
    for k in range(nc):
        n = int(mbp * 1e6 / nc)
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
        s[n // 3:n // 3 + n // 100] = ord("N")
        f.write(b">chr%d\n" % k)
        rows = s[:n - n % 60].reshape(-1, 60); out = np.empty((rows.shape[0], 61), dtype=np.uint8); out[:, :60] = rows; out[:, 60] = 10
        f.write(out.tobytes()); f.write(s[n - n % 60:].tobytes() + b"\n")
cli = os.path.join(ROOT, "dwgsim_b200", "bin", "dwgsim")
args = ["-1", "150", "-2", "150", "-e", "0.001-0.01", "-E", "0.001-0.01", "-z", "1", "-C", str(cov)]
for pipe in ("1", "0", "1", "0"):
    t = time.perf_counter()
    r = subprocess.run([cli] + args + [fa, os.path.join(wd, "o")], capture_output=True, env=dict(os.environ, DWGSIM_STATS="1", DWGSIM_PIPELINE=pipe))
    dt = time.perf_counter() - t
    tail = [l for l in r.stderr.decode(errors="ignore").split("\n") if l.startswith("[dwgsim_b200]")]
    print("DWGSIM_PIPELINE=%s rc=%d wall %.2f s | %s" % (pipe, r.returncode, dt, tail[-1][14:] if tail else r.stderr.decode()[-200:]), flush=True)
shutil.rmtree(wd, ignore_errors=True)
