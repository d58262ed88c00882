"""End-to-end timing of the host shell (dwgsim_b200/bin/dwgsim) against the compiled reference on the same FASTA:
host prologue (FASTA + mut_diref + mut_print), read loop, and the two sinks (plain files, block-parallel gzip)."""
import os, subprocess, sys, time, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
wd = "/dev/shm/dwgsim_cli_bench"
shutil.rmtree(wd, ignore_errors=True); os.makedirs(wd)
fa = os.path.join(wd, "g.fa")
bench.write_sample_fasta(fa, int(mbp * 1e6))
cli = os.path.join(ROOT, "dwgsim_b200", "bin", "dwgsim")
ref = os.path.join(ROOT, "oracle", "_ref", "dwgsim_ref")
common = ["-1", "150", "-2", "150", "-e", "0.001-0.01", "-E", "0.001-0.01", "-z", "1"]
def run(argv, env=None):
    t = time.perf_counter()
    r = subprocess.run(argv, capture_output=True, env=dict(os.environ, **(env or {})))
    dt = time.perf_counter() - t
    tail = [l for l in r.stderr.decode(errors="ignore").split("\n") if l.startswith("[dwgsim_b200]")]
    return dt, (tail[-1] if tail else ""), r.returncode
for label, extra in (("b200 prologue only (-C 0)", ["-C", "0"]), ("b200 --uncompressed", ["-C", str(cov), "--uncompressed"]), ("b200 .gz", ["-C", str(cov)])):
    dt, tail, rc = run([cli] + common + extra + [fa, os.path.join(wd, "b200")], {"DWGSIM_STATS": "1"})
    print("%-28s rc=%d wall %.2f s | %s" % (label, rc, dt, tail), flush=True)
    for f in os.listdir(wd):
        if f.startswith("b200."): print("     ", f, os.path.getsize(os.path.join(wd, f)))
if os.path.exists(ref):
    dt0, _, _ = run([ref] + common + ["-C", "0", fa, os.path.join(wd, "ref")])
    n = 200000
    dt1, _, _ = run([ref] + common + ["-N", str(n), fa, os.path.join(wd, "ref")])
    print("reference prologue only (-C 0): %.2f s (%.1f ns/base); -N %d: %.2f s => loop %.1f kpairs/s (1 core, .gz)" % (dt0, 1e9 * dt0 / (mbp * 1e6), n, dt1, n / max(dt1 - dt0, 1e-9) / 1e3))
shutil.rmtree(wd, ignore_errors=True)
