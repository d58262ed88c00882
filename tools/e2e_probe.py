#!/usr/bin/env python
"""tools/e2e_probe.py [gz] [batch_log2]: per-step breakdown of the e2e leg of bench.py (run under gpurun)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dwgsim_b200 import DwgsimGpu, params_from_options

gz = "gz" in sys.argv
lg = [int(a) for a in sys.argv[1:] if a.isdigit()]
batch = 1 << (lg[0] if lg else 18)
seq, hap = bench.dense_contig(bench.E2E_CONTIG_LEN, 7)
n_pairs = int(bench.E2E_CONTIG_LEN * bench.COVERAGE / 300.0 / 0.95 + 0.5)
g = DwgsimGpu(params_from_options(**bench.OPTS), device=0)
g.set_batch(batch, 3)
if gz:
    g.set_compression(1)
for i in range(14):
    t0 = time.perf_counter()
    g.add_contig(i, "chrE%d" % i, seq.ctypes.data, bench.E2E_CONTIG_LEN, hap[0].ctypes.data, hap[1].ctypes.data, None, 0, None, 0, n_pairs)
    t1 = time.perf_counter()
    st = g.run_count()
    t2 = time.perf_counter()
    if i >= 6:
        print("step %2d add %.1f ms run %.1f ms | lib total %.1f pack %.1f sim %.2f lay %.2f fmt %.2f gz %.2f | d2h %.0f MB -> %.1f GB/s if copy-bound | %.1f M pairs/s" % (
            i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st.ms_total, st.ms_pack, st.ms_simulate, st.ms_layout, st.ms_format, st.ms_compress,
            st.d2h_bytes / 1e6, st.d2h_bytes / 1e9 / ((t2 - t1)), n_pairs / (t2 - t0) / 1e6))
g.close()
