import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from dwgsim_b200 import DwgsimGpu, params_from_options
torch.cuda.set_device(0)
seq, hap = bench.dense_contig(bench.E2E_CONTIG_LEN, 7)
n_pairs_c = int(bench.E2E_CONTIG_LEN * 30 / 300.0 / 0.95 + 0.5)
for mode in (0, 1):
    g = DwgsimGpu(params_from_options(**bench.OPTS)); g.set_batch(1 << 18, 3); g.set_compression(mode)
    for i in range(10):
        t0 = time.perf_counter()
        g.add_contig(i, "chrE%d" % i, seq.ctypes.data, bench.E2E_CONTIG_LEN, hap[0].ctypes.data, hap[1].ctypes.data, None, 0, None, 0, n_pairs_c)
        st = g.run_count(); dt = time.perf_counter() - t0
        if i >= 5: print("mode %d step %d: %.1f ms total (pack %.1f, kernels %.1f, gzip %.1f) d2h %.0f MB -> %.1f Mpairs/s" % (mode, i, dt * 1e3, st.ms_pack, st.ms_simulate + st.ms_layout + st.ms_format, st.ms_compress, st.d2h_bytes / 1e6, n_pairs_c / dt / 1e6))
    g.close()
