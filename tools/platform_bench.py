"""Throughput of the other BASELINE configs (SOLiD 2x50, Ion Torrent 400 SE) on the resident synthetic genome."""
import sys, time, json
sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
from dwgsim_b200 import DwgsimGpu, params_from_options
import bench
FLOW = "TACGTACGTCTGAGCATCGATCGATGTACAGC"
torch.cuda.set_device(0)
lengths = [int(x * float(sys.argv[1] if len(sys.argv) > 1 else 0.25)) for x in bench.GRCH38]
for name, opts, cov, B in (("illumina_2x150", bench.OPTS, 30.0, 1 << 20),
                           ("solid_2x50", dict(length=(50, 50), data_type=1, seed=1), 30.0, 1 << 20),
                           ("ion_400_se", dict(length=(400, 0), data_type=2, e=0.01, flow_order=FLOW, seed=1), 20.0, 1 << 18)):
    g = DwgsimGpu(params_from_options(**opts))
    g.genome_synthetic(lengths, 20261017, 0.001, 0.1, 0.01, cov)
    g.genome_finalize()
    rb = 0
    for k in range(2):
        b = g.simulate_resident(k * B, B, rb); rb += b.n_random
    torch.cuda.synchronize(); t0 = time.perf_counter(); ms = [0, 0, 0]; nb = 0
    K = 10
    for k in range(2, 2 + K):
        b = g.simulate_resident(k * B, B, rb); rb += b.n_random
        ms[0] += b.ms_simulate; ms[1] += b.ms_layout; ms[2] += b.ms_format; nb += sum(b.n_bytes)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"config": name, "pairs_per_s": K * B / dt, "ms_per_step": dt / K * 1e3, "pairs_per_step": B,
                      "ms_simulate": ms[0] / K, "ms_layout": ms[1] / K, "ms_format": ms[2] / K, "fastq_bytes_per_pair": nb / (K * B)}))
    g.close()
