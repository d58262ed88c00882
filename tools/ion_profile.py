import sys; sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch, bench
from dwgsim_b200 import DwgsimGpu, params_from_options
FLOW = "TACGTACGTCTGAGCATCGATCGATGTACAGC"
torch.cuda.set_device(0)
g = DwgsimGpu(params_from_options(length=(400, 0), data_type=2, e=0.01, flow_order=FLOW, seed=1))
g.genome_synthetic([int(x * 0.05) for x in bench.GRCH38], 20261017, 0.001, 0.1, 0.01, 20.0)
g.genome_finalize()
for k in range(3):
    g.simulate_resident(k * (1 << 17), 1 << 17, 0)
g.close()
