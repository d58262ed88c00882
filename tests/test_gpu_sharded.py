"""-m gpu: the sharded run path (world_size 2 over gloo, both ranks on cuda:0) must reproduce the unsharded bytes."""
import os
import pickle
import socket
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

pytestmark = pytest.mark.gpu
OPTS = dict(seed=3, N=6000, length=(100, 100), mut_rate=0.02, indel_frac=0.5, rand_read=0.2)


def _worker(rank, world, port, fasta, tmp, batch):
    import gpu_harness as gh
    from oracle import pyoracle
    from dwgsim_b200 import DwgsimGpu, params_from_options, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        opt = pyoracle.make_opt(**OPTS)
        sess = pyoracle.Session(opt, fasta, None, mode=pyoracle.RNG_PHILOX, keep=True)   # plays the reference host
        got = [[], [], []]
        with DwgsimGpu(params_from_options(**{k: v for k, v in OPTS.items() if k in gh.GPU_KEYS})) as gpu:
            gpu.set_batch(batch, 2)
            gpu.set_shard(rank, world)
            gpu.set_exchange(shard.make_exchange())
            for k in range(sess.n_contigs):
                c = sess.contig(k)
                gpu.add_contig(c["contig_i"], c["name"], c["seq"], c["len"], c["hap"][0], c["hap"][1], c["ins"][0],
                               c["n_ins"][0], c["ins"][1], c["n_ins"][1], c["n_pairs"])
            st = gpu.run(lambda fid, data: got[fid].append(data))
        sess.close()
        with open(os.path.join(tmp, "rank%d.pkl" % rank), "wb") as f:
            pickle.dump((got, st.n_pairs, st.n_random), f)
    finally:
        dist.destroy_process_group()


def test_two_ranks_reproduce_unsharded_bytes(oracle, synth_fa, tmp_path):
    import gpu_harness as gh
    from dwgsim_b200 import shard
    sess, want = gh.oracle_expected(oracle, OPTS, synth_fa, str(tmp_path / "orc"))
    n_total, n_random = sess.stats.n_pairs_total, sess.stats.n_random
    sess.close()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ps = [ctx.Process(target=_worker, args=(r, 2, port, synth_fa, str(tmp_path), 700)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(timeout=300)
        assert p.exitcode == 0
    parts = [pickle.load(open(str(tmp_path / ("rank%d.pkl" % r)), "rb")) for r in range(2)]
    assert sum(p[1] for p in parts) == n_total and sum(p[2] for p in parts) == n_random
    for fid, name in enumerate(gh.FILE_NAMES):
        merged = shard.interleave([parts[0][0][fid], parts[1][0][fid]])
        assert merged == want[fid], "%s: %s" % (name, gh.first_diff(want[fid], merged))
