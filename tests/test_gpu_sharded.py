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


def test_device_side_rand_base_equals_host_path():
    """resident_begin + resident_finish_dev (rand_serial_base read from device memory, the NCCL path of bench.py) writes the
    bytes of simulate_resident with the same base on the host, and the device counter holds the batch's random pairs"""
    import torch
    from dwgsim_b200 import DwgsimGpu, params_from_options
    with DwgsimGpu(params_from_options(seed=5, length=(100, 100), rand_read=0.1)) as gpu:
        gpu.genome_synthetic([300000, 200000], 11, 0.002, 0.2, 0.01, 10.0)
        gpu.genome_finalize()
        n, base = 20000, 123456789
        b0 = gpu.simulate_resident(1000, n, base)
        want = [gpu.copy_stream(k, b0.n_bytes[k]) for k in range(3)]
        gpu.resident_begin(1000, n, False)
        stream = torch.cuda.ExternalStream(gpu.cuda_stream())

        class Arr:
            def __init__(self, ptr, nbytes):
                self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

        cnt = torch.as_tensor(Arr(gpu.resident_count_ptr(), 8), device="cuda").view(torch.int64)
        base_t = torch.zeros(1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            base_t.fill_(base)
            b1 = gpu.resident_finish_dev(base_t.data_ptr())
            seen = int(cnt.item())
        assert seen == b1.n_random == b0.n_random and b0.n_random > 0
        assert [gpu.copy_stream(k, b1.n_bytes[k]) for k in range(3)] == want
        assert b"rand_0_0_0_0_1_1_0:0:0_0:0:0_%x" % base in want[0]          # the first random pair carries the base as rand_ii


def test_finish_gathered_takes_the_prefix_of_the_counts_on_the_device():
    """dwgsim_gpu_resident_finish_gathered (bench.py at N > 1): rand_serial_base = running count + counts of the ranks before
    this one, computed by a kernel of the library from the all-gathered counts; the running count advances by all of them"""
    import torch
    from dwgsim_b200 import DwgsimGpu, params_from_options
    with DwgsimGpu(params_from_options(seed=5, length=(100, 100), rand_read=0.1)) as gpu:
        gpu.genome_synthetic([300000, 200000], 11, 0.002, 0.2, 0.01, 10.0)
        gpu.genome_finalize()
        n, start = 10000, 1000
        stream = torch.cuda.ExternalStream(gpu.cuda_stream())

        class Arr:
            def __init__(self, ptr, nbytes):
                self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

        gpu.resident_set_running(700)
        running = 700
        for rnd, (before, after) in enumerate(((17, 5), (0, 9))):          # this rank is rank 1 of 3
            first = start + rnd * n
            gpu.resident_begin(first, n, False)
            cnt = torch.as_tensor(Arr(gpu.resident_count_ptr(), 8), device="cuda").view(torch.int64)
            allc = torch.zeros(3, dtype=torch.int64, device="cuda")
            with torch.cuda.stream(stream):
                allc[0] = before
                allc[1:2].copy_(cnt)
                allc[2] = after
                gpu.resident_finish_gathered(allc.data_ptr(), 3, 1)
            b = gpu.resident_wait()
            got = [gpu.copy_stream(k, b.n_bytes[k]) for k in range(3)]
            want_b = gpu.simulate_resident(first, n, running + before)
            assert got == [gpu.copy_stream(k, want_b.n_bytes[k]) for k in range(3)] and b.n_random == want_b.n_random > 0
            running += before + b.n_random + after
