"""-m gpu: statistics of the CUDA path at a size the CPU oracle would not finish quickly (1 M reads), checked against
the closed forms of the REFERENCE's distributions (not against the oracle): per-cycle error rate p_i = start + by*i
(src/dwgsim.c:237), per-cycle quality histogram, insert size N(500, 50), strand, random-pair
fraction -y; plus determinism (same seed => same bytes, different seed => different bytes)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
LEN, NPAIRS, GENOME = 100, 500_000, 2_000_000
COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    COMP[a] = b


def run_gpu(seed, n_pairs):
    from dwgsim_b200 import DwgsimGpu, params_from_options
    rng = np.random.default_rng(99)
    codes = rng.integers(0, 4, GENOME, dtype=np.uint8)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].copy()
    hap = codes.astype(np.uint64)
    p = params_from_options(seed=seed, length=(LEN, LEN), e="0.002-0.03", E="0.01-0.05", rand_read=0.07, reads_output_type=1)
    with DwgsimGpu(p) as gpu:
        gpu.add_contig(0, "c1", seq.ctypes.data, GENOME, hap.ctypes.data, hap.ctypes.data, None, 0, None, 0, n_pairs)
        streams, st = gpu.run_collect()
    return seq, streams, st


def parse(stream, n):
    """fixed-geometry parse: every record = name line + LEN bases + '+' + LEN quals"""
    lines = stream.split(b"\n")
    assert len(lines) == 4 * n + 1
    names = lines[0::4][:n]
    seqs = np.frombuffer(b"".join(lines[1::4][:n]), dtype=np.uint8).reshape(n, LEN)
    quals = np.frombuffer(b"".join(lines[3::4][:n]), dtype=np.uint8).reshape(n, LEN) - 33
    return names, seqs, quals


def test_gpu_fastq_statistics_match_reference_closed_forms():
    from scipy.stats import norm
    ref, streams, st = run_gpu(seed=5, n_pairs=NPAIRS)
    assert st.n_pairs == NPAIRS
    z = (st.n_random - NPAIRS * 0.07) / math.sqrt(NPAIRS * 0.07 * 0.93)
    assert abs(z) < 4.5
    ins, strand0 = [], []
    for end, (s, e) in enumerate(((0.002, 0.03), (0.01, 0.05))):
        names, seqs, quals = parse(streams[end], NPAIRS)
        f = [nm.rsplit(b"_", 9) for nm in names]
        genomic = np.array([x[0] != b"@rand" for x in f])
        pos = np.array([int(x[1 + end]) for x in f])[genomic] - 1
        strand = np.array([int(x[3 + end]) for x in f])[genomic]
        idx = pos[:, None] + np.arange(LEN)[None, :]
        truth = ref[idx]
        rc = COMP[truth][:, ::-1]
        truth = np.where(strand[:, None] == 1, rc, truth)
        mism = (seqs[genomic] != truth).sum(0)
        n = genomic.sum()
        p = s + (e - s) / LEN * np.arange(LEN)
        zc = (mism - n * p) / np.sqrt(n * p * (1 - p))
        assert np.abs(zc).max() < 5.0, (end, float(np.abs(zc).max()))
        assert abs(zc.sum() / math.sqrt(LEN)) < 4.5
        # quality histogram of a few cycles against the closed form (sigma 2, truncation toward zero, clamp 0..40)
        for i in (0, 37, 99):
            pi = s + (e - s) / LEN * i
            qb = int(-10.0 * math.log(pi) / math.log(10.0) + 0.499)
            prob = np.zeros(41)
            for d in range(-20, 21):
                if d == 0:
                    pr = norm.cdf(0.25) - norm.cdf(-0.75)
                elif d > 0:
                    pr = norm.cdf((d + 0.5) / 2) - norm.cdf((d - 0.5) / 2)
                else:
                    pr = norm.cdf((d - 0.5) / 2) - norm.cdf((d - 1.5) / 2)
                prob[min(max(qb + d, 0), 40)] += pr
            obs = np.bincount(quals[:, i], minlength=41)
            keep = prob * NPAIRS > 20
            zq = (obs - prob * NPAIRS)[keep] / np.sqrt((prob * NPAIRS * (1 - prob))[keep])
            assert np.abs(zq).max() < 5.0, (end, i, float(np.abs(zq).max()))
        if end == 0:
            p2 = np.array([int(x[2]) for x in f])[genomic] - 1
            ins = np.abs(p2 - pos) + LEN
            strand0 = strand
    assert abs(ins.mean() - 500.0) < 5 * 50 / math.sqrt(len(ins)) and abs(ins.std() - 50.0) < 0.5
    assert abs(strand0.mean() - 0.5) < 4.5 * 0.5 / math.sqrt(len(strand0))


def test_determinism_and_seed_sensitivity():
    _, a, _ = run_gpu(seed=11, n_pairs=20000)
    _, b, _ = run_gpu(seed=11, n_pairs=20000)
    _, c, _ = run_gpu(seed=12, n_pairs=20000)
    assert a == b
    assert a[0] != c[0]
