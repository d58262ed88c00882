"""Link 2 of the parity chain for the other two platforms: the oracle's Philox backend samples the same
distributions as its drand48 backend (= the reference, byte for byte) in SOLiD colour space (BASELINE configs[3]) and under
the Ion Torrent flow model (configs[4]).  Compared between the backends, and against the closed form where there is one:
per-cycle colour-error rate, per-cycle quality histogram, error-count fields of the names, Ion Torrent read lengths."""
import math
import os
import re

import numpy as np
import pytest

from test_oracle_statistics import write_fasta, z_two

COMP = bytes.maketrans(b"ACGTN", b"TGCAN")
NAME = re.compile(rb"^@(.+)_(\d+)_(\d+)_([01])_([01])_([01])_([01])_(\d+):(\d+):(\d+)_(\d+):(\d+):(\d+)_([0-9a-f]+)$")
FLOW = "TACGTACGTCTGAGCATCGATCGATGTACAGC"


def records(path):
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(0, len(lines) - 1, 4):
        yield lines[i], lines[i + 1], lines[i + 3]


# ---- SOLiD 2x50 ---------------------------------------------------------------------------------------------------------
S_LEN, S_PAIRS = 50, 20000


def solid_stats(oracle, mode, fasta, ref, tmp):
    opts = dict(seed=21, N=S_PAIRS, data_type=1, length=(S_LEN, S_LEN), e="0.005-0.04", E="0.02", mut_rate=0, rand_read=0.05,
                reads_output_type=2)
    prefix = os.path.join(tmp, "solid%d" % mode)
    with oracle.Session(oracle.make_opt(**opts), fasta, prefix, mode=mode) as s:
        assert s.stats.error == 0
    err = np.zeros((2, S_LEN)); n = np.zeros(2); qual = np.zeros((2, S_LEN, 41)); nerr = [[], []]
    code = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}
    for k, (name, seq, q) in enumerate(records(prefix + ".bfast.fastq")):       # end 1 then end 2 of every pair
        end = k & 1
        g = NAME.match(name).groups()
        qual[end, np.arange(S_LEN), np.frombuffer(q, dtype=np.uint8) - 33] += 1
        if g[0] == b"rand":
            continue
        assert seq[:1] == b"A" and len(seq) == S_LEN + 1
        pos, strand = int(g[1 + end]), int(g[3 + end])
        truth = ref[pos - 1:pos - 1 + S_LEN]
        if strand:
            truth = truth.translate(COMP)[::-1]
        b = np.array([code[c] for c in truth])
        colours = b ^ np.concatenate(([0], b[:-1]))                             # adaptor base A = 0, src/dwgsim.c:845-858
        got = np.frombuffer(seq[1:], dtype=np.uint8) - ord("0")
        err[end] += got != colours
        n[end] += 1
        nerr[end].append(int(g[7 + 3 * end]))
    return dict(err=err, n=n, qual=qual, nerr=[np.array(x) for x in nerr])


@pytest.fixture(scope="module")
def solid(oracle, tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("solid"))
    fasta = os.path.join(tmp, "ref.fa")
    ref = write_fasta(fasta)
    return solid_stats(oracle, oracle.RNG_DRAND48, fasta, ref, tmp), solid_stats(oracle, oracle.RNG_PHILOX, fasta, ref, tmp)


def test_solid_per_cycle_colour_error_rate(solid):
    a, b = solid
    for end, (s, e) in enumerate(((0.005, 0.04), (0.02, 0.02))):
        p = s + (e - s) / S_LEN * np.arange(S_LEN)
        for x in (a, b):
            z = (x["err"][end] - x["n"][end] * p) / np.sqrt(x["n"][end] * p * (1 - p))
            assert np.abs(z).max() < 4.8 and abs(z.sum() / math.sqrt(S_LEN)) < 4.5
        zz = [z_two(a["err"][end][i], a["n"][end], b["err"][end][i], b["n"][end]) for i in range(S_LEN)]
        assert np.abs(zz).max() < 4.8


def test_solid_quality_histograms_and_error_fields(solid):
    a, b = solid
    for end in range(2):
        ha, hb = a["qual"][end].sum(0), b["qual"][end].sum(0)                   # pooled over cycles: 41 bins
        keep = (ha + hb) >= 20
        tot_a, tot_b = ha.sum(), hb.sum()
        exp_a = (ha + hb) * tot_a / (tot_a + tot_b); exp_b = (ha + hb) * tot_b / (tot_a + tot_b)
        chi = (((ha - exp_a) ** 2 / np.maximum(exp_a, 1e-9))[keep] + ((hb - exp_b) ** 2 / np.maximum(exp_b, 1e-9))[keep]).sum()
        dof = int(keep.sum()) - 1
        assert chi < dof + 5 * math.sqrt(2 * dof), (end, chi, dof)
        ma, mb = a["nerr"][end], b["nerr"][end]
        z = (ma.mean() - mb.mean()) / math.sqrt(ma.var() / len(ma) + mb.var() / len(mb))
        assert abs(z) < 4.5


# ---- Ion Torrent 200 bp single-end ---------------------------------------------------------------------------------------
I_LEN, I_READS = 200, 8000


def ion_stats(oracle, mode, fasta, tmp):
    opts = dict(seed=31, N=I_READS, data_type=2, length=(I_LEN, 0), e=0.02, flow_order=FLOW, mut_rate=0, rand_read=0.05,
                reads_output_type=1)
    prefix = os.path.join(tmp, "ion%d" % mode)
    with oracle.Session(oracle.make_opt(**opts), fasta, prefix, mode=mode) as s:
        assert s.stats.error == 0
    lens, nerr = [], []
    for name, seq, q in records(prefix + ".bwa.read1.fastq"):
        g = NAME.match(name[:-2]).groups()                                      # "/1" suffix
        assert len(seq) == len(q)
        if g[0] == b"rand":
            assert len(seq) == I_LEN                                              # random reads carry no flow errors
            continue
        lens.append(len(seq)); nerr.append(int(g[7]))
    return np.array(lens), np.array(nerr)


def test_ion_torrent_read_lengths_and_error_counts(oracle, tmp_path):
    fasta = str(tmp_path / "ref.fa")
    write_fasta(fasta)
    (la, ea), (lb, eb) = ion_stats(oracle, oracle.RNG_DRAND48, fasta, str(tmp_path)), ion_stats(oracle, oracle.RNG_PHILOX, fasta, str(tmp_path))
    for xa, xb in ((la, lb), (ea, eb)):
        z = (xa.mean() - xb.mean()) / math.sqrt(xa.var() / len(xa) + xb.var() / len(xb))
        assert abs(z) < 4.5, (xa.mean(), xb.mean())
        assert 0.8 < xa.std() / xb.std() < 1.25
    assert ea.mean() > 3 and la.std() > 1                                       # the flow model did change the reads
    # the empty-flow insertions of the second pass (src/dwgsim.c:366-406) make reads grow on average
    assert la.mean() > I_LEN and lb.mean() > I_LEN
