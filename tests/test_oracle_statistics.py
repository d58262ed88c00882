"""Link 2 of the parity chain: the oracle's Philox backend samples the same distributions as
its drand48 backend (= the reference, byte for byte).  Every statistic the north star names is compared between the
two backends and against its closed form within sampling error: per-cycle error rate, per-cycle quality histogram,
insert size, strand, haplotype, random-pair fraction, position uniformity, name-field count distributions."""
import math
import os
import re

import numpy as np
import pytest

N_PAIRS = 30000
LEN = 100
COMP = bytes.maketrans(b"ACGTN", b"TGCAN")
NAME = re.compile(rb"^@(.+)_(\d+)_(\d+)_([01])_([01])_([01])_([01])_(\d+):(\d+):(\d+)_(\d+):(\d+):(\d+)_([0-9a-f]+)/([12])$")


def write_fasta(path, n=200000, seed=5):
    rng = np.random.default_rng(seed)
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)]
    with open(path, "wb") as f:
        f.write(b">c1\n")
        f.write(s.tobytes())
        f.write(b"\n")
    return s.tobytes()


def parse(path):
    recs = []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(0, len(lines) - 1, 4):
        m = NAME.match(lines[i])
        assert m, lines[i]
        recs.append((m.groups(), lines[i + 1], lines[i + 3]))
    return recs


def collect(oracle, mode, fasta, ref, tmp, opts):
    prefix = os.path.join(tmp, "m%d" % mode)
    with oracle.Session(oracle.make_opt(**opts), fasta, prefix, mode=mode) as s:
        assert s.stats.error == 0
        n_random = s.stats.n_random
    out = dict(n_random=n_random)
    err = np.zeros((2, LEN))
    nreads = np.zeros(2)
    qual = np.zeros((2, LEN, 41))
    isize, strand0, pos = [], [], []
    nerr_field = [[], []]
    for end, fn in enumerate(("bwa.read1.fastq", "bwa.read2.fastq")):
        for (g, seq, q) in parse(prefix + "." + fn):
            qa = np.frombuffer(q, dtype=np.uint8) - 33
            qual[end, np.arange(LEN), qa] += 1
            if g[0] == b"rand":
                continue
            p = [int(g[1]), int(g[2])]
            st = [int(g[3]), int(g[4])]
            truth = ref[p[end] - 1:p[end] - 1 + LEN]
            if st[end]:
                truth = truth.translate(COMP)[::-1]
            mism = np.frombuffer(seq, dtype=np.uint8) != np.frombuffer(truth, dtype=np.uint8)
            err[end] += mism
            nreads[end] += 1
            nerr_field[end].append(int(g[7 + 3 * end]))
            assert int(g[7 + 3 * end]) >= mism.sum()          # a substitution may restore nothing: errors always change the base
            if end == 0:
                isize.append(abs(p[1] - p[0]) + LEN)
                strand0.append(st[0])
                pos.append(min(p))
    out.update(err=err, nreads=nreads, qual=qual, isize=np.array(isize), strand0=np.array(strand0), pos=np.array(pos),
               nerr=[np.array(x) for x in nerr_field])
    return out


@pytest.fixture(scope="module")
def both(oracle, tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("stat"))
    fasta = os.path.join(tmp, "ref.fa")
    ref = write_fasta(fasta)
    opts = dict(seed=11, N=N_PAIRS, length=(LEN, LEN), e="0.002-0.03", E="0.01-0.05", mut_rate=0, rand_read=0.07,
                reads_output_type=1)
    return (collect(oracle, oracle.RNG_DRAND48, fasta, ref, tmp, opts),
            collect(oracle, oracle.RNG_PHILOX, fasta, ref, tmp, opts), opts)


def z_two(k1, n1, k2, n2):
    p = (k1 + k2) / max(n1 + n2, 1)
    v = p * (1 - p) * (1 / max(n1, 1) + 1 / max(n2, 1))
    return 0.0 if v <= 0 else (k1 / n1 - k2 / n2) / math.sqrt(v)


def test_random_pair_fraction(both):
    a, b, o = both
    for x in (a, b):
        z = (x["n_random"] - N_PAIRS * o["rand_read"]) / math.sqrt(N_PAIRS * o["rand_read"] * (1 - o["rand_read"]))
        assert abs(z) < 4.5
    assert abs(z_two(a["n_random"], N_PAIRS, b["n_random"], N_PAIRS)) < 4.5


def test_per_cycle_error_rate(both):
    a, b, o = both
    for end, (s, e) in enumerate(((0.002, 0.03), (0.01, 0.05))):
        p = s + (e - s) / LEN * np.arange(LEN)                # src/dwgsim_opt.c:459-460, src/dwgsim.c:237
        for x in (a, b):
            n = x["nreads"][end]
            z = (x["err"][end] - n * p) / np.sqrt(n * p * (1 - p))
            assert np.abs(z).max() < 4.8, (end, np.abs(z).max())
            assert abs(z.sum() / math.sqrt(LEN)) < 4.5           # no systematic bias over the cycles
        zz = [z_two(a["err"][end][i], a["nreads"][end], b["err"][end][i], b["nreads"][end]) for i in range(LEN)]
        assert np.abs(zz).max() < 4.8


def test_per_cycle_quality_histogram(both):
    """closed form of SURVEY.md App. A.11 (truncation toward zero, clamp 0..40) vs both backends"""
    from scipy.stats import norm, chi2
    a, b, o = both
    sd = 2.0
    for end, (s, e) in enumerate(((0.002, 0.03), (0.01, 0.05))):
        for x in (a, b):
            stat, dof = 0.0, 0
            n = x["qual"][end][0].sum()
            for i in range(0, LEN, 7):
                p = s + (e - s) / LEN * i
                qb = int(-10.0 * math.log(p) / math.log(10.0) + 0.499)
                prob = np.zeros(41)
                for d in range(-20, 21):
                    if d == 0:
                        pr = norm.cdf(0.5 / sd) - norm.cdf(-1.5 / sd)
                    elif d > 0:
                        pr = norm.cdf((d + 0.5) / sd) - norm.cdf((d - 0.5) / sd)
                    else:
                        pr = norm.cdf((d - 0.5) / sd) - norm.cdf((d - 1.5) / sd)
                    prob[min(max(qb + d, 0), 40)] += pr
                keep = prob * n > 5
                obs = x["qual"][end][i]
                stat += (((obs - prob * n) ** 2)[keep] / (prob * n)[keep]).sum()
                dof += keep.sum() - 1
            assert chi2.sf(stat, dof) > 1e-6, (end, stat, dof)
        # backend against backend, all cycles pooled per quality value
        ha, hb = a["qual"][end].sum(0), b["qual"][end].sum(0)
        keep = (ha + hb) > 20
        tot_a, tot_b = ha.sum(), hb.sum()
        exp_a = (ha + hb) * tot_a / (tot_a + tot_b)
        exp_b = (ha + hb) * tot_b / (tot_a + tot_b)
        stat = (((ha - exp_a) ** 2 / exp_a)[keep] + ((hb - exp_b) ** 2 / exp_b)[keep]).sum()
        assert chi2.sf(stat, keep.sum() - 1) > 1e-6


def test_insert_size_strand_position(both):
    from scipy.stats import ks_2samp
    a, b, o = both
    for x in (a, b):
        n = len(x["isize"])
        assert abs(x["isize"].mean() - 500.0) < 5 * 50 / math.sqrt(n)          # -d 500 -s 50, outer distance
        assert abs(x["isize"].std() - 50.0) < 1.5
        assert abs(x["strand0"].mean() - 0.5) < 4.5 * 0.5 / math.sqrt(n)
        # positions uniform over [0, l-d]: compare decile counts
        h, _ = np.histogram(x["pos"], bins=10, range=(0, 200000 - 500))
        assert np.abs((h - n / 10) / math.sqrt(n / 10 * 0.9)).max() < 4.8
    assert ks_2samp(a["isize"], b["isize"]).pvalue > 1e-5
    assert ks_2samp(a["pos"], b["pos"]).pvalue > 1e-5


def test_name_error_count_distribution(both):
    a, b, o = both
    for end in (0, 1):
        ma, mb = a["nerr"][end].mean(), b["nerr"][end].mean()
        sa = a["nerr"][end].std() / math.sqrt(len(a["nerr"][end]))
        assert abs(ma - mb) < 5 * math.sqrt(2) * sa


def test_haplotype_and_mutation_fields(oracle, tmp_path):
    """with het-only-ish mutations the sub/indel name fields depend on the haplotype draw (-F): compare backends"""
    fasta = str(tmp_path / "ref.fa")
    write_fasta(fasta, n=100000, seed=9)
    opts = dict(seed=4, N=12000, length=(LEN, LEN), mut_rate=0.01, indel_frac=0.3, mut_freq=0.3, reads_output_type=1)
    tot = []
    for mode in (oracle.RNG_DRAND48, oracle.RNG_PHILOX):
        prefix = str(tmp_path / ("h%d" % mode))
        with oracle.Session(oracle.make_opt(**opts), fasta, prefix, mode=mode) as s:
            assert s.stats.error == 0
        sub = indel = n = 0
        for (g, _, _) in parse(prefix + ".bwa.read1.fastq"):
            if g[0] == b"rand":
                continue
            sub += int(g[8]); indel += int(g[9]); n += 1
        tot.append((sub / n, indel / n, n))
    # both backends mutate with the same drand48 stream only until the first read draw, so the genomes differ:
    # compare the per-read means loosely (they estimate the same rate x read length)
    assert abs(tot[0][0] - tot[1][0]) < 0.25 * max(tot[0][0], tot[1][0])
    assert abs(tot[0][1] - tot[1][1]) < 0.35 * max(tot[0][1], tot[1][1])
