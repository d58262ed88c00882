"""-x (targeted regions, src/dwgsim.c:539-581,677-713 + src/regions_bed.c): properties of the oracle's two backends.

Byte parity of the drand48 backend with the compiled reference is in test_oracle_matrix.py (regions_* cases); here the
Philox backend (the kernels' specification) is checked against the drand48 backend statistically, and both against the
reference's containment rule."""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

REGIONS = {"chrA": [(1000, 12000), (15000, 22000)], "chrB": [(2000, 10000)], "hp": [(100, 7000)]}
NAME = re.compile(rb"^@(\w+)_(\d+)_(\d+)_([01])_([01])_([01])_([01])_")


def fragments(path):
    """(contig, leftmost 0-based start, rightmost 0-based end) of every genomic pair in a bwa read-1 file"""
    out = []
    with open(path, "rb") as f:
        for i, line in enumerate(f):
            if i % 4:
                continue
            m = NAME.match(line)
            if m is None or m.group(1) == b"rand":
                continue
            out.append((m.group(1).decode(), int(m.group(2)), int(m.group(3))))
    return out


def run(oracle, mode, fasta, tmp, seed):
    opts = dict(seed=seed, C=40, length=(100, 100), rand_read=0.02, mut_rate=0,
                regions=[(c, a, b) for c, rs in REGIONS.items() for a, b in rs])
    opts = make_golden.materialize(opts, tmp)
    prefix = os.path.join(tmp, "m%d" % mode)
    with oracle.Session(oracle.make_opt(**opts), fasta, prefix, mode=mode) as s:
        assert s.stats.error == 0
        failed = s.stats.n_failed_attempts
    return fragments(prefix + ".bwa.read1.fastq"), failed


def test_regions_containment_and_backend_agreement(oracle, synth_fa, tmp_path):
    stats = {}
    for mode in (oracle.RNG_DRAND48, oracle.RNG_PHILOX):
        fr, failed = run(oracle, mode, synth_fa, str(tmp_path), seed=41 + mode)
        assert len(fr) > 3000
        per_contig = {}
        for contig, p1, p2 in fr:
            lo, hi = min(p1, p2) - 1, max(p1, p2) - 1 + 100     # 0-based fragment [lo, hi)
            assert any(a <= lo and hi <= b + 1 for a, b in REGIONS[contig]), (contig, lo, hi)   # regions_bed_query is end-inclusive
            per_contig.setdefault(contig, []).append(lo)
        stats[mode] = {c: (len(v), float(np.mean(v))) for c, v in per_contig.items()}
        if mode == oracle.RNG_PHILOX:
            assert failed > 0                                   # fragments crossing a region edge are redrawn as new attempts
    a, b = stats[oracle.RNG_DRAND48], stats[oracle.RNG_PHILOX]
    assert set(a) == set(b) == set(REGIONS)
    for c in a:
        span = sum(y - x for x, y in REGIONS[c])
        # the pair budget is deterministic; only the number of random pairs (2 %) differs between the backends
        assert abs(a[c][0] - b[c][0]) < 6 * np.sqrt(a[c][0] * 0.02 + 1)
        assert abs(a[c][1] - b[c][1]) < 6 * span / np.sqrt(12 * min(a[c][0], b[c][0]))   # mean position, 6 sigma of a uniform
