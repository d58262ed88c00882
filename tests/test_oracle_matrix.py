"""Pin the oracle (drand48 backend) against the compiled reference on option sets the reference's own
tests do not cover.  Fixtures: tests/golden/ref_matrix.json (md5s written by tests/golden/make_golden.py
from oracle/_ref/dwgsim_ref).  When oracle/_ref/dwgsim_ref is present the binary is also run live."""
import gzip
import hashlib
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

FIX = json.load(open(os.path.join(HERE, "golden", "ref_matrix.json")))


def md5(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        return hashlib.md5(f.read()).hexdigest()


def test_fixture_fasta_is_reproducible(synth_fa):
    assert md5(synth_fa) == FIX["_fasta_md5"]


@pytest.mark.parametrize("case", sorted(make_golden.MATRIX))
def test_oracle_equals_reference(oracle, synth_fa, tmp_path, case):
    opts = make_golden.materialize(make_golden.MATRIX[case], str(tmp_path))
    prefix = str(tmp_path / "orc")
    with oracle.Session(oracle.make_opt(**opts), synth_fa, prefix) as s:
        assert s.stats.error == 0
    for f in make_golden.FILES:
        want = FIX[case][f]
        p = prefix + "." + f
        if want is None:
            assert not os.path.exists(p), f
        else:
            assert md5(p) == want, "%s %s" % (case, f)


@pytest.mark.parametrize("case", ["illumina_indel_heavy", "solid_2x50", "ion_paired_higherr"])
def test_live_reference_binary(oracle, synth_fa, tmp_path, case):
    """same comparison against the binary itself (skipped where oracle/_ref did not travel)"""
    if oracle.ref_binary() is None:
        pytest.skip("oracle/_ref/dwgsim_ref not built")
    got = make_golden.run_ref(synth_fa, make_golden.MATRIX[case], str(tmp_path))
    assert got == FIX[case]
