"""-m gpu: INTEGRATION.md as a tested artefact.  oracle/_ref/dwgsim_ref_gpu is the reference itself (its option parser, FASTA
reader, mut_diref, mut_print, gzFile writers) with the read-pair loop of dwgsim_core replaced by the binding in
integration/dwgsim_b200_binding.c (built by `make -C oracle ref_gpu` from /root/reference + oracle/patch_reference.py; the
binary travels to the GPU box with the snapshot).  Its FASTQ files must be the oracle's (Philox backend), its mutation
files the stock reference's `-C 0` files (the loop no longer draws from drand48)."""
import gzip
import hashlib
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import gpu_harness as gh  # noqa: E402
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu


def md5_of(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        return hashlib.md5(f.read()).hexdigest()


@pytest.fixture(scope="module")
def ref_gpu(oracle):
    p = oracle.ref_gpu_binary()
    if p is None:
        pytest.skip("oracle/_ref/dwgsim_ref_gpu was not built (needs /root/reference at build time)")
    return p


CASES = {
    "golden_config": (dict(seed=13, N=10000), "ex1"),                     # the reference's own test (testdata/test.sh:18)
    "illumina_2x150_slope": (make_golden.MATRIX["illumina_2x150_slope_C"], "synth"),
    "solid_2x50": (make_golden.MATRIX["solid_2x50"], "synth"),
    "ion_400_se": (make_golden.MATRIX["ion_400_se"], "synth"),
    "regions": (make_golden.MATRIX["regions_N"], "synth"),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_reference_with_the_binding_writes_the_oracles_files(oracle, ref_gpu, synth_fa, ex1_fa, tmp_path, case):
    opts, which = CASES[case]
    fasta = ex1_fa if which == "ex1" else synth_fa
    opts = make_golden.materialize(dict(opts), str(tmp_path))
    a, b = str(tmp_path / "ref_gpu"), str(tmp_path / "orc")
    r = subprocess.run([ref_gpu] + [str(x) for x in oracle.opt_to_ref_argv(**opts)] + [fasta, a], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    with oracle.Session(oracle.make_opt(**opts), fasta, b, mode=oracle.RNG_PHILOX) as s:
        assert s.stats.error == 0
    for f in gh.FILE_NAMES:
        if os.path.exists(b + "." + f):
            assert md5_of(a + "." + f + ".gz") == md5_of(b + "." + f), f
    for f in ("mutations.txt", "mutations.vcf"):
        assert md5_of(a + "." + f) == md5_of(b + "." + f), f


def test_mutation_files_equal_the_stock_reference_C0(oracle, ref_gpu, synth_fa, tmp_path):
    """against the UNMODIFIED reference binary: same options with -C 0 (zero read draws), multi-contig FASTA with indels"""
    stock = oracle.ref_binary()
    if stock is None:
        pytest.skip("oracle/_ref/dwgsim_ref not present")
    opts = dict(seed=3, N=4000, length=(100, 100), mut_rate=0.02, indel_frac=0.5, indel_extend=0.7)
    a, b = str(tmp_path / "ref_gpu"), str(tmp_path / "stock")
    r = subprocess.run([ref_gpu] + [str(x) for x in oracle.opt_to_ref_argv(**opts)] + [synth_fa, a], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    r = subprocess.run([stock] + [str(x) for x in oracle.opt_to_ref_argv(**dict(opts, C=0))] + [synth_fa, b], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    for f in ("mutations.txt", "mutations.vcf"):
        assert md5_of(a + "." + f) == md5_of(b + "." + f), f
