"""-m gpu: the CUDA path through the C ABI must equal the oracle (Philox backend) byte for byte."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import gpu_harness as gh  # noqa: E402
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu

CASES = {k: v for k, v in make_golden.MATRIX.items() if v.get("output_type", 0) == 0 and k != "coverage_zero"}


def check(oracle, opts, fasta, tmp_path, **kw):
    opts = make_golden.materialize(opts, str(tmp_path))
    sess, want = gh.oracle_expected(oracle, opts, fasta, str(tmp_path / "orc"))
    try:
        assert sess.stats.error == 0
        got, stats = gh.gpu_actual(sess, opts, orc_opt=sess.opt, **kw)
        for i, name in enumerate(gh.FILE_NAMES):
            assert got[i] == want[i], "%s: %s" % (name, gh.first_diff(want[i], got[i]))
        assert sum(s.n_pairs for s in stats) == sess.stats.n_pairs_total
        assert sum(s.n_random for s in stats) == sess.stats.n_random
        assert sum(s.n_failed_attempts for s in stats) == sess.stats.n_failed_attempts
        return stats
    finally:
        sess.close()


@pytest.mark.parametrize("case", sorted(CASES))
def test_matrix_case_bit_exact(oracle, synth_fa, tmp_path, case):
    check(oracle, CASES[case], synth_fa, tmp_path)


def test_reference_golden_config_bit_exact(oracle, ex1_fa, tmp_path):
    """the reference's own test configuration (testdata/test.sh:18) with Philox draws"""
    check(oracle, dict(seed=13, N=10000), ex1_fa, tmp_path)


def test_config1_2x100(oracle, ex1_fa, tmp_path):
    """BASELINE.json configs[0]: ex1.fa 2x100bp -N 10000 Illumina"""
    check(oracle, dict(seed=13, N=10000, length=(100, 100), data_type=0), ex1_fa, tmp_path)


def test_small_batches_and_per_contig_runs_give_same_bytes(oracle, synth_fa, tmp_path):
    """output must not depend on the batch size or on calling run() once per contig (pair-indexed Philox)"""
    opts = dict(seed=3, N=5000, length=(100, 100), mut_rate=0.02, indel_frac=0.5, indel_extend=0.7)
    check(oracle, opts, synth_fa, tmp_path, batch=777)
    check(oracle, opts, synth_fa, tmp_path, batch=1024, per_contig_runs=True)


def test_ion_torrent_config5_shape(oracle, synth_fa, tmp_path):
    """BASELINE.json configs[4]: Ion Torrent 400 bp single-end, 32-flow order, -e 0.01"""
    check(oracle, dict(seed=5, C=4, data_type=2, length=(400, 0), e=0.01, flow_order=make_golden.FLOW), synth_fa, tmp_path)


def test_ion_torrent_heavy_errors_many_shifts(oracle, synth_fa, tmp_path):
    check(oracle, dict(seed=6, N=3000, data_type=2, length=(150, 80), e=0.12, E=0.2, flow_order="TACGTACGTCTGAGCATCGATCGATGTACAGC",
                       dist=300, std_dev=30), synth_fa, tmp_path)


def test_ion_torrent_warp_kernel_fallback(oracle, synth_fa, tmp_path, monkeypatch):
    """reads too long for the thread-per-pair rows use the warp-per-pair kernel; force it on a short case"""
    monkeypatch.setenv("DWGSIM_ION_KERNEL", "warp")
    check(oracle, dict(seed=26, N=1500, data_type=2, length=(200, 100), e=0.05, E=0.03, flow_order=make_golden.FLOW,
                       mut_rate=0.01, indel_frac=0.4), synth_fa, tmp_path)


def test_long_paired_reads_shrink_the_format_cta(oracle, synth_fa, tmp_path):
    """2 x 1,200-base Ion Torrent reads: a pair's records need ~20 KB of staging per warp, so the format kernel runs with
    fewer warps per CTA (and the simulate kernel falls back to warp-per-pair)"""
    check(oracle, dict(seed=28, N=200, data_type=2, length=(1200, 1200), dist=3000, std_dev=100, e=0.01, E=0.01,
                       flow_order=make_golden.FLOW), synth_fa, tmp_path)


def test_ion_torrent_long_reads_use_fallback(oracle, synth_fa, tmp_path):
    check(oracle, dict(seed=27, N=300, data_type=2, length=(1500, 0), e=0.01, flow_order=make_golden.FLOW), synth_fa, tmp_path)


@pytest.mark.parametrize("case", ["illumina_2x150_slope_C", "solid_2x50", "ion_400_se", "illumina_single_end", "illumina_q0_bfast"])
def test_device_gzip_members_decode_to_the_same_bytes(oracle, synth_fa, tmp_path, case):
    """sink receives gzip members written on the device; zlib must decode them to the oracle's FASTQ bytes"""
    stats = check(oracle, CASES[case], synth_fa, tmp_path, compression=1, batch=2500)
    assert sum(sum(s.bytes) for s in stats) < 0.7 * sum(sum(s.raw_bytes) for s in stats)


def test_device_gzip_large_batch(oracle, synth_fa, tmp_path):
    """several 64 KiB members per stream and batch, ragged last member"""
    check(oracle, dict(seed=2, N=60000, length=(150, 150), e="0.001-0.01", E="0.001-0.01"), synth_fa, tmp_path, compression=1)


def test_derived_tables_equal_oracle(oracle):
    """the 32-bit threshold tables the kernels sample from == the oracle's own derivation"""
    from dwgsim_b200 import DwgsimGpu, params_from_options
    opts = dict(seed=1, C=1, length=(150, 120), e="0.001-0.01", E="0.0-0.3", std_dev=37.5, dist=420, quality_std=3.3,
                rand_read=0.123, mut_freq=0.37)
    o = oracle.make_opt(**opts)
    t = oracle.lib().orc_tables_build(o).contents
    with DwgsimGpu(params_from_options(**{k: v for k, v in opts.items() if k in gh.GPU_KEYS})) as gpu:
        g = gpu.tables()
        assert (g.thr_genomic, g.thr_hap0) == (t.thr_genomic, t.thr_hap0)
        assert (g.isize_lo, g.isize_n, g.qdelta_lo, g.qdelta_n) == (t.isize_lo, t.isize_n, t.qdelta_lo, t.qdelta_n)
        assert [g.isize_cdf[i] for i in range(g.isize_n)] == [t.isize_cdf[i] for i in range(t.isize_n)]
        assert [g.qdelta_cdf[i] for i in range(g.qdelta_n)] == [t.qdelta_cdf[i] for i in range(t.qdelta_n)]
        for e in range(2):
            n = opts["length"][e]
            assert [g.err_gap[e][i] for i in range(n)] == [t.err_gap[e][i] for i in range(n)]
            assert [g.err_acc[e][i] for i in range(n)] == [t.err_acc[e][i] for i in range(n)]
            assert [g.qbase[e][i] for i in range(n)] == [t.qbase[e][i] for i in range(n)]


def test_wide_insert_size_table(oracle, synth_fa, tmp_path):
    """-s 5000: the insert-size table has more steps than the 16-bit guide can index (searched whole on the device)"""
    check(oracle, dict(seed=21, N=3000, dist=6000, std_dev=5000), synth_fa, tmp_path)


@pytest.fixture(scope="module")
def growing_names_fa(tmp_path_factory):
    """three contigs whose names grow from 2 to 70 characters (hg38-style alt contig names after 'c1')"""
    import numpy as np
    rng = np.random.default_rng(99)
    p = str(tmp_path_factory.mktemp("names") / "names.fa")
    with open(p, "w") as f:
        for name in ("c1", "chr1_KI270706v1_random", "chrUn_" + "x" * 58 + "_alt_v2"):
            f.write(">%s\n" % name)
            b = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 6000)]).decode()
            for i in range(0, len(b), 60):
                f.write(b[i:i + 60] + "\n")
    return p


@pytest.mark.parametrize("data_type", [0, 1])
def test_contig_names_grow_between_runs(oracle, growing_names_fa, tmp_path, data_type):
    """run() once per contig with ever longer contig names: the name, stream and pinned buffers must grow with them"""
    opts = dict(seed=8, N=6000, data_type=data_type, length=(50, 50) if data_type else (100, 100))
    check(oracle, opts, growing_names_fa, tmp_path, per_contig_runs=True, batch=4096)


def test_flow_gap_tables_equal_oracle(oracle):
    """Ion Torrent: the geometric gap table of the per-flow error coin == the oracle's"""
    from dwgsim_b200 import DwgsimGpu, params_from_options
    opts = dict(seed=1, C=1, data_type=2, length=(200, 100), e=0.013, E=0.04, flow_order=make_golden.FLOW)
    o = oracle.make_opt(**opts)
    t = oracle.lib().orc_tables_build(o).contents
    with DwgsimGpu(params_from_options(**{k: v for k, v in opts.items() if k in gh.GPU_KEYS})) as gpu:
        g = gpu.tables()
        for e in range(2):
            assert g.flow_gap_n[e] == oracle.FLOW_GAP_N
            assert [g.flow_gap[e][i] for i in range(oracle.FLOW_GAP_N)] == [t.flow_gap[e][i] for i in range(oracle.FLOW_GAP_N)]


def test_pack_then_add_packed_and_the_file_sink(oracle, synth_fa, tmp_path):
    """the two-call form of add_contig (pack on one thread, queue later) and the library's file sink (one writer thread
    per file, positional writes) must give the oracle's bytes"""
    from dwgsim_b200 import DwgsimGpu, params_from_options
    from concurrent.futures import ThreadPoolExecutor
    opts = dict(seed=4, N=5000, length=(100, 100), mut_rate=0.01)
    sess, want = gh.oracle_expected(oracle, opts, synth_fa, str(tmp_path / "orc"))
    try:
        names = [str(tmp_path / ("out." + f)) for f in gh.FILE_NAMES]
        fds = [os.open(p, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644) for p in names]
        offs = [0, 0, 0]
        with DwgsimGpu(params_from_options(**{k: v for k, v in opts.items() if k in gh.GPU_KEYS})) as gpu, ThreadPoolExecutor(1) as pool:
            gpu.set_batch(900, 3)
            gpu.set_host_threads(2)
            cs = [sess.contig(k) for k in range(sess.n_contigs)]

            def pack(c):
                return gpu.pack_contig(c["contig_i"], c["name"], c["seq"], c["len"], c["hap"][0], c["hap"][1], c["ins"][0],
                                       c["n_ins"][0], c["ins"][1], c["n_ins"][1], c["n_pairs"])
            nxt = pool.submit(pack, cs[0])
            for k in range(len(cs)):                       # contig k+1 is packed while contig k runs
                gpu.add_packed(nxt.result())
                if k + 1 < len(cs):
                    nxt = pool.submit(pack, cs[k + 1])
                st, offs = gpu.run_to_files(fds, offs)
        for fd in fds:
            os.close(fd)
        for i, p in enumerate(names):
            got = open(p, "rb").read()
            assert got == want[i], "%s: %s" % (gh.FILE_NAMES[i], gh.first_diff(want[i], got))
    finally:
        sess.close()
