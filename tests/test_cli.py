"""The drop-in `dwgsim` host shell (dwgsim_b200/bin/dwgsim).

CPU: option surface, and .mutations.txt/.vcf byte-identical to the reference run with -C 0 / -M 2 (SURVEY.md section 0:
zero read draws from drand48), checked against the md5 fixtures written from oracle/_ref and against the pinned oracle.
GPU (-m gpu): whole runs, FASTQ bytes (after gunzip) equal to the oracle's Philox backend."""
import gzip
import hashlib
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

FIX = json.load(open(os.path.join(HERE, "golden", "ref_matrix.json")))
FLOW = make_golden.FLOW


@pytest.fixture(scope="module")
def cli():
    from dwgsim_b200 import build
    build.build()
    return build.build_cli()


def md5(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        return hashlib.md5(f.read()).hexdigest()


def run(cli, args, check=True):
    return subprocess.run([cli] + [str(a) for a in args], capture_output=True, check=check)


@pytest.mark.parametrize("case", ["mutations_only", "coverage_zero"])
def test_mutation_files_equal_reference_fixture(cli, oracle, synth_fa, tmp_path, case):
    prefix = str(tmp_path / "out")
    run(cli, oracle.opt_to_ref_argv(**make_golden.MATRIX[case]) + [synth_fa, prefix])
    for f in ("mutations.txt", "mutations.vcf"):
        assert md5(prefix + "." + f) == FIX[case][f]
    if case == "coverage_zero":            # three empty gzip members, like the reference
        for f in ("bwa.read1.fastq.gz", "bwa.read2.fastq.gz", "bfast.fastq.gz"):
            assert os.path.getsize(prefix + "." + f) == 20 and md5(prefix + "." + f) == hashlib.md5(b"").hexdigest()


MUT_CASES = {
    "defaults_ex1": (dict(seed=13), "ex1"),
    "haploid_indel_min": (dict(seed=31, mut_rate=0.02, indel_frac=0.6, indel_extend=0.5, indel_min=2, is_hap=1), "synth"),
    "long_insertions": (dict(seed=5, mut_rate=0.01, indel_frac=0.9, indel_extend=0.97, indel_min=3), "synth"),
    "high_rate": (dict(seed=8, mut_rate=0.2, indel_frac=0.3), "synth"),
    "ion_base_error_calibration": (dict(seed=28, data_type=2, length=(100, 0), e=0.02, flow_order=FLOW, use_base_error=1), "synth"),
    "skips_short_contig_paired": (dict(seed=9, dist=3000, std_dev=2000, mut_rate=0.01), "synth"),
    "regions_skip_rules": (dict(make_golden.MATRIX["regions_skip_n"], mut_rate=0.01), "synth"),
    # -m / -v / -b: mutations replayed from a file (src/mut.c:644-745); the oracle is pinned on these by test_oracle_matrix.py
    "replay_txt": (make_golden.MATRIX["replay_txt"], "synth"),
    "replay_vcf": (make_golden.MATRIX["replay_vcf"], "synth"),
    "replay_bed": (make_golden.MATRIX["replay_bed"], "synth"),
    "replay_bed_haploid": (make_golden.MATRIX["replay_bed_hap_C"], "synth"),
}


@pytest.mark.parametrize("case", sorted(MUT_CASES))
def test_mutation_files_equal_oracle_C0(cli, oracle, synth_fa, ex1_fa, tmp_path, case):
    opts, which = MUT_CASES[case]
    fasta = ex1_fa if which == "ex1" else synth_fa
    opts = make_golden.materialize(dict(opts, C=0), str(tmp_path))
    a, b = str(tmp_path / "cli"), str(tmp_path / "orc")
    run(cli, oracle.opt_to_ref_argv(**opts) + [fasta, a])
    with oracle.Session(oracle.make_opt(**opts), fasta, b) as s:
        assert s.stats.error == 0
    for f in ("mutations.txt", "mutations.vcf"):
        assert md5(a + "." + f) == md5(b + "." + f), f


@pytest.mark.parametrize("case", ["defaults_ex1", "haploid_indel_min", "long_insertions", "high_rate", "skips_short_contig_paired"])
@pytest.mark.parametrize("threads", [2, 5])
def test_parallel_mut_diref_equals_oracle_C0(cli, oracle, synth_fa, ex1_fa, tmp_path, case, threads):
    """mut_diref on several threads (hit list of the LCG iterates + one sequential step per mutation; long contigs only by
    default, forced here for every contig) must write the files of the serial path"""
    opts, which = MUT_CASES[case]
    fasta = ex1_fa if which == "ex1" else synth_fa
    opts = make_golden.materialize(dict(opts, C=0), str(tmp_path))
    a, b = str(tmp_path / "cli"), str(tmp_path / "orc")
    env = dict(os.environ, DWGSIM_DIREF_PAR_MIN="0", DWGSIM_DIREF_THREADS=str(threads))
    subprocess.run([cli] + [str(x) for x in oracle.opt_to_ref_argv(**opts) + [fasta, a]], capture_output=True, check=True, env=env)
    with oracle.Session(oracle.make_opt(**opts), fasta, b) as s:
        assert s.stats.error == 0
    for f in ("mutations.txt", "mutations.vcf"):
        assert md5(a + "." + f) == md5(b + "." + f), f


def test_option_surface(cli, synth_fa, tmp_path):
    assert run(cli, ["-h"], check=False).returncode == 1
    assert run(cli, [synth_fa], check=False).returncode == 1                       # needs <ref> <prefix>
    r = run(cli, ["-N", "5", "-C", "3", synth_fa, str(tmp_path / "x")], check=False)  # -C resets -N: fine
    assert r.returncode in (0, 1)
    r = run(cli, ["-c", "7", "-C", "0", synth_fa, str(tmp_path / "x")], check=False)
    assert r.returncode == 1 and b"-c was out of range" in r.stderr
    r = run(cli, ["-c", "2", "-C", "0", synth_fa, str(tmp_path / "x")], check=False)
    assert r.returncode == 1 and b"-f is required" in r.stderr
    r = run(cli, ["-m", "no_such_muts.txt", "-C", "0", synth_fa, str(tmp_path / "x")], check=False)
    assert r.returncode == 1
    r = run(cli, ["-m", "a.txt", "-b", "b.bed", "-C", "0", synth_fa, str(tmp_path / "x")], check=False)   # one input at most
    assert r.returncode == 1
    r = run(cli, ["-d", "abc", "-C", "0", synth_fa, str(tmp_path / "x")], check=False)
    assert r.returncode == 1 and b"is not a number" in r.stderr


def test_replayed_txt_reproduces_itself(cli, synth_fa, tmp_path):
    """-m with the reference's own .mutations.txt gives that file back (the purpose of the option), and rejected
    inputs end like the reference's parsers do (src/mut_txt.c:58-69, src/mut_bed.c:57-80)"""
    src = os.path.join(HERE, "golden", "replay_muts.txt")
    run(cli, ["-M", "2", "-z", "3", "-m", src, synth_fa, str(tmp_path / "x")])     # -M 2 like the run that wrote it (no contig skipped)
    assert open(str(tmp_path / "x.mutations.txt")).read() == open(src).read()
    bad = tmp_path / "bad.txt"
    bad.write_text("chrA\t500\tA\tC\t3\nchrA\t100\tA\tG\t3\nchrZ\t5\tA\tG\t3\n")
    r = run(cli, ["-C", "0", "-m", str(bad), synth_fa, str(tmp_path / "y")], check=False)
    assert r.returncode == 1 and b"mutation contig not found or out of order [chrZ]" in r.stderr
    bad.write_text("chrA\t500\tA\tC\t1\n")                       # heterozygous substitutions must be IUPAC codes
    r = run(cli, ["-C", "0", "-m", str(bad), synth_fa, str(tmp_path / "y")], check=False)
    assert r.returncode == 1 and b"heterozygous bases must be in IUPAC form" in r.stderr
    (tmp_path / "empty.txt").write_text("")                        # the reference cannot hold zero records (realloc(p, 0))
    r = run(cli, ["-C", "0", "-m", str(tmp_path / "empty.txt"), synth_fa, str(tmp_path / "y")], check=False)
    assert r.returncode == 1 and b"memory allocation failed in muts_txt_init" in r.stderr
    same = tmp_path / "same.bed"                                     # a substitution that keeps the base: the reference aborts in mut_debug
    seq = "".join(l.strip() for l in open(synth_fa).read().split(">")[1].split("\n")[1:])
    same.write_text("chrA\t100\t101\t%s\tsnp\n" % seq[100].upper())
    r = run(cli, ["-C", "0", "-H", "-b", str(same), synth_fa, str(tmp_path / "y")], check=False)
    assert r.returncode == 1 and b"inconsistent substitution at chrA:101" in r.stderr
    bed = tmp_path / "bad.bed"
    bed.write_text("chrA\t10\t40\t*\tins\n")
    r = run(cli, ["-C", "0", "-b", str(bed), synth_fa, str(tmp_path / "y")], check=False)
    assert r.returncode == 1 and b"exceeded the maximum supported length of 26" in r.stderr


def test_prologue_variants_write_the_same_mutation_files(cli, synth_fa, tmp_path):
    """the producer thread (DWGSIM_PIPELINE=1) and the event-list driven left-justification / writers against the inline,
    every-base forms (DWGSIM_PIPELINE=0, DWGSIM_FULL_SCAN=1): same .mutations.txt/.vcf"""
    args = ["-C", "0", "-z", "77", "-r", "0.02", "-R", "0.5", "-X", "0.8", "-I", "2"]
    outs = []
    for k, env in enumerate(({"DWGSIM_PIPELINE": "1"}, {"DWGSIM_PIPELINE": "0"}, {"DWGSIM_PIPELINE": "0", "DWGSIM_FULL_SCAN": "1"})):
        prefix = str(tmp_path / ("v%d" % k))
        subprocess.run([cli] + args + [synth_fa, prefix], capture_output=True, check=True, env=dict(os.environ, **env))
        outs.append((md5(prefix + ".mutations.txt"), md5(prefix + ".mutations.vcf")))
    assert outs[0] == outs[1] == outs[2]
    assert os.path.getsize(str(tmp_path / "v0.mutations.txt")) > 10000


def test_fai_census(cli, oracle, synth_fa, tmp_path):
    """with <ref.fa>.fai next to the FASTA the census pass reads names and lengths from it (src/dwgsim.c:467-478)"""
    fa = str(tmp_path / "s.fa")
    data = open(synth_fa, "rb").read()
    open(fa, "wb").write(data)
    recs, name, n = [], None, 0
    for line in data.split(b"\n"):
        if line.startswith(b">"):
            if name is not None:
                recs.append((name, n))
            name, n = line[1:].split()[0].decode(), 0
        else:
            n += len(line)
    recs.append((name, n))
    open(fa + ".fai", "w").write("".join("%s\t%d\t0\t60\t61\n" % r for r in recs))
    opts = dict(seed=6, C=0, mut_rate=0.02, indel_frac=0.4)
    a, b = str(tmp_path / "cli"), str(tmp_path / "orc")
    r = run(cli, oracle.opt_to_ref_argv(**opts) + [fa, a])
    assert b"[dwgsim_core] hp length: 8000" in r.stderr
    with oracle.Session(oracle.make_opt(**opts), fa, b) as s:
        assert s.stats.error == 0
    for f in ("mutations.txt", "mutations.vcf"):
        assert md5(a + "." + f) == md5(b + "." + f), f


def test_reference_stderr_lines(cli, synth_fa, tmp_path):
    r = run(cli, ["-C", "0", "-z", "1", synth_fa, str(tmp_path / "x")])
    err = r.stderr.decode()
    assert "[dwgsim_core] chrA length: 30000" in err
    assert "[dwgsim_core] 4 sequences, total length: 50400" in err
    assert "#3 skip sequence 'tiny' as it is shorter than 650.000000!" in err
    assert "[dwgsim_core] Complete!" in err


def test_regions_skip_messages(cli, oracle, synth_fa, tmp_path):
    """-x: the reference's skip rules #0 (no region on the contig) and #1 (regions are > 95 % non-ACGT), src/dwgsim.c:547-580"""
    opts = make_golden.materialize(dict(make_golden.MATRIX["regions_skip_n"], C=0), str(tmp_path))
    r = run(cli, oracle.opt_to_ref_argv(**opts) + [synth_fa, str(tmp_path / "x")])
    err = r.stderr.decode()
    assert "#0 skip sequence 'tiny' as it is not in the targeted region" in err
    assert "#1 skip sequence 'chrA' as 581 out of 580 bases are non-ACGT" in err
    bad = tmp_path / "bad.bed"
    bad.write_text("chrB\t500\t900\nchrB\t100\t300\n")
    r = run(cli, ["-C", "0", "-x", str(bad), synth_fa, str(tmp_path / "y")], check=False)
    assert r.returncode == 1 and b"the input was not sorted" in r.stderr
    r = run(cli, ["-C", "0", "-a", "-x", str(bad), synth_fa, str(tmp_path / "y")], check=False)
    assert r.returncode == 1 and b"cannot use a regions BED file" in r.stderr


GPU_CASES = {
    "illumina_gz": (dict(seed=7, N=4000, length=(100, 100), mut_rate=0.01, indel_frac=0.3), []),
    "illumina_plain_config1": (dict(seed=13, N=10000, length=(100, 100), data_type=0), ["--uncompressed"]),
    "solid_gz": (dict(seed=22, N=3000, data_type=1, length=(50, 50), mut_rate=0.02, indel_frac=0.5), ["--batch", "1000"]),
    "illumina_host_gzip": (dict(seed=8, N=3000, length=(100, 100)), ["--host-gzip"]),
    "ion_plain": (dict(seed=25, N=1200, data_type=2, length=(200, 0), e=0.02, flow_order=FLOW), ["--uncompressed"]),
    "regions_N": (make_golden.MATRIX["regions_N"], ["--uncompressed"]),
    "regions_skip_n_gz": (make_golden.MATRIX["regions_skip_n"], []),
    "replay_bed_gz": (make_golden.MATRIX["replay_bed"], []),
    "replay_vcf_plain": (make_golden.MATRIX["replay_vcf"], ["--uncompressed"]),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(GPU_CASES))
def test_whole_run_equals_oracle(cli, oracle, synth_fa, ex1_fa, tmp_path, case):
    opts, extra = GPU_CASES[case]
    opts = make_golden.materialize(opts, str(tmp_path))
    fasta = ex1_fa if case == "illumina_plain_config1" else synth_fa
    a, b = str(tmp_path / "cli"), str(tmp_path / "orc")
    run(cli, oracle.opt_to_ref_argv(**opts) + extra + [fasta, a])
    with oracle.Session(oracle.make_opt(**opts), fasta, b, mode=oracle.RNG_PHILOX) as s:
        assert s.stats.error == 0
    ext = "" if "--uncompressed" in extra else ".gz"
    for f in ("bwa.read1.fastq", "bwa.read2.fastq", "bfast.fastq"):
        assert md5(a + "." + f + ext) == md5(b + "." + f), f
    for f in ("mutations.txt", "mutations.vcf"):
        assert md5(a + "." + f) == md5(b + "." + f), f
