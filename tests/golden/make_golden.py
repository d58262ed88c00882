#!/usr/bin/env python
"""Generate tests/golden/ref_matrix.json: md5s of the five output files of the UNMODIFIED reference
(oracle/_ref/dwgsim_ref, compiled from /root/reference by oracle/Makefile) over a matrix of option
sets the reference's own test-suite does not cover (SOLiD, Ion Torrent, single-end, sloped error
rates, indel-heavy genomes, N runs ...).  tests/test_oracle_matrix.py replays the same matrix through
the oracle (drand48 backend) and requires identical md5s, which pins the oracle beyond the five
golden files of the reference's testdata/.

Run from the repo root in the build container (needs /root/reference):  python tests/golden/make_golden.py
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

FILES = ["bwa.read1.fastq", "bwa.read2.fastq", "bfast.fastq", "mutations.txt", "mutations.vcf"]
FLOW = "TACGTACGTCTGAGCATCGATCGATGTACAGC"


def synth_fasta(path, seed=20261017, with_fai=False):
    """4 contigs: 30 kb with two N runs, 12 kb with lowercase + IUPAC, 400 bp (skipped: too short),
    8 kb homopolymer-rich (exercises left-justification)"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rnd(n):
        return acgt[rng.integers(0, 4, n)].copy()

    c1 = rnd(30000)
    c1[5000:5600] = ord("N")
    c1[20000:20040] = ord("N")
    c1[:50] = ord("N")
    c2 = rnd(12000)
    c2[3000:3300] += 32  # lowercase
    c2[7000] = ord("R")
    c2[7100] = ord("Y")
    c3 = rnd(400)
    runs = rng.integers(1, 9, 4000)
    c4 = np.repeat(acgt[rng.integers(0, 4, 4000)], runs)[:8000]
    with open(path, "w") as f:
        for name, s in (("chrA", c1), ("chrB desc text", c2), ("tiny", c3), ("hp", c4)):
            f.write(">%s\n" % name)
            b = s.tobytes().decode()
            for i in range(0, len(b), 60):
                f.write(b[i:i + 60] + "\n")
    if with_fai:
        raise NotImplementedError
    return path


# option sets by oracle keyword (pyoracle.make_opt / opt_to_ref_argv)
MATRIX = {
    "illumina_default_N": dict(seed=7, N=3000),
    "illumina_2x150_slope_C": dict(seed=1, C=3, length=(150, 150), e="0.001-0.01", E="0.001-0.01"),
    "illumina_indel_heavy": dict(seed=3, N=4000, length=(100, 100), mut_rate=0.02, indel_frac=0.5, indel_extend=0.7),
    "illumina_long_ins": dict(seed=5, N=3000, length=(100, 100), mut_rate=0.01, indel_frac=0.9, indel_extend=0.97,
                              indel_min=3),
    "illumina_inner_matepair": dict(seed=11, N=3000, is_inner=1, dist=300, strandedness=1),
    "illumina_readone_fwd": dict(seed=12, N=2000, read_one_strand=1, strandedness=2),
    "illumina_readone_rev": dict(seed=13, N=2000, read_one_strand=2),
    "illumina_single_end": dict(seed=14, N=3000, length=(120, 0)),
    "illumina_maxn_hap": dict(seed=15, N=3000, max_n=5, is_hap=1, mut_freq=0.8, rand_read=0.2),
    "illumina_fixedq_prefix": dict(seed=16, N=1500, fixed_quality="I", read_prefix="pfx", reads_output_type=1),
    "illumina_q0_bfast": dict(seed=17, N=1500, quality_std=0, reads_output_type=2),
    "illumina_bigqstd": dict(seed=18, N=1500, quality_std=30, e="0.0-0.3"),
    "illumina_tight_insert": dict(seed=19, N=2000, dist=150, std_dev=40, length=(100, 100)),
    "illumina_norand_noerr": dict(seed=20, N=2000, rand_read=0, e=0, E=0, mut_rate=0),
    "mutations_only": dict(seed=21, output_type=2, mut_rate=0.01, indel_frac=0.3),
    "coverage_zero": dict(seed=21, C=0, mut_rate=0.01, indel_frac=0.3),
    "solid_2x50": dict(seed=22, N=4000, data_type=1, length=(50, 50), mut_rate=0.02, indel_frac=0.5),
    "solid_single_bwa": dict(seed=23, N=2000, data_type=1, length=(35, 0), reads_output_type=1),
    "solid_maxn": dict(seed=24, N=2000, data_type=1, length=(50, 50), max_n=3, strandedness=2),
    "ion_400_se": dict(seed=25, N=1500, data_type=2, length=(400, 0), e=0.01, flow_order=FLOW),
    "ion_paired_higherr": dict(seed=26, N=1500, data_type=2, length=(200, 100), e=0.05, E=0.03, flow_order=FLOW,
                               mut_rate=0.01, indel_frac=0.4),
    "ion_short_flow4": dict(seed=27, N=1500, data_type=2, length=(30, 30), e=0.1, E=0.1, flow_order="TACG",
                            dist=200, std_dev=10),
    "ion_base_error_B": dict(seed=28, N=500, data_type=2, length=(100, 0), e=0.02, flow_order=FLOW, use_base_error=1),
    "amplicons": dict(seed=29, N=1000, amplicons=1, length=(100, 100), max_n=60),
    # -x: `regions` is written to a BED file by materialize(); the first set has an overlap that merges,
    # a contig without regions (skip #0) and, under -N, the last contig keeping its full length
    "regions_N": dict(seed=30, N=3000, regions=[("chrA", 1000, 9000), ("chrA", 8500, 12000), ("chrA", 15000, 22000),
                                                 ("chrB", 2000, 10000), ("hp", 100, 7000)]),
    "regions_C": dict(seed=31, C=4, length=(100, 100), regions=[("chrA", 1000, 9000), ("chrA", 15000, 22000),
                                                                 ("chrB", 2000, 10000), ("hp", 100, 7000)]),
    "regions_single_end": dict(seed=32, C=3, length=(100, 0), regions=[("chrB", 2000, 3000), ("chrB", 3001, 5000),
                                                                        ("hp", 1, 2000), ("hp", 2000, 2600)]),
    "regions_skip_n": dict(seed=33, C=3, length=(80, 80), dist=300, std_dev=20,
                           regions=[("chrA", 5010, 5590), ("chrB", 1000, 11000), ("hp", 500, 7500)]),
    # -m / -v / -b: mutations replayed from a file (fixtures next to this script; the .txt and .vcf are what the
    # reference itself wrote for `-z 40 -r 0.01 -R 0.3 -X 0.6 -M 2`, the .vcf with two untagged records appended,
    # the .bed is hand-written: explicit and random bases, an overlap that is ignored, an N, a 26-base insertion)
    "replay_txt": dict(seed=41, N=2000, muts_txt="replay_muts.txt"),
    "replay_vcf": dict(seed=42, N=2000, muts_vcf="replay_muts.vcf"),
    "replay_bed": dict(seed=43, N=2000, length=(100, 100), muts_bed="replay_muts.bed"),
    "replay_bed_hap_C": dict(seed=44, C=2, is_hap=1, muts_bed="replay_muts.bed"),
}
REPLAY_SOURCE = dict(seed=40, output_type=2, mut_rate=0.01, indel_frac=0.3, indel_extend=0.6)
REPLAY_VCF_EXTRA = "hp\t7996\t.\tA\tC\t.\t.\tnote=untagged\nhp\t7998\t.\tAC\tGT\t.\t.\tnote=untagged_again\n"


def materialize(opts, outdir):
    """options whose values are files: `regions` -> a BED file under outdir and the option fn_regions_bed"""
    opts = dict(opts)
    if "regions" in opts:
        path = os.path.join(outdir, "regions.bed")
        with open(path, "w") as f:
            for name, a, b in opts.pop("regions"):
                f.write("%s\t%d\t%d\n" % (name, a, b))
        opts["fn_regions_bed"] = path
    for k in ("muts_txt", "muts_bed", "muts_vcf"):
        if k in opts:
            opts["fn_" + k] = os.path.join(HERE, opts.pop(k))
    return opts


def md5_of(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        return hashlib.md5(f.read()).hexdigest()


def run_ref(fasta, opts, outdir):
    import shutil
    sub = os.path.join(outdir, "case")
    shutil.rmtree(sub, ignore_errors=True)
    os.makedirs(sub)
    prefix = os.path.join(sub, "ref")
    argv = [po.ref_binary()] + po.opt_to_ref_argv(**materialize(opts, sub)) + [fasta, prefix]
    subprocess.run(argv, check=True, stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
    out = {}
    for f in FILES:
        p = prefix + "." + f + (".gz" if f.endswith("fastq") else "")
        out[f] = md5_of(p) if os.path.exists(p) else None
    return out


def main():
    po.build()
    res = {}
    with tempfile.TemporaryDirectory() as td:
        fasta = synth_fasta(os.path.join(td, "synth.fa"))
        res["_fasta_md5"] = md5_of(fasta)
        # the replay fixtures: the reference's own mutation files of one run
        sub = os.path.join(td, "replay_src")
        os.makedirs(sub)
        subprocess.run([po.ref_binary()] + po.opt_to_ref_argv(**REPLAY_SOURCE) + [fasta, os.path.join(sub, "src")], check=True,
                       stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
        with open(os.path.join(sub, "src.mutations.txt")) as f, open(os.path.join(HERE, "replay_muts.txt"), "w") as g:
            g.write(f.read())
        with open(os.path.join(sub, "src.mutations.vcf")) as f, open(os.path.join(HERE, "replay_muts.vcf"), "w") as g:
            g.write(f.read() + REPLAY_VCF_EXTRA)
        for name, opts in MATRIX.items():
            res[name] = run_ref(fasta, opts, td)
            print(name, res[name])
    with open(os.path.join(ROOT, "tests", "golden", "ref_matrix.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
