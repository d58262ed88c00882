"""-m gpu: oracle parity at production coordinates.  One contig of 210 Mbp (nine-digit coordinates, block-index entries
beyond 2^20, the multi-threaded packer and its part merge, pool re-basing of long insertions) followed by a contig with a
name of more than 200 characters; -r 0.001 -R 0.15 with long insertions, N runs, through dwgsim_gpu_add_contig and through
the drop-in binary with a .fai.  Every output stream is compared with the oracle byte for byte."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import gpu_harness as gh  # noqa: E402

pytestmark = pytest.mark.gpu
BIG = 210_000_000
LONG_NAME = "chrUn_" + "KI270706v1_random_" * 11 + "alt"       # 207 characters


def write_fasta(path):
    rng = np.random.default_rng(20261017)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as f, open(path + ".fai", "w") as fai:
        off = 0
        for name, n in (("chrBig", BIG), (LONG_NAME, 300_000)):
            s = acgt[rng.integers(0, 4, n, dtype=np.uint8)]
            s[:10_000] = ord("N"); s[n - 10_000:] = ord("N")        # telomeres
            s[n // 3: n // 3 + n // 100] = ord("N")                 # one long run
            head = (">%s\n" % name).encode()
            f.write(head)
            off += len(head)
            rows = n // 60
            body = np.empty((rows, 61), dtype=np.uint8)
            body[:, :60] = s[:rows * 60].reshape(rows, 60)
            body[:, 60] = 10
            f.write(body.tobytes())
            if n % 60:
                f.write(s[rows * 60:].tobytes() + b"\n")
            fai.write("%s\t%d\t%d\t60\t61\n" % (name, n, off))
            off += n + rows + (1 if n % 60 else 0)
    return path


@pytest.fixture(scope="module")
def big_fa(tmp_path_factory):
    return write_fasta(str(tmp_path_factory.mktemp("big") / "big.fa"))


OPTS = dict(seed=11, N=200000, length=(150, 150), e="0.001-0.01", E="0.001-0.01", mut_rate=0.001, indel_frac=0.15,
            indel_extend=0.9, read_prefix="prod")


def test_big_contig_through_the_c_abi(oracle, big_fa, tmp_path):
    sess, want = gh.oracle_expected(oracle, OPTS, big_fa, str(tmp_path / "orc"))
    try:
        assert sess.stats.error == 0
        assert sess.stats.n_pairs_total == OPTS["N"]
        got, stats = gh.gpu_actual(sess, OPTS, orc_opt=sess.opt, batch=1 << 16)
    finally:
        sess.close()
    for i, name in enumerate(gh.FILE_NAMES):
        assert got[i] == want[i], "%s: %s" % (name, gh.first_diff(want[i], got[i]))
    # the coordinates really are nine digits wide and the long name made it into the records
    import re
    assert re.search(rb"@prod_chrBig_\d{9}_\d{9}_", want[0]) and LONG_NAME.encode() in want[0]


def test_big_contig_through_the_drop_in(oracle, big_fa, tmp_path):
    from dwgsim_b200 import build
    exe = build.build_cli()
    opt = oracle.make_opt(**OPTS)
    orc = str(tmp_path / "orc")
    with oracle.Session(opt, big_fa, orc, mode=oracle.RNG_PHILOX) as sess:
        assert sess.stats.error == 0
    prefix = str(tmp_path / "cli")
    argv = [exe] + oracle.opt_to_ref_argv(**OPTS) + ["--uncompressed", "--batch", "65536", big_fa, prefix]
    r = subprocess.run(argv, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]

    def md5(p):
        h = hashlib.md5()
        with open(p, "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
        return h.hexdigest()
    for f in gh.FILE_NAMES + ["mutations.txt", "mutations.vcf"]:
        assert md5(prefix + "." + f) == md5(orc + "." + f), f
