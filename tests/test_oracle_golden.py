"""Pin the oracle (drand48 backend) against the reference's own golden files and primitives.

Reference test being mirrored: testdata/test.sh:18 (`dwgsim -z 13 -N 10000 ex1.fa`, five files diffed).
"""
import ctypes as C
import hashlib
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def md5(path):
    with open(path, "rb") as f:
        return hashlib.md5(f.read()).hexdigest()


def test_drand48_known_answers(oracle):
    # glibc srand48(13) stream, state after seeding = (seed << 16) + 0x330E
    L = oracle.lib()
    L.orc_srand48(13)
    assert L.orc_drand48_state() == (13 << 16) + 0x330E
    got = [L.orc_drand48() for _ in range(4)]
    assert got == [0.49125804875894019, 0.9095780156526132, 0.69616396868708463, 0.92223566756269548]


def test_philox4x32_10_known_answers(oracle):
    # Random123 kat_vectors for philox4x32-10
    L = oracle.lib()

    def run(ctr, key):
        out = (C.c_uint32 * 4)()
        L.orc_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        return list(out)

    assert run([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_reference_golden_files(oracle, ex1_fa, tmp_path):
    assert md5(ex1_fa) == "2be5bfebdd7764be3af95881ddcc1471"
    want = json.load(open(os.path.join(HERE, "golden", "ex1_golden_md5.json")))
    opt = oracle.make_opt(seed=13, N=10000)
    prefix = str(tmp_path / "ex1.test")
    with oracle.Session(opt, ex1_fa, prefix) as s:
        assert s.stats.error == 0
        assert s.stats.n_pairs_total == 10000
    for name, digest in want.items():
        if name.startswith("_"):
            continue
        assert md5(prefix + "." + name) == digest, name
    # record counts the reference test implies: 40,000 / 40,000 / 80,000 lines
    assert sum(1 for _ in open(prefix + ".bwa.read1.fastq")) == 40000
    assert sum(1 for _ in open(prefix + ".bfast.fastq")) == 80000


def test_philox_backend_mutations_equal_reference_C0(oracle, ex1_fa, tmp_path):
    """mutation files of the Philox backend == drand48 backend with -C 0"""
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    with oracle.Session(oracle.make_opt(seed=13, C=0, mut_rate=0.01), ex1_fa, a) as s:
        assert s.stats.error == 0
    with oracle.Session(oracle.make_opt(seed=13, N=2000, mut_rate=0.01), ex1_fa, b, mode=oracle.RNG_PHILOX) as s:
        assert s.stats.error == 0 and s.stats.n_pairs_total == 2000
    for f in ("mutations.txt", "mutations.vcf"):
        assert md5(a + "." + f) == md5(b + "." + f)
