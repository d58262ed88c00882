"""-m gpu: BASELINE.json's full-size configuration (3.1 Gbp synthetic reference, Illumina 2x150, -C 30) through
size-independent properties -- the oracle cannot run at this size.  Record grammar and geometry of every record of a
batch, coordinates inside the contig, determinism, and independence from how the pair range is cut into batches."""
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
          133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
          58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
NAME = re.compile(rb"^@(chr\d+|rand)_(\d+)_(\d+)_([01])_([01])_([01])_([01])_(\d+):(\d+):(\d+)_(\d+):(\d+):(\d+)_([0-9a-f]+)$")


@pytest.fixture(scope="module")
def gpu():
    from dwgsim_b200 import DwgsimGpu, params_from_options
    g = DwgsimGpu(params_from_options(length=(150, 150), e="0.001-0.01", E="0.001-0.01", seed=1))
    g.genome_synthetic(GRCH38, 20261017, 0.001, 0.1, 0.01, 30.0)
    g.genome_finalize()
    yield g
    g.close()


def streams(gpu, first, n, rand_base=0):
    b = gpu.simulate_resident(first, n, rand_base)
    return [gpu.copy_stream(k, b.n_bytes[k]) for k in range(3)], b


def test_pair_budget_is_the_reference_formula(gpu):
    # src/dwgsim.c:589: (uint64)(l * C / (s0 + s1) / (1 - y) + 0.5) per contig
    want = sum(int(l * 30.0 / 300.0 / 0.95 + 0.5) for l in GRCH38)
    assert abs(gpu.genome_pairs() - want) <= len(GRCH38)


def test_record_grammar_and_geometry_at_full_size(gpu):
    n = 1 << 17
    first = gpu.genome_pairs() // 2 + 12345          # middle of the job: several hundred million pairs in
    (r1, r2, bf), b = streams(gpu, first, n)
    assert b.n_pairs == n and 0.03 * n < b.n_random < 0.07 * n
    for s, suffix in ((r1, b"/1"), (r2, b"/2")):
        lines = s.split(b"\n")
        assert len(lines) == 4 * n + 1 and lines[-1] == b""
        names, seqs, plus, quals = lines[0:-1:4], lines[1:-1:4], lines[2:-1:4], lines[3:-1:4]
        assert all(p == b"+" for p in plus[:2000]) and set(plus) == {b"+"}
        assert set(map(len, seqs)) == {150} and set(map(len, quals)) == {150}
        assert set(b"".join(seqs[:5000])) <= set(b"ACGTN")
        q = np.frombuffer(b"".join(quals[:5000]), dtype=np.uint8)
        assert q.min() >= 33 and q.max() <= 73
        for nm in names[:20000]:
            assert nm.endswith(suffix)
            m = NAME.match(nm[:-2])
            assert m, nm
            if m.group(1) != b"rand":
                c = int(m.group(1)[3:]) - 1
                p1, p2 = int(m.group(2)), int(m.group(3))
                assert 1 <= p1 <= GRCH38[c] - 149 and 1 <= p2 <= GRCH38[c] - 149
                assert m.group(4) != m.group(5)                      # Illumina: opposite strands
                assert abs(abs(p2 - p1) + 150 - 500) < 8 * 50 + 40   # insert size within 8 sigma (+ indels)
    bl = bf.split(b"\n")
    assert len(bl) == 8 * n + 1
    assert bl[0] + b"/1" == r1.split(b"\n", 1)[0] and bl[4] + b"/2" == r2.split(b"\n", 1)[0]
    assert bl[1::8] == r1.split(b"\n")[1:-1:4] and bl[5::8] == r2.split(b"\n")[1:-1:4]


def test_determinism_and_batch_split_independence(gpu):
    first, n = 200_000_000, 60_000
    whole, b = streams(gpu, first, n, rand_base=777)
    again, _ = streams(gpu, first, n, rand_base=777)
    assert whole == again
    h1, b1 = streams(gpu, first, 25_000, rand_base=777)
    h2, _ = streams(gpu, first + 25_000, n - 25_000, rand_base=777 + b1.n_random)
    for k in range(3):
        assert h1[k] + h2[k] == whole[k]


@pytest.mark.parametrize("base", [0xFFFFFFF0, (1 << 36) + 5])
def test_random_pair_serials_of_eight_and_more_hex_digits(gpu, base):
    """the running count of random pairs (rand_ii, src/dwgsim.c:1096) is printed in hexadecimal at the end of the name: the
    name writer assembles eight digits per store, so cross 2^32 and go past it"""
    first, n = 150_000_000, 4_000
    (r1, r2, bf), b = streams(gpu, first, n, rand_base=base)
    (s1, _, _), _ = streams(gpu, first, n, rand_base=0)
    l0, l1 = s1.split(b"\n"), r1.split(b"\n")
    assert l0[1::4] == l1[1::4] and l0[3::4] == l1[3::4]          # bases and qualities do not depend on the count
    k = 0
    for small, big in zip(l0[0:-1:4], l1[0:-1:4]):
        ms, mb = NAME.match(small[:-2]), NAME.match(big[:-2])
        assert ms and mb, (small, big)
        if mb.group(1) == b"rand":
            assert int(ms.group(14), 16) == k and int(mb.group(14), 16) == base + k
            assert mb.group(14) == b"%x" % (base + k)              # no leading zeros
            k += 1
        else:
            assert small == big
    assert k == b.n_random and k > 100
    names2 = r2.split(b"\n")[0:-1:4]
    assert [x[:-2] for x in names2] == [x[:-2] for x in l1[0:-1:4]] and bf.split(b"\n")[0::8][:-1] == [x[:-2] for x in l1[0:-1:4]]


def test_queued_batches_equal_waited_batches(gpu):
    """dwgsim_gpu_resident_enqueue queues batches without a host wait; rand_ii continues from the counter in device memory"""
    first, n, base = 120_000_000, 30_000, 4242
    want, run = [], base
    for k in range(3):                                   # three waited batches, the running count carried on the host
        s, b = streams(gpu, first + k * n, n, rand_base=run)
        want.append(s)
        run += b.n_random
        per_batch = b.n_launches
    gpu.resident_set_running(base)
    for k in range(3):
        gpu.resident_enqueue(first + k * n, n)
    b = gpu.resident_wait()
    got = [gpu.copy_stream(k, b.n_bytes[k]) for k in range(3)]
    assert got == want[2] and b.n_pairs == n and b.n_launches == 3 * per_batch
    gpu.resident_enqueue(first + 3 * n, n)               # the counter went on: batch 3 continues where batch 2 ended
    b3 = gpu.resident_wait()
    got3 = [gpu.copy_stream(k, b3.n_bytes[k]) for k in range(3)]
    s3, _ = streams(gpu, first + 3 * n, n, rand_base=run)
    assert got3 == s3
