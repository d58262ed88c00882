"""Host tables of the device gzip writer (dwgsim_b200/csrc/gz_host.h): length-limited Huffman codes, the constant member
prefix and the CRC-32 tables, validated by decoding the CPU-encoded members with zlib (gzip module)."""
import ctypes as C
import gzip
import os

import pytest


@pytest.fixture(scope="module")
def lib():
    from dwgsim_b200 import build, _lib
    build.build()
    return _lib.load()


def encode(lib, data):
    out = C.create_string_buffer(2 * len(data) + 4096)
    n = C.c_uint64()
    assert lib.dwgsim_gpu_gz_host_encode(data, len(data), out, len(out), C.byref(n)) == 0
    return out.raw[:n.value]


@pytest.mark.parametrize("name", ["fastq", "random", "zeros", "one_byte", "skewed", "exact_member", "member_plus_one"])
def test_members_decode_with_zlib(lib, name):
    fq = (b"@chr1_1234567_1234999_0_1_0_0_1:0:0_2:0:0_1f3a/1\n" + b"ACGTTGCAAN" * 15 + b"\n+\n" + b"5678:;<=>?" * 15 + b"\n") * 700
    data = {
        "fastq": fq,
        "random": os.urandom(200001),
        "zeros": bytes(100000),
        "one_byte": b"A",
        "skewed": b"A" * 300000 + bytes(range(256)),        # forces the 15-bit length limit
        "exact_member": os.urandom(65536),
        "member_plus_one": os.urandom(65537),
    }[name]
    z = encode(lib, data)
    assert z[:4] == b"\x1f\x8b\x08\x00"
    assert gzip.decompress(z) == data
    if name == "fastq":
        assert len(z) < 0.6 * len(data)
