"""CPU-side checks of the drop-in boundary: the shared library loads and exports every entry point
include/dwgsim_gpu.h declares; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dwgsim_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dwgsim_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dwgsim_gpu_[a-z_0-9]+)\s*\(", text)) - {"dwgsim_gpu_sink_fn"})


def test_every_declared_symbol_is_exported(lib):
    from dwgsim_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libdwgsim_b200.so does not export %s" % n
        assert n in _lib.SYMBOLS, "dwgsim_b200/_lib.py has no prototype for %s" % n
    assert sorted(_lib.SYMBOLS) == names


def test_abi_version_and_strerror(lib):
    assert lib.dwgsim_gpu_abi_version() == 2
    assert lib.dwgsim_gpu_strerror(0) == b"ok"
    assert b"10001 trials" in lib.dwgsim_gpu_strerror(-5)      # reference message, src/dwgsim.c:838


def test_struct_sizes_match_header(lib, tmp_path):
    """compile the header as plain C (gcc) and compare sizeof with the ctypes mirrors"""
    import subprocess
    from dwgsim_b200._lib import Params, Stats, Batch, Tables
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "dwgsim_gpu.h"\nint main(void){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(dwgsim_gpu_params_t),sizeof(dwgsim_gpu_stats_t),sizeof(dwgsim_gpu_batch_t),'
                   'sizeof(dwgsim_gpu_tables_t));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(Params), C.sizeof(Stats), C.sizeof(Batch), C.sizeof(Tables)]


def test_option_mirror_applies_reference_parse_rules():
    from dwgsim_b200 import params_from_options
    p = params_from_options(e="0.001-0.01", E="0.002,0.004", length=(150, 100), seed=7)
    assert p.e_start[0] == 0.001 and abs(p.e_by[0] - (0.01 - 0.001) / 150) < 1e-18
    assert p.e_start[1] == 0.002 and abs(p.e_by[1] - (0.004 - 0.002) / 100) < 1e-18
    q = params_from_options()
    assert (q.dist, q.std_dev, q.length[0], q.rand_read, q.quality_std) == (500, 50.0, 70, 0.05, 2.0)


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dwgsim_b200 import DwgsimGpu, DwgsimGpuError, params_from_options
    with pytest.raises(DwgsimGpuError) as e:
        DwgsimGpu(params_from_options(seed=1))
    assert e.value.code == -2


def test_product_never_references_the_oracle():
    """the product tree must not import, link or execute anything under oracle/"""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "dwgsim_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")):
                t = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"(?m)^\s*(from|import)\s+oracle|liboracle|orc_run|#include\s+\"[^\"]*oracle", t):
                    bad.append(f)
    assert not bad, bad
