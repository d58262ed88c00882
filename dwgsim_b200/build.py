"""Build libdwgsim_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

nvcc cross-compiles without a GPU, so this also is the "does it build" check of __graft_entry__.build().
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libdwgsim_b200.so")
SOURCES = ["dwgsim_gpu.cu"]
DEPS = ["dwgsim_gpu.cu", "kernels.cuh", "layout.h", "flow_model.h", "gz_device.cuh", "gz_host.h", os.path.join("..", "..", "include", "dwgsim_gpu.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared"]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


CLI = os.path.join(HERE, "bin", "dwgsim")
CLI_SRC = os.path.join(CSRC, "host", "dwgsim_cli.cpp")


def build_cli(force=False):
    """the drop-in `dwgsim` host shell (g++), linked against libdwgsim_b200.so and zlib"""
    if not force and os.path.exists(CLI) and os.path.getmtime(CLI) > max(os.path.getmtime(CLI_SRC), os.path.getmtime(SO)):
        return CLI
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-o", CLI, CLI_SRC, "-L" + HERE, "-ldwgsim_b200", "-lz", "-lpthread",
           "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building the dwgsim host shell")
    return CLI


def build(force=False, verbose=False):
    if not force and not stale():
        build_cli()
        return SO
    cmd = [nvcc()] + NVCC_FLAGS + os.environ.get("NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libdwgsim_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    build_cli(force=True)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
