#!/usr/bin/env python
"""bench.py -- read-pairs/s of the dwgsim_core read-pair path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

Workload (BASELINE.json configs[1], SURVEY.md 8d "Config 2"): 3.1 Gbp synthetic reference (24 contigs with
GRCh38 lengths, i.i.d. ACGT, ~1 % N, SNP+indel events at -r 0.001 -R 0.1), Illumina 2x150,
-e/-E 0.001-0.01, -d 500 -s 50 -y 0.05, all three FASTQ outputs (-o 0), -C 30 => ~326 M pairs.
A step = one batch of PAIRS_PER_STEP consecutive pair indices of that job through the whole hot path
(simulate -> layout -> format), genome resident in HBM, FASTQ bytes left in HBM (`value`).
`e2e` drives the C ABI the way the reference host would: dense seq_t/mutseq_t host arrays in
(dwgsim_gpu_add_contig), FASTQ bytes out to host memory in pair order (dwgsim_gpu_run + sink).
"""
import argparse
import ctypes as C
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# GRCh38 primary assembly chr1..22, X, Y
GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
          133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
          58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
OPTS = dict(length=(150, 150), e="0.001-0.01", E="0.001-0.01", seed=1)       # everything else: reference defaults
REF_ARGV = ["-1", "150", "-2", "150", "-e", "0.001-0.01", "-E", "0.001-0.01"]
COVERAGE = 30.0
MUT_RATE, INDEL_FRAC, N_FRAC = 0.001, 0.1, 0.01
PAIRS_PER_STEP = 1 << 20
ALGO_BYTES_PER_PAIR = 1524.0      # 118 B read (2-bit ref + N mask + mutation table) + 1,406 B FASTQ written
E2E_CONTIG_LEN = 8 << 20          # contig handed over per e2e step (dense host arrays: 17 B/base)
# dram__bytes_read.sum + dram__bytes_write.sum of the eight launches of one 2^20-pair step, from the ncu --set full
# capture profiles/r01c_ncu_key_metrics.txt (reads 0.743 GB + writes 1.685 GB)
NCU_DRAM_BYTES_PER_STEP = 2.427e9


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=fd,
                                         stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ---------------------------------------------------------------------------------------------------
# CPU reference arm
# ---------------------------------------------------------------------------------------------------
def write_sample_fasta(path, n_bases, seed=20261017):
    import numpy as np
    rng = np.random.default_rng(seed)
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n_bases)].copy()
    s[n_bases // 3:n_bases // 3 + n_bases // 100] = ord("N")
    with open(path, "wb") as f:
        f.write(b">chrS\n")
        rows = s[:n_bases - n_bases % 60].reshape(-1, 60)
        out = np.empty((rows.shape[0], 61), dtype=np.uint8)
        out[:, :60] = rows
        out[:, 60] = 10
        f.write(out.tobytes())
        if n_bases % 60:
            f.write(s[n_bases - n_bases % 60:].tobytes() + b"\n")


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "dwgsim_ref")
    return p if os.path.exists(p) else None


def run_ref_round(binary, fasta, workdir, n_proc, n_pairs, seed0, coverage_zero=False):
    """n_proc concurrent single-threaded reference processes; returns wall seconds"""
    procs = []
    t0 = time.perf_counter()
    for i in range(n_proc):
        prefix = os.path.join(workdir, "r%d" % i)
        argv = [binary] + REF_ARGV + ["-z", str(seed0 + i)] + (["-C", "0"] if coverage_zero else ["-N", str(n_pairs)]) + \
               [fasta, prefix]
        procs.append(subprocess.Popen(argv, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference binary failed")
    return time.perf_counter() - t0


def cpu_reference(n_proc, n_pairs, rounds, warmup, sample_bases=2_000_000):
    """time oracle/_ref/dwgsim_ref (the unmodified reference) on a bounded sample of the workload.
    Loop time = wall time of a round minus the prologue (FASTA census + mut_diref), measured with -C 0."""
    binary = ref_binary()
    if binary is None:
        raise RuntimeError("oracle/_ref/dwgsim_ref not present")
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    wd = tempfile.mkdtemp(prefix="dwgsim_ref_", dir=base)
    try:
        fasta = os.path.join(wd, "sample.fa")
        write_sample_fasta(fasta, sample_bases)
        pro = []
        for w in range(max(warmup, 1)):
            pro.append(run_ref_round(binary, fasta, wd, n_proc, 0, 1000 + w, coverage_zero=True))
        prologue = statistics.median(pro)
        times = [run_ref_round(binary, fasta, wd, n_proc, n_pairs, 1 + r * n_proc) for r in range(rounds)]
        loop = [max(t - prologue, 1e-9) for t in times]
        total = sum(loop)
        return dict(value=n_proc * n_pairs * rounds / total, ms_per_step=1e3 * total / rounds, prologue_s=prologue,
                    wall_s=sum(times))
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_pairs = int(max(3000, min(12000, 600000 // max(args.steps, 1))))
    sample = ("oracle/_ref/dwgsim_ref (unmodified reference, gcc -O3), %d concurrent single-threaded processes x -N %d pairs "
              "per step on a 2 Mbp sample of the synthetic reference, all three .fastq.gz outputs; prologue "
              "(census + mut_diref, measured with -C 0) subtracted" % (cores, n_pairs))
    line = {"impl": "reference", "metric": "read-pairs/sec (2x150bp, 3.1Gbp ref)", "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": config_dict(args.gpus), "gpu_launches": 0}
    try:
        r = cpu_reference(cores, n_pairs, args.steps, args.warmup)
        line.update(value=r["value"], ms_per_step=r["ms_per_step"],
                    cpu_baseline={"value": r["value"], "unit": "pairs/s", "cores": cores, "kind": "reference", "sample": sample},
                    e2e={"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    except Exception as e:  # the oracle port is the fallback the tier allows
        line.update(unavailable="reference binary could not run: %s" % e)
    emit(line)


def config_dict(n_gpus):
    return {"workload": "configs[1]: 3.1 Gbp synthetic reference (24 contigs, GRCh38 lengths), Illumina 2x150bp, "
                        "-e/-E 0.001-0.01, -r 0.001 -R 0.1, -C 30 (~326M pairs), -o 0 (bwa1+bwa2+bfast)",
            "pairs_per_step": PAIRS_PER_STEP, "l2": "inputs (1.5 GB genome blob) and outputs (1.5 GB/step) larger than L2",
            "parallelism": "pair-index shards x%d" % n_gpus}


# ---------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------
class _CudaArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def dense_contig(n_bases, seed):
    """seq_t + 2 x mutseq_t.s of one synthetic contig as the reference host would hold them (numpy)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, n_bases, dtype=np.uint8)
    codes[n_bases // 3:n_bases // 3 + n_bases // 100] = 4
    seq = np.frombuffer(b"ACGTN", dtype=np.uint8)[codes]
    hap = [codes.astype(np.uint64), codes.astype(np.uint64)]
    n_mut = int(n_bases * MUT_RATE)
    pos = np.unique(rng.integers(0, n_bases, n_mut))
    pos = pos[codes[pos] < 4]
    kind = rng.random(pos.size)
    zyg = rng.integers(0, 3, pos.size)                       # 0 hom, 1 hap1, 2 hap2
    c = codes[pos].astype(np.uint64)
    sub = (0x20 | ((c + 1 + rng.integers(0, 3, pos.size).astype(np.uint64)) & 3)).astype(np.uint64)
    dele = (0x30 | c).astype(np.uint64)
    nins = rng.integers(1, 4, pos.size).astype(np.uint64)
    bases = rng.integers(0, 64, pos.size).astype(np.uint64) & ((np.uint64(1) << (2 * nins)) - 1)
    ins = ((nins << np.uint64(59)) | (bases << np.uint64(6)) | np.uint64(0x10) | c).astype(np.uint64)
    val = np.where(kind >= INDEL_FRAC, sub, np.where(kind < INDEL_FRAC / 2, dele, ins))
    for h in (0, 1):
        m = (zyg == 0) | (zyg == h + 1)
        hap[h][pos[m]] = val[m]
    return np.ascontiguousarray(seq), hap


_REAL_STDOUT = None


def protect_stdout():
    """the driver reads ONE JSON line from stdout; libraries (NCCL prints its version) must not get in the way:
    everything written to fd 1 from now on goes to stderr, emit() writes to the real stdout"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--genome-scale", type=float, default=1.0, help="shrink the synthetic genome (debug only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from dwgsim_b200 import DwgsimGpu, params_from_options, build
    if not os.path.exists(build.SO):          # normally prebuilt in-tree and shipped with the snapshot
        if rank == 0:
            build.build()
        else:
            for _ in range(600):
                if os.path.exists(build.SO):
                    break
                time.sleep(0.5)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the read-pair path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    gpu = DwgsimGpu(params_from_options(**OPTS), device=local_rank)
    lengths = [max(int(x * args.genome_scale), 200000) for x in GRCH38]
    # ---- genome: rank 0 builds and packs it; NCCL broadcasts the packed blob to the other ranks ----
    t0 = time.perf_counter()
    keep = None
    if rank == 0:
        gpu.genome_synthetic(lengths, 20261017, MUT_RATE, INDEL_FRAC, N_FRAC, COVERAGE)
        gpu.genome_finalize()
        ptr, nbytes = gpu.genome_blob()
    if world > 1:
        meta = torch.zeros(1, dtype=torch.int64, device="cuda")
        if rank == 0:
            meta[0] = nbytes
        dist.broadcast(meta, 0)
        nbytes = int(meta.item())
        if rank == 0:
            blob_t = torch.as_tensor(_CudaArray(ptr, nbytes), device="cuda")
        else:
            blob_t = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        tb = time.perf_counter()
        dist.broadcast(blob_t, 0)
        torch.cuda.synchronize()
        bcast_s = time.perf_counter() - tb
        if rank != 0:
            gpu.genome_import(blob_t.data_ptr(), nbytes, take_ownership=False)
        keep = blob_t
    else:
        bcast_s = 0.0
    setup_s = time.perf_counter() - t0
    total_pairs = gpu.genome_pairs()
    B = PAIRS_PER_STEP
    n_steps_avail = total_pairs // (B * world)
    if n_steps_avail < 1:
        B = int(total_pairs // world)
        n_steps_avail = 1

    stream = torch.cuda.ExternalStream(gpu.cuda_stream(), device=torch.device("cuda", local_rank))
    rand_base = 0

    # N > 1: the one exchange of the path -- random-pair counts of the round, so that rand_ii in the names stays the
    # global running count (src/dwgsim.c:1096) -- as an NCCL all-gather of device counters enqueued on the library's
    # stream between the simulate passes and the layout kernels: no host round trip inside a step
    if world > 1:
        allc = torch.zeros(world, dtype=torch.int64, device="cuda")
        base_t = torch.zeros(1, dtype=torch.int64, device="cuda")
        running_t = torch.zeros(1, dtype=torch.int64, device="cuda")

    def step(k):
        nonlocal rand_base
        first = ((k % n_steps_avail) * world + rank) * B
        if world > 1:
            gpu.resident_begin(first, B, False)
            cnt = torch.as_tensor(_CudaArray(gpu.resident_count_ptr(), 8), device="cuda").view(torch.int64)
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(allc, cnt)
                torch.add(running_t, allc[:rank].sum(), out=base_t)
                b = gpu.resident_finish_dev(base_t.data_ptr())
                running_t.add_(allc.sum())
        else:
            b = gpu.simulate_resident(first, B, rand_base)
            rand_base += b.n_random
        return b

    for k in range(args.warmup):
        step(k)
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    ms = [0.0, 0.0, 0.0]
    out_bytes = 0
    launches = 0
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        b = step(args.warmup + k)
        ms[0] += b.ms_simulate; ms[1] += b.ms_layout; ms[2] += b.ms_format
        out_bytes += sum(b.n_bytes)
        launches += b.n_launches
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # ---- roofline of the kernels (device time from CUDA events inside the library, per launch group) ----
    peak, peak_src = peak_hbm()
    kern_ms = sum(ms) / args.steps
    achieved = ALGO_BYTES_PER_PAIR * B / (kern_ms * 1e-3) / 1e9
    dom = max(range(3), key=lambda i: ms[i])
    names = ["simulate_pairs_tp_kernel (2 passes)", "layout_* (5 scan kernels)", "format_fastq_kernel"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_PER_STEP if B == PAIRS_PER_STEP else None,
                "traffic_source": "ncu --set full capture of the same step (profiles/r01c_ncu_key_metrics.txt), bytes per step",
                "peak_source": peak_src,
                "kernel": "whole step = simulate (2 passes) + layout + format (8 launches); dominant: %s" % names[dom],
                "algorithmic_bytes_per_pair": ALGO_BYTES_PER_PAIR, "fastq_bytes_per_pair": out_bytes / (B * args.steps),
                "ms_per_step_by_kernel": {n: m / args.steps for n, m in zip(names, ms)}}
    # the dominant kernel on its own: the formatter's algorithmic bytes are the FASTQ text it writes (its inputs are the
    # intermediate records and codes); achieved = those bytes / its CUDA-event time
    fmt_ms = ms[2] / args.steps
    if fmt_ms > 0:
        fmt_bytes = out_bytes / args.steps
        roofline["dominant_kernel"] = {"name": "format_fastq_kernel", "ms": fmt_ms, "algorithmic_bytes": fmt_bytes,
                                       "achieved": fmt_bytes / (fmt_ms * 1e-3) / 1e9, "unit": "GB/s",
                                       "frac": fmt_bytes / (fmt_ms * 1e-3) / 1e9 / peak}

    # ---- e2e: the C ABI with host buffers (dense arrays in, FASTQ bytes out to host memory) ----
    # Headline leg: the drop-in's default output, .fastq.gz bytes (what the reference writes to its gzFiles,
    # src/dwgsim.c:1151-1157), compressed on the GPU before the device->host copy.  Second leg: plain FASTQ text.
    e2e = None
    if not args.no_e2e:
        seq, hap = dense_contig(E2E_CONTIG_LEN, 7 + rank)
        n_pairs_c = int(E2E_CONTIG_LEN * COVERAGE / 300.0 / 0.95 + 0.5)
        n_warm = 6       # the first steps allocate the pinned ring and first-touch its pages (100+ ms stalls on a fresh box)
        n_e2e = max(5, min(args.steps, 20))

        def e2e_leg(compress):
            g2 = DwgsimGpu(params_from_options(**OPTS), device=local_rank)
            g2.set_batch(1 << 18, 3)
            if compress:
                g2.set_compression(1)

            def one(i):
                g2.add_contig(i, "chrE%d" % i, seq.ctypes.data, E2E_CONTIG_LEN, hap[0].ctypes.data, hap[1].ctypes.data,
                              None, 0, None, 0, n_pairs_c)
                return g2.run_count()

            for i in range(n_warm):
                one(i)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            te = time.perf_counter()
            h2d = d2h = raw = 0
            pack_ms = 0.0
            for i in range(n_e2e):
                st = one(n_warm + i)
                h2d += st.h2d_bytes; d2h += st.d2h_bytes; pack_ms += st.ms_pack; raw += sum(st.raw_bytes)
            torch.cuda.synchronize()
            secs = time.perf_counter() - te
            if world > 1:
                t = torch.tensor([secs], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                secs = float(t.item())
            g2.close()
            return {"value": world * n_pairs_c * n_e2e / secs, "unit": "pairs/s", "h2d_bytes_per_step": h2d // n_e2e,
                    "d2h_bytes_per_step": d2h // n_e2e, "steps": n_e2e, "pairs_per_step": n_pairs_c,
                    "host_pack_ms_per_step": pack_ms / n_e2e, "fastq_bytes_per_step": raw // n_e2e}

        what = ("per step: dwgsim_gpu_add_contig(%d Mbp contig as dense seq_t + 2 x mut_t[len] host arrays, packed on the host, "
                "copied to the device) + dwgsim_gpu_run -> %s of all three files delivered in pair order to host memory "
                "(library counting sink)")
        raw_leg = e2e_leg(False)
        raw_leg["what"] = what % (E2E_CONTIG_LEN >> 20, "FASTQ text")
        try:
            e2e = e2e_leg(True)
            e2e["compression_ratio"] = e2e["d2h_bytes_per_step"] / max(e2e["fastq_bytes_per_step"], 1)
            e2e["what"] = what % (E2E_CONTIG_LEN >> 20, ".fastq.gz bytes (gzip members written on the GPU, "
                                                         "dwgsim_gpu_set_compression(1), the host shell's default)")
            e2e["raw_sink"] = raw_leg
        except Exception as ex:  # keep a headline if the compressed leg fails
            e2e = raw_leg
            e2e["gzip_sink"] = {"error": str(ex)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and ref_binary():
        try:
            n = 60000
            r = cpu_reference(1, n, 1, 1)
            cpu_baseline = {"value": r["value"], "unit": "pairs/s", "cores": 1, "kind": "reference",
                            "sample": "oracle/_ref/dwgsim_ref -N %d on a 2 Mbp sample of the synthetic reference, same options, "
                                      "all three .fastq.gz outputs, 1 process; prologue (%.2f s, -C 0) subtracted" % (n, r["prologue_s"])}
        except Exception as e:
            cpu_baseline = {"value": None, "unit": "pairs/s", "cores": 1, "kind": "reference", "sample": "failed: %s" % e}

    if rank == 0:
        line = {"metric": "read-pairs/sec (2x150bp, 3.1Gbp ref)", "value": value, "unit": "pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_dict(world),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "setup": {"genome_build_s": setup_s, "nccl_broadcast_s": bcast_s, "genome_pairs": total_pairs,
                          "wall_s_timed_region": t_wall}}
        emit(line)
    gpu.close()
    del keep
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
