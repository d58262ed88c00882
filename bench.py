#!/usr/bin/env python
"""bench.py -- read-pairs/s of the dwgsim_core read-pair path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

Workload (BASELINE.json configs[1], SURVEY.md 8d "Config 2"): 3.1 Gbp synthetic reference (24 contigs with
GRCh38 lengths, i.i.d. ACGT, ~1 % N, SNP+indel events at -r 0.001 -R 0.1), Illumina 2x150,
-e/-E 0.001-0.01, -d 500 -s 50 -y 0.05, all three FASTQ outputs (-o 0), -C 30 => ~326 M pairs.
A step = `device_batches_per_step` consecutive device batches of PAIRS_PER_BATCH pair indices of that job through the
whole hot path (simulate -> layout -> format), genome resident in HBM, FASTQ bytes left in HBM (`value`); the number of
batches per step is chosen after the warm-up so that the K timed steps last at least MIN_TIMED_S seconds (the whole
326 M-pair job is well under a second of kernel time, so the timed region walks the job's pair indices more than once).
The batches of a step are queued back to back on the library's stream (dwgsim_gpu_resident_enqueue; rand_ii continues in
device memory) with one host wait per step; the per-kernel times of the roofline are those of each step's last batch.
`configs` repeats the kernel-path measurement for BASELINE.json configs[2..4] (-R 0.15, SOLiD 2x50, Ion Torrent 400 bp SE).
`e2e` drives the C ABI the way the reference host would: dense seq_t/mutseq_t host arrays in (dwgsim_gpu_add_contig),
.fastq.gz bytes out through dwgsim_gpu_run to host memory in pair order; `e2e.file_sink` also writes them to new files
on tmpfs every step, like the reference arm's gzFiles (bounded by the kernel's page-cache copy, a few GB/s per file).
`e2e_cli` (1 GPU) is the wall time of the drop-in binary on the full 3.1 Gbp FASTA, prologue included.
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
PAIRS_PER_BATCH = 1 << 20
MIN_TIMED_S = 1.25                # the K timed steps of the headline last at least this long (>= 5 clock samples)
MIN_TIMED_S_OTHER = 0.4           # ... and those of the other configs this long
ALGO_BYTES_PER_PAIR = 1524.0      # 118 B read (2-bit ref + N mask + mutation table) + 1,406 B FASTQ written
E2E_CONTIG_LEN = 8 << 20          # contig handed over per e2e step (dense host arrays: 17 B/base)
FLOW = "TACGTACGTCTGAGCATCGATCGATGTACAGC"
# BASELINE.json configs[1..4]; algorithmic bytes per unit: packed reference + N mask + mutation-table bytes
# read (sum over ends of ceil(len/4) * 1.5, + 5) + FASTQ bytes written (measured per run; 1,524 B fixed for the headline)
WORKLOADS = {
    "illumina_2x150": dict(opts=OPTS, coverage=30.0, indel_frac=0.1, unit="pairs/s",
                           label="configs[1]: Illumina 2x150bp, -e/-E 0.001-0.01, -r 0.001 -R 0.1, -C 30 (~326M pairs)"),
    "illumina_2x150_R0.15": dict(opts=OPTS, coverage=30.0, indel_frac=0.15, unit="pairs/s",
                                 label="configs[2]: Illumina 2x150bp, -r 0.001 -R 0.15 (SNP+indel), -C 30, pair-index shards"),
    "solid_2x50": dict(opts=dict(length=(50, 50), data_type=1, seed=1), coverage=30.0, indel_frac=0.1, unit="pairs/s",
                       label="configs[3]: SOLiD colour space -c 1, 2x50bp, -C 30 (~979M pairs)"),
    "iontorrent_400se": dict(opts=dict(length=(400, 0), data_type=2, e=0.01, flow_order=FLOW, seed=1), coverage=20.0,
                             indel_frac=0.1, unit="reads/s",
                             label="configs[4]: Ion Torrent -c 2 flow model, 400bp single-end, -e 0.01, -C 20 (~163M reads)"),
}


def traffic_record():
    """DRAM bytes of one 2^20-pair step from the committed ncu --set full capture (tools/ncu_read.py --traffic writes it)"""
    p = os.path.join(ROOT, "profiles", "r02b_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


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
            "pairs_per_device_batch": PAIRS_PER_BATCH,
            "l2": "inputs (1.5 GB genome blob) and outputs (1.5 GB per device batch) larger than L2",
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


class KernelPath:
    """one workload on the resident synthetic genome: rank 0 builds and packs it, NCCL broadcasts the packed blob"""

    def __init__(self, name, rank, local_rank, world, genome_scale):
        import torch
        import torch.distributed as dist
        from dwgsim_b200 import DwgsimGpu, params_from_options
        self.torch, self.dist = torch, dist
        self.name, self.w = name, WORKLOADS[name]
        self.rank, self.world = rank, world
        self.gpu = DwgsimGpu(params_from_options(**self.w["opts"]), device=local_rank)
        lengths = [max(int(x * genome_scale), 200000) for x in GRCH38]
        t0 = time.perf_counter()
        self.keep = None
        self.bcast_s = 0.0
        if rank == 0:
            self.gpu.genome_synthetic(lengths, 20261017, MUT_RATE, self.w["indel_frac"], N_FRAC, self.w["coverage"])
            self.gpu.genome_finalize()
            ptr, nbytes = self.gpu.genome_blob()
        if world > 1:
            meta = torch.zeros(1, dtype=torch.int64, device="cuda")
            if rank == 0:
                meta[0] = nbytes
            dist.broadcast(meta, 0)
            nbytes = int(meta.item())
            blob_t = torch.as_tensor(_CudaArray(ptr, nbytes), device="cuda") if rank == 0 else \
                torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            tb = time.perf_counter()
            dist.broadcast(blob_t, 0)
            torch.cuda.synchronize()
            self.bcast_s = time.perf_counter() - tb
            if rank != 0:
                self.gpu.genome_import(blob_t.data_ptr(), nbytes, take_ownership=False)
            self.keep = blob_t
        self.setup_s = time.perf_counter() - t0
        self.total_pairs = self.gpu.genome_pairs()
        self.B = PAIRS_PER_BATCH
        self.n_avail = self.total_pairs // (self.B * world)
        if self.n_avail < 1:
            self.B = int(self.total_pairs // world)
            self.n_avail = 1
        self.stream = torch.cuda.ExternalStream(self.gpu.cuda_stream(), device=torch.device("cuda", local_rank))
        self.rand_base = 0
        self.batch_no = 0
        # N > 1: the one exchange of the path -- random-pair counts of the round, so that rand_ii in the names stays the
        # global running count (src/dwgsim.c:1096) -- as an NCCL all-gather of device counters enqueued on the library's
        # stream between the simulate passes and the layout kernels: no host round trip inside a batch
        if world > 1:
            self.allc = torch.zeros(world, dtype=torch.int64, device="cuda")
            self.base_t = torch.zeros(1, dtype=torch.int64, device="cuda")
            self.running_t = torch.zeros(1, dtype=torch.int64, device="cuda")

    def enqueue(self):
        """one device batch queued behind the previous ones on the library's stream: no host wait (rand_ii continues in device
        memory: the library's running counter at N = 1, the all-gathered counts at N > 1)"""
        torch, dist, gpu = self.torch, self.dist, self.gpu
        first = ((self.batch_no % self.n_avail) * self.world + self.rank) * self.B
        self.batch_no += 1
        if self.world > 1:
            gpu.resident_begin(first, self.B, False)
            cnt = torch.as_tensor(_CudaArray(gpu.resident_count_ptr(), 8), device="cuda").view(torch.int64)
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(self.allc, cnt)
                gpu.resident_finish_gathered(self.allc.data_ptr(), self.world, self.rank)   # prefix of the counts: one kernel of the library
        else:
            gpu.resident_enqueue(first, self.B)

    def batch(self):
        """one device batch, waited for"""
        self.enqueue()
        return self.gpu.resident_wait()

    def measure(self, steps, warmup, min_seconds, sampler=None):
        """`warmup` untimed steps, then exactly `steps` timed steps of `per_step` device batches each"""
        torch, dist, world = self.torch, self.dist, self.world
        # calibration: device batches per step so that the timed region lasts min_seconds
        for _ in range(3):
            self.batch()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(8):
            self.enqueue()
        self.gpu.resident_wait()
        torch.cuda.synchronize()
        t_batch = (time.perf_counter() - t0) / 8
        if world > 1:
            t = torch.tensor([t_batch], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_batch = float(t.item())
        per_step = max(1, int(1.05 * min_seconds / (steps * t_batch) + 0.999))     # (the calibration carries some start-up time)
        for _ in range(warmup):
            for _ in range(per_step):
                self.batch()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        ms = [0.0, 0.0, 0.0]
        out_bytes = launches = sampled = 0
        t_wall0 = time.perf_counter()
        for _ in range(steps):
            # the batches of a step are queued back to back; one host wait per step.  The kernel times and sizes of the step's
            # last batch are the sample of the per-kernel figures below
            for _ in range(per_step):
                self.enqueue()
            b = self.gpu.resident_wait()
            ms[0] += b.ms_simulate; ms[1] += b.ms_layout; ms[2] += b.ms_format
            out_bytes += sum(b.n_bytes)
            launches += b.n_launches
            sampled += 1
        ev1.record(self.stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop() if sampler else None
        elapsed_ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed_ms = float(t.item())
        n_batches = steps * per_step
        units = world * self.B * n_batches
        # roofline of the kernels (device time from CUDA events inside the library, per launch group; this rank's)
        peak, peak_src = peak_hbm()
        L = self.w["opts"]["length"]
        read_bytes = sum(((x + 3) // 4) * 1.5 for x in L if x > 0) + 5.0
        fastq_per_unit = out_bytes / (self.B * sampled)
        algo = ALGO_BYTES_PER_PAIR if self.name == "illumina_2x150" else read_bytes + fastq_per_unit
        kern_ms = sum(ms) / sampled
        achieved = algo * self.B / (kern_ms * 1e-3) / 1e9
        names = ["simulate_pairs_tp_kernel", "layout_* (5 scan kernels)", "format_fastq_kernel"]
        dom = max(range(3), key=lambda i: ms[i])
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src, "algorithmic_bytes_per_unit": algo, "fastq_bytes_per_unit": fastq_per_unit,
                    "kernel": "whole device batch = simulate + layout + format; dominant: %s" % names[dom],
                    "ms_per_batch_by_kernel": {n: m / sampled for n, m in zip(names, ms)},
                    "sampled_batches": sampled}
        fmt_ms = ms[2] / sampled
        if fmt_ms > 0:        # the formatter alone: its algorithmic bytes are the FASTQ text it writes
            fmt_bytes = out_bytes / sampled
            roofline["dominant_kernel" if dom == 2 else "format_kernel"] = {
                "name": "format_fastq_kernel", "ms": fmt_ms, "algorithmic_bytes": fmt_bytes,
                "achieved": fmt_bytes / (fmt_ms * 1e-3) / 1e9, "unit": "GB/s", "frac": fmt_bytes / (fmt_ms * 1e-3) / 1e9 / peak}
        sim_ms = ms[0] / sampled
        if dom == 0 and sim_ms > 0:
            roofline["dominant_kernel"] = {"name": "simulate_pairs_tp_kernel", "ms": sim_ms,
                                           "algorithmic_bytes": algo * self.B, "achieved": algo * self.B / (sim_ms * 1e-3) / 1e9,
                                           "unit": "GB/s", "frac": algo * self.B / (sim_ms * 1e-3) / 1e9 / peak,
                                           "note": "measured against the whole batch's algorithmic bytes"}
        return {"value": units / (elapsed_ms * 1e-3), "unit": self.w["unit"], "elapsed_ms": elapsed_ms, "ms_per_step": elapsed_ms / steps,
                "device_batches_per_step": per_step, "pairs_per_step": world * self.B * per_step, "timed_s": elapsed_ms * 1e-3,
                "wall_s_timed_region": t_wall, "gpu_launches": launches, "roofline": roofline, "clocks": clocks}

    def close(self):
        self.gpu.close()
        self.keep = None


def e2e_legs(args, rank, local_rank, world):
    """The C ABI with host buffers.  Per step: dwgsim_gpu_add_contig (dense seq_t + 2 x mut_t[len] host arrays, packed on
    the host, copied to the device) + dwgsim_gpu_run.  Sinks: "memory" = the bytes are delivered in pair order to host
    (pinned) memory and counted (dwgsim_gpu_sink_count: what a consumer in the same process sees); "files" = they are
    also written to three NEW files on tmpfs per step (dwgsim_gpu_sink_files, one writer thread per file), removed by a
    helper thread afterwards -- that leg is bounded by the kernel's page-cache copy (a few GB/s per file)."""
    import threading
    import torch
    import torch.distributed as dist
    from dwgsim_b200 import DwgsimGpu, params_from_options
    seq, hap = dense_contig(E2E_CONTIG_LEN, 7 + rank)
    n_pairs_c = int(E2E_CONTIG_LEN * COVERAGE / 300.0 / 0.95 + 0.5)
    n_warm = 6       # the first steps allocate the pinned ring and first-touch its pages (100+ ms stalls on a fresh box)
    n_e2e = max(5, min(args.steps, 20))
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None

    def leg(compress, files):
        wd = tempfile.mkdtemp(prefix="dwgsim_e2e_r%d_" % rank, dir=base) if files else None
        g2 = DwgsimGpu(params_from_options(**OPTS), device=local_rank)
        g2.set_batch(1 << 18, 3)
        if compress:
            g2.set_compression(1)
        doomed, lock, stop = [], threading.Lock(), [False]

        def reaper():
            while True:
                with lock:
                    paths = doomed[:]
                    del doomed[:]
                for q in paths:
                    try:
                        os.unlink(q)
                    except OSError:
                        pass
                if stop[0] and not paths:
                    return
                time.sleep(0.002)

        th = threading.Thread(target=reaper, daemon=True)
        th.start()

        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(1)
        g2.set_host_threads(max(1, min(32, (os.cpu_count() or 1) // max(world, 1))))

        def pack(i):           # the host half of add_contig; the next step's runs while this step's batches are on the device
            return g2.pack_contig(i, "chrE%d" % i, seq.ctypes.data, E2E_CONTIG_LEN, hap[0].ctypes.data, hap[1].ctypes.data,
                                  None, 0, None, 0, n_pairs_c)

        nxt = [pool.submit(pack, 0)]

        def one(i):
            g2.add_packed(nxt[0].result())
            nxt[0] = pool.submit(pack, i + 1)
            if not files:
                st = g2.run_count()
                return st, sum(st.bytes)
            names = [os.path.join(wd, "s%d.%s%s" % (i, f, ".gz" if compress else "")) for f in
                     ("bwa.read1.fastq", "bwa.read2.fastq", "bfast.fastq")]
            fds = [os.open(q, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644) for q in names]
            try:
                st, offs = g2.run_to_files(fds)
            finally:
                for fd in fds:
                    os.close(fd)
            with lock:
                doomed.extend(names)
            return st, sum(offs)

        try:
            for i in range(n_warm):
                one(i)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            te = time.perf_counter()
            h2d = d2h = raw = delivered = 0
            pack_ms = 0.0
            for i in range(n_e2e):
                st, nb = one(n_warm + i)
                h2d += st.h2d_bytes; d2h += st.d2h_bytes; pack_ms += st.ms_pack; raw += sum(st.raw_bytes); delivered += nb
            torch.cuda.synchronize()
            secs = time.perf_counter() - te
            if world > 1:
                t = torch.tensor([secs], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                secs = float(t.item())
        finally:
            stop[0] = True
            th.join(timeout=30)
            try:
                g2._L.dwgsim_gpu_packed_free(nxt[0].result())
            except Exception:
                pass
            pool.shutdown()
            g2.close()
            if wd:
                shutil.rmtree(wd, ignore_errors=True)
        out = {"value": world * n_pairs_c * n_e2e / secs, "unit": "pairs/s", "h2d_bytes_per_step": h2d // n_e2e,
               "d2h_bytes_per_step": d2h // n_e2e, "steps": n_e2e, "pairs_per_step": n_pairs_c, "ms_per_step": 1e3 * secs / n_e2e,
               "host_pack_ms_per_step": pack_ms / n_e2e, "fastq_bytes_per_step": raw // n_e2e, "sink_bytes_per_step": delivered // n_e2e}
        what = ("dwgsim_gpu_pack_contig + add_packed (%d Mbp contig, dense host arrays; the next step's packing runs on a helper "
                "thread while this step's batches are on the device) + dwgsim_gpu_run -> %s of all three files in pair order, %s") % (
            E2E_CONTIG_LEN >> 20,
            ".fastq.gz bytes (gzip members written on the GPU, dwgsim_gpu_set_compression(1), the host shell's default)" if compress else "FASTQ text",
            "written to three new files on tmpfs per step (dwgsim_gpu_sink_files, one writer thread per file)" if files else
            "delivered to host memory (library counting sink)")
        out["what"] = "per step: " + what
        return out

    e2e = leg(True, False)                      # headline: the drop-in's default output delivered to host memory
    e2e["compression_ratio"] = e2e["d2h_bytes_per_step"] / max(e2e["fastq_bytes_per_step"], 1)
    try:
        e2e["raw_sink"] = leg(False, False)
    except Exception as ex:
        e2e["raw_sink"] = {"error": str(ex)}
    try:
        e2e["file_sink"] = leg(True, True)
    except Exception as ex:
        e2e["file_sink"] = {"error": str(ex)}
    return e2e


def write_genome_fasta(path, scale=1.0):
    """the 24-contig synthetic reference of configs[1] as a FASTA + .fai (i.i.d. ACGT, 10 kb telomeres and one long N run)"""
    import numpy as np
    rng = np.random.default_rng(20261017)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    total = 0
    with open(path, "wb") as f, open(path + ".fai", "w") as fai:
        off = 0
        for i, full in enumerate(GRCH38):
            n = max(int(full * scale), 200000)
            name = "chr%d" % (i + 1)
            s = acgt[np.frombuffer(rng.bytes(n), dtype=np.uint8) & 3]
            s[:10000] = ord("N"); s[n - 10000:] = ord("N")
            s[n // 3:n // 3 + int(n * N_FRAC)] = ord("N")
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
            total += n
    return total


def e2e_cli(args, cpu_baseline):
    """wall time of the drop-in binary on the whole synthetic reference: FASTA (+ .fai) in, -C 30, the three .fastq.gz streams to
    /dev/null-backed files (symlinks), the mutation files to tmpfs; host prologue (mut_diref, mutation files, packing) included"""
    import re
    from dwgsim_b200 import build
    exe = build.build_cli()
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    wd = tempfile.mkdtemp(prefix="dwgsim_cli_", dir=base)
    try:
        fa = os.path.join(wd, "genome.fa")
        t0 = time.perf_counter()
        bases = write_genome_fasta(fa, args.genome_scale)
        t_fa = time.perf_counter() - t0
        prefix = os.path.join(wd, "out")
        for f in ("bwa.read1.fastq.gz", "bwa.read2.fastq.gz", "bfast.fastq.gz"):
            os.symlink("/dev/null", prefix + "." + f)
        argv = [exe] + REF_ARGV + ["-C", "%g" % COVERAGE, "-z", "1", fa, prefix]
        env = dict(os.environ, DWGSIM_STATS="1")
        t0 = time.perf_counter()
        r = subprocess.run(argv, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": r.stderr[-400:]}
        m = re.search(r"\[dwgsim_b200\] bases (\d+) pairs (\d+) fastq_bytes (\d+) \| total ([\d.]+) s: mut_diref ([\d.]+) s .*?mut_print ([\d.]+) s, "
                      r"read loop ([\d.]+) s \(host pack ([\d.]+) s, kernels ([\d.]+) s", r.stderr)
        out = {"wall_s": wall, "genome_bases": bases, "fasta_write_s": t_fa, "gpus": 1,
               "what": "dwgsim_b200/bin/dwgsim %s on the 24-contig synthetic reference (FASTA + .fai on tmpfs); .fastq.gz streams "
                       "(device gzip) to /dev/null-backed files, mutations.txt/.vcf to tmpfs; wall time of the process" % " ".join(argv[1:-2])}
        if m:
            pairs = int(m.group(2))
            out.update(pairs=pairs, pairs_per_s=pairs / wall, fastq_bytes=int(m.group(3)), mut_diref_s=float(m.group(5)),
                       mut_print_s=float(m.group(6)), read_loop_s=float(m.group(7)), host_pack_s=float(m.group(8)),
                       kernels_s=float(m.group(9)))
            if cpu_baseline and cpu_baseline.get("value"):
                ref_prologue = cpu_baseline.get("prologue_ns_per_base", 0.0) * 1e-9 * bases
                out["reference_extrapolated_s"] = pairs / cpu_baseline["value"] + ref_prologue
                out["reference_extrapolation"] = "pairs / (1-core reference rate of cpu_baseline) + measured reference prologue ns/base x bases"
        return out
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-cli", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip configs[2..4]")
    ap.add_argument("--only", default=None, help="measure this workload as the headline (debug / profiling)")
    ap.add_argument("--genome-scale", type=float, default=1.0, help="shrink the synthetic genome (debug only)")
    ap.add_argument("--min-seconds", type=float, default=MIN_TIMED_S, help="lower bound of the timed region (profiling: 0 = one device batch per step)")
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
    from dwgsim_b200 import build
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

    # ---- headline: configs[1] on the kernel path -------------------------------------------------------------------
    head_name = args.only or "illumina_2x150"
    kp = KernelPath(head_name, rank, local_rank, world, args.genome_scale)
    head = kp.measure(args.steps, args.warmup, args.min_seconds, ClockSampler(local_rank))
    setup = {"genome_build_s": kp.setup_s, "nccl_broadcast_s": kp.bcast_s, "genome_pairs": kp.total_pairs,
             "wall_s_timed_region": head["wall_s_timed_region"], "timed_s": head["timed_s"]}
    kp.close()
    tr = traffic_record()
    roofline = head["roofline"]
    roofline["traffic"] = tr["dram_bytes_per_batch"] if tr and head_name == "illumina_2x150" else None
    roofline["traffic_source"] = tr.get("source") if tr else None

    # ---- configs[2..4] ---------------------------------------------------------------------------------------------------
    configs = {}
    if not args.no_configs and not args.only:
        for name in ("illumina_2x150_R0.15", "solid_2x50", "iontorrent_400se"):
            try:
                k2 = KernelPath(name, rank, local_rank, world, args.genome_scale)
                r = k2.measure(max(5, min(args.steps, 20)), 3, MIN_TIMED_S_OTHER)
                k2.close()
                configs[name] = {"workload": WORKLOADS[name]["label"] + ", 3.1 Gbp synthetic reference, -o 0", "value": r["value"],
                                 "unit": r["unit"], "n_gpus": world, "timed_s": r["timed_s"], "pairs_per_step": r["pairs_per_step"],
                                 "roofline": {k: r["roofline"][k] for k in ("achieved", "peak", "unit", "frac", "algorithmic_bytes_per_unit",
                                                                            "fastq_bytes_per_unit", "ms_per_batch_by_kernel")}}
            except Exception as ex:
                configs[name] = {"error": str(ex)}

    # ---- e2e: the C ABI with host buffers, output to files on tmpfs -----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = e2e_legs(args, rank, local_rank, world)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and ref_binary():
        try:
            n = 60000
            r = cpu_reference(1, n, 1, 1)
            cpu_baseline = {"value": r["value"], "unit": "pairs/s", "cores": 1, "kind": "reference",
                            "prologue_ns_per_base": r["prologue_s"] / 2_000_000 * 1e9,
                            "sample": "oracle/_ref/dwgsim_ref -N %d on a 2 Mbp sample of the synthetic reference, same options, "
                                      "all three .fastq.gz outputs, 1 process; prologue (%.2f s, -C 0) subtracted" % (n, r["prologue_s"])}
        except Exception as e:
            cpu_baseline = {"value": None, "unit": "pairs/s", "cores": 1, "kind": "reference", "sample": "failed: %s" % e}

    cli = None
    if rank == 0 and world == 1 and not args.no_e2e_cli and not args.only:
        try:
            cli = e2e_cli(args, cpu_baseline)
        except Exception as ex:
            cli = {"error": str(ex)}

    if rank == 0:
        cfg = config_dict(world)
        if args.only:
            cfg["workload"] = WORKLOADS[head_name]["label"]
        line = {"metric": "read-pairs/sec (2x150bp, 3.1Gbp ref)", "value": head["value"], "unit": head["unit"], "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
                "step": {"device_batches_per_step": head["device_batches_per_step"], "pairs_per_step": head["pairs_per_step"],
                         "timed_s": head["timed_s"]},
                "clocks": head["clocks"], "e2e": e2e, "e2e_cli": cli, "gpu_launches": head["gpu_launches"], "roofline": roofline,
                "configs": configs, "cpu_baseline": cpu_baseline, "setup": setup}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
