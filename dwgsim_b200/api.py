"""Host-side mirror of the C ABI (include/dwgsim_gpu.h) over ctypes.

Names and argument meaning follow the reference's dwgsim_opt_t (src/dwgsim_opt.h:21-60) and the call
order a reference maintainer would use at the seam of dwgsim_core (src/dwgsim.c:628-636):

    gpu = DwgsimGpu(params_from_options(length=(150, 150), e="0.001-0.01", ...))
    for each contig:   gpu.add_contig(contig_i, name, seq_ptr, len, hap1_ptr, hap2_ptr, ins..., n_pairs)
    gpu.run(sink)      # sink(file_id, bytes) in pair order
"""
import ctypes as C

from . import _lib
from ._lib import Batch, Params, Stats, Tables

FILE_BWA1, FILE_BWA2, FILE_BFAST = 0, 1, 2
ILLUMINA, SOLID, IONTORRENT = 0, 1, 2
_NT4 = {ord("A"): 0, ord("a"): 0, ord("C"): 1, ord("c"): 1, ord("G"): 2, ord("g"): 2, ord("T"): 3, ord("t"): 3}


class DwgsimGpuError(RuntimeError):
    def __init__(self, code, text):
        RuntimeError.__init__(self, "dwgsim_gpu error %d: %s" % (code, text))
        self.code = code


def _parse_rate(v):
    """-e/-E syntax of the reference: 'a', 'a-b' or 'a,b', split at the first '-' or ',' (src/dwgsim_opt.c:162-179)"""
    if isinstance(v, (tuple, list)):
        return float(v[0]), float(v[1])
    if not isinstance(v, str):
        return float(v), float(v)

    def atof(s):
        import re
        m = re.match(r"\s*[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)", s)
        return float(m.group(0)) if m else 0.0

    start = atof(v)
    i = next((i for i, ch in enumerate(v) if ch in ",-"), len(v))
    return (start, atof(v[i + 1:])) if i < len(v) - 1 else (start, start)


def params_from_options(**kw):
    """dwgsim options by their dwgsim_opt_t names -> Params, applying what dwgsim_opt_parse applies
    (defaults src/dwgsim_opt.c:40-80, slope :459-460, flow order codes :404-407)."""
    o = dict(e=0.02, E=0.02, is_inner=0, dist=500, std_dev=50.0, length=(70, 70), mut_freq=0.5, rand_read=0.05,
             max_n=0, data_type=0, strandedness=0, read_one_strand=0, flow_order=None, seed=-1,
             fixed_quality=None, quality_std=2.0, read_prefix=None, reads_output_type=0, amplicons=0)
    for k, v in kw.items():
        if k not in o:
            raise KeyError(k)
        o[k] = v
    p = Params()
    p._keep = []
    for i, key in enumerate(("e", "E")):
        s, e = _parse_rate(o[key])
        p.e_start[i] = s
        p.e_by[i] = (e - s) / o["length"][i] if o["length"][i] else 0.0
    p.is_inner, p.dist, p.std_dev = int(o["is_inner"]), int(o["dist"]), float(o["std_dev"])
    p.length[0], p.length[1] = int(o["length"][0]), int(o["length"][1])
    p.mut_freq, p.rand_read, p.max_n = float(o["mut_freq"]), float(o["rand_read"]), int(o["max_n"])
    p.data_type, p.strandedness, p.read_one_strand = int(o["data_type"]), int(o["strandedness"]), int(o["read_one_strand"])
    if o["flow_order"]:
        codes = bytes(_NT4.get(ch, 4) for ch in o["flow_order"].encode())
        buf = C.create_string_buffer(codes, len(codes))
        p._keep.append(buf)
        p.flow_order = C.cast(buf, C.c_void_p)
        p.flow_order_len = len(codes)
    p.seed = int(o["seed"])
    fq = o["fixed_quality"]
    p.fixed_quality = 0 if not fq else (ord(fq) if isinstance(fq, str) else int(fq))
    p.quality_std = float(o["quality_std"])
    if o["read_prefix"] is not None:
        p.read_prefix = o["read_prefix"].encode() if isinstance(o["read_prefix"], str) else o["read_prefix"]
    p.reads_output_type, p.amplicons = int(o["reads_output_type"]), int(o["amplicons"])
    return p


class DwgsimGpu:
    """one dwgsim_gpu_t handle on one CUDA device, or (devices=[...]) on a group of devices of the box"""

    def __init__(self, params, device=0, devices=None):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self._params = params
        if devices is not None:
            arr = (C.c_int32 * len(devices))(*devices)
            rc = self._L.dwgsim_gpu_create_group(C.byref(self._h), C.byref(params), arr, len(devices))
        else:
            rc = self._L.dwgsim_gpu_create(C.byref(self._h), C.byref(params), device)
        if rc:
            self._h = None
            raise DwgsimGpuError(rc, self._L.dwgsim_gpu_strerror(rc).decode())

    def _check(self, rc):
        if rc:
            detail = self._L.dwgsim_gpu_last_error(self._h).decode()
            raise DwgsimGpuError(rc, self._L.dwgsim_gpu_strerror(rc).decode() + (": " + detail if detail else ""))

    def close(self):
        if self._h:
            self._L.dwgsim_gpu_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the seam ---------------------------------------------------------------------------------
    def add_contig(self, contig_i, name, seq, length, hap1, hap2, ins1=None, ins1_n=0, ins2=None, ins2_n=0, n_pairs=0):
        """seq / hap1 / hap2 / ins1 / ins2 are host addresses (ints or ctypes pointers) of the reference's
        seq_t.s, mutseq_t.s and mutseq_t.ins arrays"""
        name = name if isinstance(name, bytes) else name.encode()
        self._check(self._L.dwgsim_gpu_add_contig(self._h, contig_i, name, seq, length, hap1, hap2, ins1, ins1_n,
                                                  ins2, ins2_n, n_pairs))

    def pack_contig(self, contig_i, name, seq, length, hap1, hap2, ins1=None, ins1_n=0, ins2=None, ins2_n=0, n_pairs=0):
        """the host half of add_contig (may run on another thread while run() is in flight); returns an opaque handle"""
        name = name if isinstance(name, bytes) else name.encode()
        out = C.c_void_p()
        rc = self._L.dwgsim_gpu_pack_contig(self._h, contig_i, name, seq, length, hap1, hap2, ins1, ins1_n, ins2, ins2_n,
                                            n_pairs, C.byref(out))
        if rc:
            raise DwgsimGpuError(rc, self._L.dwgsim_gpu_strerror(rc).decode())
        return out

    def add_packed(self, packed):
        """queue the result of pack_contig (ownership passes to the handle)"""
        self._check(self._L.dwgsim_gpu_add_packed(self._h, packed))

    def set_host_threads(self, n):
        self._check(self._L.dwgsim_gpu_set_host_threads(self._h, n))

    def set_regions(self, regions, sample_len):
        """-x for the contig just queued: regions = [(start, end), ...] merged and sorted (BED half-open),
        sample_len = the reference's `l` after src/dwgsim.c:539-553"""
        n = len(regions)
        a = (C.c_uint32 * max(n, 1))(*[r[0] for r in regions])
        b = (C.c_uint32 * max(n, 1))(*[r[1] for r in regions])
        self._check(self._L.dwgsim_gpu_set_regions(self._h, a, b, n, sample_len))

    def run(self, sink):
        """sink(file_id:int, data:bytes) is called in pair order; returns Stats"""
        err = []

        def _cb(user, file_id, buf, n):
            try:
                sink(file_id, C.string_at(buf, n))
                return 0
            except Exception as e:  # propagate after the C call returns
                err.append(e)
                return 1

        st = Stats()
        rc = self._L.dwgsim_gpu_run(self._h, _lib.SINK_FN(_cb), None, C.byref(st))
        if err:
            raise err[0]
        self._check(rc)
        return st

    def run_count(self):
        """run() with the library's own counting sink (no Python in the data path); returns Stats"""
        counts = (C.c_int64 * 4)()
        st = Stats()
        cb = C.cast(self._L.dwgsim_gpu_sink_count, _lib.SINK_FN)
        self._check(self._L.dwgsim_gpu_run(self._h, cb, C.cast(counts, C.c_void_p), C.byref(st)))
        assert list(counts[:3]) == list(st.bytes)
        return st

    def run_to_fds(self, fds):
        """run() writing each stream to a file descriptor (-1 discards); returns Stats"""
        arr = (C.c_int * 3)(*fds)
        st = Stats()
        cb = C.cast(self._L.dwgsim_gpu_sink_fd, _lib.SINK_FN)
        self._check(self._L.dwgsim_gpu_run(self._h, cb, C.cast(arr, C.c_void_p), C.byref(st)))
        return st

    def run_to_files(self, fds, offsets=(0, 0, 0)):
        """run() through the library's file sink (one background writer per file, positional writes); returns
        (Stats, end offsets)"""
        fs = self._L.dwgsim_gpu_file_sink_open((C.c_int32 * 3)(*fds), (C.c_int64 * 3)(*offsets))
        st = Stats()
        cb = C.cast(self._L.dwgsim_gpu_sink_files, _lib.SINK_FN)
        rc = self._L.dwgsim_gpu_run(self._h, cb, fs, C.byref(st))
        end = (C.c_int64 * 3)()
        wrc = self._L.dwgsim_gpu_file_sink_close(fs, end)
        self._check(rc)
        if wrc:
            raise DwgsimGpuError(-7, "writing the output files failed")
        return st, list(end)

    def run_collect(self):
        """run() into three bytes objects (tests)"""
        parts = ([], [], [])
        st = self.run(lambda fid, data: parts[fid].append(data))
        return [b"".join(p) for p in parts], st

    # -- knobs ------------------------------------------------------------------------------------
    def set_batch(self, pairs_per_batch, ring_slots=3):
        self._check(self._L.dwgsim_gpu_set_batch(self._h, pairs_per_batch, ring_slots))

    def set_compression(self, mode):
        """0: FASTQ text to the sink; 1: gzip members written on the device"""
        self._check(self._L.dwgsim_gpu_set_compression(self._h, mode))

    def set_shard(self, rank, world):
        self._check(self._L.dwgsim_gpu_set_shard(self._h, rank, world))

    def set_exchange(self, fn):
        """fn(round, my_random) -> (before_me, round_total); called once per round on every rank (a collective)"""
        def _cb(user, rnd, mine, before, total):
            try:
                b, t = fn(int(rnd), int(mine))
                before[0], total[0] = int(b), int(t)
                return 0
            except Exception as e:
                self._exchange_error = e
                return 1
        self._exchange_cb = _lib.EXCHANGE_FN(_cb)      # keep alive
        self._check(self._L.dwgsim_gpu_set_exchange(self._h, self._exchange_cb, None))

    def set_origin(self, first_pair_index, first_rand_serial=0):
        self._check(self._L.dwgsim_gpu_set_origin(self._h, first_pair_index, first_rand_serial))

    # -- device-resident interface ------------------------------------------------------------------
    def genome_finalize(self):
        self._check(self._L.dwgsim_gpu_genome_finalize(self._h))

    def genome_blob(self):
        p, n = C.c_uint64(), C.c_uint64()
        self._check(self._L.dwgsim_gpu_genome_blob(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def genome_import(self, device_ptr, n_bytes, take_ownership=False):
        self._check(self._L.dwgsim_gpu_genome_import(self._h, device_ptr, n_bytes, 1 if take_ownership else 0))

    def genome_pairs(self):
        return self._L.dwgsim_gpu_genome_pairs(self._h)

    def genome_synthetic(self, lengths, seed, mut_rate, indel_frac, n_frac, coverage):
        arr = (C.c_int32 * len(lengths))(*lengths)
        self._check(self._L.dwgsim_gpu_genome_synthetic(self._h, len(lengths), arr, seed, mut_rate, indel_frac, n_frac,
                                                        coverage))

    def simulate_resident(self, first, n, rand_serial_base=0):
        b = Batch()
        self._check(self._L.dwgsim_gpu_simulate_resident(self._h, first, n, rand_serial_base, C.byref(b)))
        return b

    def resident_begin(self, first, n, want_count=True):
        c = C.c_int64()
        self._check(self._L.dwgsim_gpu_resident_begin(self._h, first, n, C.byref(c) if want_count else None))
        return c.value

    def resident_finish(self, rand_serial_base):
        b = Batch()
        self._check(self._L.dwgsim_gpu_resident_finish(self._h, rand_serial_base, C.byref(b)))
        return b

    def resident_count_ptr(self):
        """device address of the uint64 random-pair count of the batch begun last"""
        p = C.c_uint64()
        self._check(self._L.dwgsim_gpu_resident_count_ptr(self._h, C.byref(p)))
        return p.value

    def resident_finish_dev(self, rand_serial_base_device_ptr):
        b = Batch()
        self._check(self._L.dwgsim_gpu_resident_finish_dev(self._h, rand_serial_base_device_ptr, C.byref(b)))
        return b

    def resident_set_running(self, rand_serial):
        self._check(self._L.dwgsim_gpu_resident_set_running(self._h, rand_serial))

    def resident_enqueue(self, first, n):
        """queue one batch behind the previous ones (no host sync); rand_ii continues in device memory"""
        self._check(self._L.dwgsim_gpu_resident_enqueue(self._h, first, n))

    def resident_finish_async(self, rand_serial_base_device_ptr):
        self._check(self._L.dwgsim_gpu_resident_finish_async(self._h, rand_serial_base_device_ptr))

    def resident_finish_gathered(self, counts_device_ptr, world, rank):
        self._check(self._L.dwgsim_gpu_resident_finish_gathered(self._h, counts_device_ptr, world, rank))

    def resident_wait(self):
        """wait for the queued batches; describes the last one"""
        b = Batch()
        self._check(self._L.dwgsim_gpu_resident_wait(self._h, C.byref(b)))
        return b

    def copy_stream(self, file_id, n_bytes):
        buf = C.create_string_buffer(max(int(n_bytes), 1))
        self._check(self._L.dwgsim_gpu_copy_stream(self._h, file_id, buf, n_bytes))
        return buf.raw[:n_bytes]

    def cuda_stream(self):
        return self._L.dwgsim_gpu_cuda_stream(self._h)

    def tables(self):
        t = Tables()
        self._check(self._L.dwgsim_gpu_tables(self._h, C.byref(t)))
        return t
