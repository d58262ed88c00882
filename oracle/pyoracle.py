"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (dwgsim_b200) never does.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RNG_DRAND48, RNG_PHILOX = 0, 1
ILLUMINA, SOLID, IONTORRENT = 0, 1, 2


class OrcOpt(C.Structure):
    _fields_ = [
        ("e_start", C.c_double * 2), ("e_end", C.c_double * 2), ("e_by", C.c_double * 2),
        ("is_inner", C.c_int32), ("dist", C.c_int32), ("std_dev", C.c_double),
        ("N", C.c_int64), ("C", C.c_double), ("length", C.c_int32 * 2),
        ("mut_rate", C.c_double), ("mut_freq", C.c_double), ("indel_frac", C.c_double),
        ("indel_extend", C.c_double), ("indel_min", C.c_int32), ("rand_read", C.c_double),
        ("max_n", C.c_int32), ("data_type", C.c_int32), ("strandedness", C.c_int32),
        ("read_one_strand", C.c_int32), ("flow_order_len", C.c_int32),
        ("flow_order", C.c_int8 * 1024), ("use_base_error", C.c_int32), ("is_hap", C.c_int32),
        ("seed", C.c_int32), ("fixed_quality", C.c_int32), ("quality_std", C.c_double),
        ("has_read_prefix", C.c_int32), ("read_prefix", C.c_char * 256),
        ("reads_output_type", C.c_int32), ("output_type", C.c_int32), ("amplicons", C.c_int32),
        ("finalized", C.c_int32), ("fn_regions_bed", C.c_char * 1024),
        ("muts_input_type", C.c_int32), ("fn_muts_input", C.c_char * 1024),
    ]


class OrcTables(C.Structure):
    _fields_ = [
        ("thr_genomic", C.c_uint64), ("thr_hap0", C.c_uint64),
        ("isize_lo", C.c_int32), ("isize_n", C.c_int32), ("isize_cdf", C.POINTER(C.c_uint32)),
        ("qdelta_lo", C.c_int32), ("qdelta_n", C.c_int32), ("qdelta_cdf", C.POINTER(C.c_uint32)),
        ("n_cycles", C.c_int32 * 2), ("err_gap", C.POINTER(C.c_uint32) * 2), ("err_acc", C.POINTER(C.c_uint32) * 2),
        ("qbase", C.POINTER(C.c_uint8) * 2), ("flow_thr", C.c_uint32 * 2),
        ("flow_gap", C.POINTER(C.c_uint32) * 2),
    ]


FLOW_GAP_N = 4096


class OrcStats(C.Structure):
    _fields_ = [
        ("n_pairs_total", C.c_int64), ("n_random", C.c_int64), ("n_failed_attempts", C.c_int64),
        ("n_contigs", C.c_int64), ("n_contigs_skipped", C.c_int64),
        ("bytes_bwa1", C.c_int64), ("bytes_bwa2", C.c_int64), ("bytes_bfast", C.c_int64),
        ("error", C.c_int32),
    ]


def build(force=False):
    """compile liboracle.so (and oracle/_ref when /root/reference is present)"""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("dwgsim_oracle.c", "dwgsim_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    ref_bin = os.path.join(_HERE, "_ref", "dwgsim_ref")
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(ref_bin)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    # the reference with the binding of INTEGRATION.md in place of its read-pair loop, linked against libdwgsim_b200.so
    gpu_bin = os.path.join(_HERE, "_ref", "dwgsim_ref_gpu")
    lib_so = os.path.join(os.path.dirname(_HERE), "dwgsim_b200", "libdwgsim_b200.so")
    deps = [os.path.join(os.path.dirname(_HERE), "integration", f) for f in ("dwgsim_b200_binding.c", "dwgsim_b200_binding.h")] + \
           [os.path.join(_HERE, "patch_reference.py"), os.path.join(os.path.dirname(_HERE), "include", "dwgsim_gpu.h")]
    if os.path.isdir("/root/reference/src") and os.path.exists(lib_so) and \
            (force or not os.path.exists(gpu_bin) or any(os.path.getmtime(d) > os.path.getmtime(gpu_bin) for d in deps)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref_gpu"])
    return so


def ref_binary():
    p = os.path.join(_HERE, "_ref", "dwgsim_ref")
    return p if os.path.exists(p) else None


def ref_gpu_binary():
    """the reference built with integration/dwgsim_b200_binding.c in place of its read-pair loop (needs a GPU to run)"""
    p = os.path.join(_HERE, "_ref", "dwgsim_ref_gpu")
    return p if os.path.exists(p) else None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_opt_init.argtypes = [C.POINTER(OrcOpt)]
        L.orc_opt_finalize.argtypes = [C.POINTER(OrcOpt)]
        L.orc_opt_finalize.restype = C.c_int
        L.orc_run.argtypes = [C.POINTER(OrcOpt), C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        L.orc_run.restype = C.c_void_p
        L.orc_close.argtypes = [C.c_void_p]
        L.orc_stats.argtypes = [C.c_void_p]
        L.orc_stats.restype = C.POINTER(OrcStats)
        L.orc_n_contigs.argtypes = [C.c_void_p]
        L.orc_n_contigs.restype = C.c_int32
        for fn, rt in (("orc_contig_name", C.c_char_p), ("orc_contig_index", C.c_int32),
                       ("orc_contig_len", C.c_int32), ("orc_contig_n_pairs", C.c_int64),
                       ("orc_contig_seq", C.c_void_p)):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_int32]
            getattr(L, fn).restype = rt
        L.orc_contig_hap.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_contig_hap.restype = C.c_void_p
        L.orc_contig_n_ins.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_contig_n_ins.restype = C.c_int32
        L.orc_contig_ins.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_contig_ins.restype = C.c_void_p
        L.orc_contig_sample_len.argtypes = [C.c_void_p, C.c_int32]
        L.orc_contig_sample_len.restype = C.c_int32
        L.orc_contig_regions.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_uint32))]
        L.orc_contig_regions.restype = C.c_int32
        L.orc_tables_build.argtypes = [C.POINTER(OrcOpt)]
        L.orc_tables_build.restype = C.POINTER(OrcTables)
        L.orc_tables_free.argtypes = [C.POINTER(OrcTables)]
        L.orc_srand48.argtypes = [C.c_int32]
        L.orc_drand48.restype = C.c_double
        L.orc_drand48_state.restype = C.c_uint64
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_philox_draw.argtypes = [C.c_int32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_philox_draw.restype = C.c_uint32
        _LIB = L
    return _LIB


def make_opt(**kw):
    """dwgsim command-line options by their reference names -> finalized OrcOpt.

    e / E take the reference's 'a-b' / 'a,b' / 'a' strings or floats; finalize() seeds drand48, so call
    this immediately before run().
    """
    L = lib()
    o = OrcOpt()
    L.orc_opt_init(C.byref(o))

    def rate(v):
        if isinstance(v, str):
            a, b = C.c_double(), C.c_double()
            L.orc_parse_error_rate(v.encode(), C.byref(a), C.byref(b))
            return a.value, b.value
        if isinstance(v, (tuple, list)):
            return float(v[0]), float(v[1])
        return float(v), float(v)

    for k, v in kw.items():
        if k == "e":
            o.e_start[0], o.e_end[0] = rate(v)
        elif k == "E":
            o.e_start[1], o.e_end[1] = rate(v)
        elif k == "N":
            o.N, o.C = int(v), -1.0
        elif k == "C":
            o.C, o.N = float(v), -1
        elif k == "length":
            o.length[0], o.length[1] = int(v[0]), int(v[1])
        elif k == "flow_order":
            b = v.encode() if isinstance(v, str) else v
            for i, ch in enumerate(b):
                o.flow_order[i] = ch
            o.flow_order[len(b)] = 0
        elif k == "read_prefix":
            o.has_read_prefix = 1
            o.read_prefix = v.encode() if isinstance(v, str) else v
        elif k == "fixed_quality":
            o.fixed_quality = ord(v) if isinstance(v, str) else int(v)
        elif k == "fn_regions_bed":
            o.fn_regions_bed = v.encode() if isinstance(v, str) else v
        elif k in ("fn_muts_txt", "fn_muts_bed", "fn_muts_vcf"):        # -m / -b / -v
            o.fn_muts_input = v.encode() if isinstance(v, str) else v
            o.muts_input_type = {"fn_muts_bed": 0, "fn_muts_txt": 1, "fn_muts_vcf": 2}[k]
        else:
            if not hasattr(o, k):
                raise KeyError(k)
            setattr(o, k, v)
    if not L.orc_opt_finalize(C.byref(o)):
        raise ValueError("options rejected (the reference would print usage)")
    return o


class Session:
    """one oracle run; with keep=True it retains every simulated contig (seq + mut_t arrays)"""

    def __init__(self, opt, fasta, prefix=None, mode=RNG_DRAND48, keep=False):
        self._L = lib()
        self._h = self._L.orc_run(C.byref(opt), fasta.encode(), prefix.encode() if prefix else None,
                                  mode, 1 if keep else 0)
        self.stats = self._L.orc_stats(self._h).contents

    def close(self):
        if self._h:
            self._L.orc_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def n_contigs(self):
        return self._L.orc_n_contigs(self._h)

    def contig(self, k):
        L, h = self._L, self._h
        rs, re_ = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
        nr = L.orc_contig_regions(h, k, C.byref(rs), C.byref(re_))
        return dict(
            sample_len=L.orc_contig_sample_len(h, k),
            regions=[(rs[i], re_[i]) for i in range(nr)],
            name=L.orc_contig_name(h, k), contig_i=L.orc_contig_index(h, k), len=L.orc_contig_len(h, k),
            n_pairs=L.orc_contig_n_pairs(h, k), seq=L.orc_contig_seq(h, k),
            hap=[L.orc_contig_hap(h, k, 0), L.orc_contig_hap(h, k, 1)],
            n_ins=[L.orc_contig_n_ins(h, k, 0), L.orc_contig_n_ins(h, k, 1)],
            ins=[L.orc_contig_ins(h, k, 0), L.orc_contig_ins(h, k, 1)],
        )


def opt_to_ref_argv(**kw):
    """same keyword options -> argv for oracle/_ref/dwgsim_ref (used to pin the oracle)"""
    m = {"dist": "-d", "std_dev": "-s", "N": "-N", "C": "-C", "mut_rate": "-r", "mut_freq": "-F",
         "indel_frac": "-R", "indel_extend": "-X", "indel_min": "-I", "rand_read": "-y", "max_n": "-n",
         "data_type": "-c", "strandedness": "-S", "read_one_strand": "-A", "seed": "-z",
         "quality_std": "-Q", "reads_output_type": "-o", "output_type": "-M", "flow_order": "-f",
         "read_prefix": "-P", "fixed_quality": "-q", "e": "-e", "E": "-E", "fn_regions_bed": "-x",
         "fn_muts_txt": "-m", "fn_muts_bed": "-b", "fn_muts_vcf": "-v"}
    argv = []
    for k, v in kw.items():
        if k == "length":
            argv += ["-1", str(v[0]), "-2", str(v[1])]
        elif k in ("is_inner", "use_base_error", "is_hap", "amplicons"):
            if v:
                argv.append({"is_inner": "-i", "use_base_error": "-B", "is_hap": "-H", "amplicons": "-a"}[k])
        else:
            argv += [m[k], str(v)]
    return argv
