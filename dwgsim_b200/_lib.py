"""ctypes loader of libdwgsim_b200.so.  Fails loudly: there is no Python / CPU fallback for the path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("DWGSIM_LIB") or os.path.join(HERE, "libdwgsim_b200.so")   # (DWGSIM_LIB: kernel-variant experiments)


class Params(C.Structure):
    """dwgsim_gpu_params_t (include/dwgsim_gpu.h)"""
    _fields_ = [
        ("e_start", C.c_double * 2), ("e_by", C.c_double * 2),
        ("is_inner", C.c_int32), ("dist", C.c_int32), ("std_dev", C.c_double),
        ("length", C.c_int32 * 2), ("mut_freq", C.c_double), ("rand_read", C.c_double),
        ("max_n", C.c_int32), ("data_type", C.c_int32), ("strandedness", C.c_int32),
        ("read_one_strand", C.c_int32), ("flow_order", C.c_void_p), ("flow_order_len", C.c_int32),
        ("seed", C.c_int32), ("fixed_quality", C.c_int32), ("quality_std", C.c_double),
        ("read_prefix", C.c_char_p), ("reads_output_type", C.c_int32), ("amplicons", C.c_int32),
    ]


class Stats(C.Structure):
    """dwgsim_gpu_stats_t"""
    _fields_ = [
        ("n_pairs", C.c_int64), ("n_random", C.c_int64), ("n_failed_attempts", C.c_int64),
        ("bytes", C.c_int64 * 3), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("ms_simulate", C.c_double), ("ms_layout", C.c_double), ("ms_format", C.c_double),
        ("ms_pack", C.c_double), ("ms_total", C.c_double), ("n_launches", C.c_int32), ("n_batches", C.c_int32),
        ("raw_bytes", C.c_int64 * 3), ("ms_compress", C.c_double),
    ]


class Batch(C.Structure):
    """dwgsim_gpu_batch_t"""
    _fields_ = [
        ("dev_ptr", C.c_uint64 * 3), ("n_bytes", C.c_uint64 * 3),
        ("n_pairs", C.c_int64), ("n_random", C.c_int64), ("n_failed_attempts", C.c_int64),
        ("ms_simulate", C.c_double), ("ms_layout", C.c_double), ("ms_format", C.c_double),
        ("n_launches", C.c_int32),
    ]


class Tables(C.Structure):
    """dwgsim_gpu_tables_t"""
    _fields_ = [
        ("thr_genomic", C.c_uint64), ("thr_hap0", C.c_uint64),
        ("isize_lo", C.c_int32), ("isize_n", C.c_int32), ("isize_cdf", C.POINTER(C.c_uint32)),
        ("qdelta_lo", C.c_int32), ("qdelta_n", C.c_int32), ("qdelta_cdf", C.POINTER(C.c_uint32)),
        ("n_cycles", C.c_int32 * 2), ("err_gap", C.POINTER(C.c_uint32) * 2), ("err_acc", C.POINTER(C.c_uint32) * 2),
        ("qbase", C.POINTER(C.c_uint8) * 2), ("flow_thr", C.c_uint32 * 2),
        ("flow_gap", C.POINTER(C.c_uint32) * 2), ("flow_gap_n", C.c_int32 * 2),
    ]


SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_char), C.c_size_t)
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64))

# every symbol include/dwgsim_gpu.h declares: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "dwgsim_gpu_abi_version": (C.c_int, []),
    "dwgsim_gpu_create": (C.c_int, [C.POINTER(_P), C.POINTER(Params), C.c_int]),
    "dwgsim_gpu_create_group": (C.c_int, [C.POINTER(_P), C.POINTER(Params), C.POINTER(C.c_int32), C.c_int32]),
    "dwgsim_gpu_group_size": (C.c_int, [_P]),
    "dwgsim_gpu_destroy": (None, [_P]),
    "dwgsim_gpu_strerror": (C.c_char_p, [C.c_int]),
    "dwgsim_gpu_last_error": (C.c_char_p, [_P]),
    "dwgsim_gpu_add_contig": (C.c_int, [_P, C.c_int32, C.c_char_p, _P, C.c_int32, _P, _P, _P, C.c_int32, _P, C.c_int32,
                                        C.c_int64]),
    "dwgsim_gpu_pack_contig": (C.c_int, [_P, C.c_int32, C.c_char_p, _P, C.c_int32, _P, _P, _P, C.c_int32, _P, C.c_int32,
                                         C.c_int64, C.POINTER(_P)]),
    "dwgsim_gpu_add_packed": (C.c_int, [_P, _P]),
    "dwgsim_gpu_packed_free": (None, [_P]),
    "dwgsim_gpu_set_host_threads": (C.c_int, [_P, C.c_int32]),
    "dwgsim_gpu_warm": (C.c_int, [_P]),
    "dwgsim_gpu_set_regions": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int32, C.c_int32]),
    "dwgsim_gpu_run": (C.c_int, [_P, SINK_FN, _P, C.POINTER(Stats)]),
    "dwgsim_gpu_set_batch": (C.c_int, [_P, C.c_int64, C.c_int32]),
    "dwgsim_gpu_set_compression": (C.c_int, [_P, C.c_int32]),
    "dwgsim_gpu_set_shard": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "dwgsim_gpu_set_exchange": (C.c_int, [_P, EXCHANGE_FN, _P]),
    "dwgsim_gpu_resident_begin": (C.c_int, [_P, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]),
    "dwgsim_gpu_resident_finish": (C.c_int, [_P, C.c_int64, C.POINTER(Batch)]),
    "dwgsim_gpu_resident_count_ptr": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "dwgsim_gpu_resident_finish_dev": (C.c_int, [_P, C.c_uint64, C.POINTER(Batch)]),
    "dwgsim_gpu_resident_set_running": (C.c_int, [_P, C.c_int64]),
    "dwgsim_gpu_resident_enqueue": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "dwgsim_gpu_resident_finish_async": (C.c_int, [_P, C.c_uint64]),
    "dwgsim_gpu_resident_finish_gathered": (C.c_int, [_P, C.c_uint64, C.c_int32, C.c_int32]),
    "dwgsim_gpu_resident_wait": (C.c_int, [_P, C.POINTER(Batch)]),
    "dwgsim_gpu_set_origin": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "dwgsim_gpu_genome_finalize": (C.c_int, [_P]),
    "dwgsim_gpu_genome_blob": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "dwgsim_gpu_genome_import": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_int32]),
    "dwgsim_gpu_genome_pairs": (C.c_int64, [_P]),
    "dwgsim_gpu_simulate_resident": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.POINTER(Batch)]),
    "dwgsim_gpu_copy_stream": (C.c_int, [_P, C.c_int, _P, C.c_uint64]),
    "dwgsim_gpu_genome_synthetic": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.c_uint64, C.c_double, C.c_double,
                                              C.c_double, C.c_double]),
    "dwgsim_gpu_cuda_stream": (_P, [_P]),
    "dwgsim_gpu_gz_host_encode": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "dwgsim_gpu_sink_count": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "dwgsim_gpu_sink_fd": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "dwgsim_gpu_file_sink_open": (_P, [C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "dwgsim_gpu_sink_files": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "dwgsim_gpu_file_sink_close": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "dwgsim_gpu_pwrite_all": (C.c_int, [C.c_int, _P, C.c_size_t, C.c_int64]),
    "dwgsim_gpu_tables": (C.c_int, [_P, C.POINTER(Tables)]),
}

_LIB = None


def load():
    """dlopen the in-tree library (build it with `python -m dwgsim_b200.build`)"""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO):
            raise ImportError(
                "dwgsim_b200: %s is missing. Build it with `python -m dwgsim_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback for the read-pair path." % SO)
        lib = C.CDLL(SO)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB
