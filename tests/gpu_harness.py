"""Shared helpers of the -m gpu parity tests: play the role of the reference host at the seam.

The oracle (Philox backend) produces the expected FASTQ bytes AND, with keep=True, the dense seq_t /
mutseq_t arrays a reference host would hold after mut_diref; those arrays are handed to the C ABI exactly
as dwgsim_core would (include/dwgsim_gpu.h), and the product's bytes must equal the oracle's.
"""
import os

FILE_NAMES = ["bwa.read1.fastq", "bwa.read2.fastq", "bfast.fastq"]
GPU_KEYS = ("e", "E", "is_inner", "dist", "std_dev", "length", "mut_freq", "rand_read", "max_n", "data_type",
            "strandedness", "read_one_strand", "flow_order", "seed", "fixed_quality", "quality_std", "read_prefix",
            "reads_output_type", "amplicons")


def oracle_expected(oracle, opts, fasta, prefix):
    """run the oracle with Philox draws; returns (session kept open, [bytes x3])"""
    opt = oracle.make_opt(**opts)
    sess = oracle.Session(opt, fasta, prefix, mode=oracle.RNG_PHILOX, keep=True)
    sess.opt = opt
    want = []
    for f in FILE_NAMES:
        p = prefix + "." + f
        want.append(open(p, "rb").read() if os.path.exists(p) else b"")
    return sess, want


def gpu_actual(sess, opts, batch=None, per_contig_runs=False, orc_opt=None, compression=0, devices=None):
    from dwgsim_b200 import DwgsimGpu, params_from_options
    params = params_from_options(**{k: v for k, v in opts.items() if k in GPU_KEYS})
    if orc_opt is not None:
        # -B (Ion Torrent base-error calibration) rescales e in option parsing (src/dwgsim_opt.c:415-457), i.e. on the
        # host before the seam: hand the device the calibrated profile the host holds
        for i in range(2):
            params.e_start[i], params.e_by[i] = orc_opt.e_start[i], orc_opt.e_by[i]
    got = [[], [], []]
    stats = []
    with DwgsimGpu(params, devices=devices) as gpu:
        if batch:
            gpu.set_batch(batch, 2)
        if compression:
            gpu.set_compression(compression)
        for k in range(sess.n_contigs):
            c = sess.contig(k)
            if c["n_pairs"] == 0 and not per_contig_runs:
                pass
            gpu.add_contig(c["contig_i"], c["name"], c["seq"], c["len"], c["hap"][0], c["hap"][1],
                           c["ins"][0], c["n_ins"][0], c["ins"][1], c["n_ins"][1], c["n_pairs"])
            if opts.get("fn_regions_bed"):
                gpu.set_regions(c["regions"], c["sample_len"])
            if per_contig_runs:
                stats.append(gpu.run(lambda fid, data: got[fid].append(data)))
        if not per_contig_runs:
            stats.append(gpu.run(lambda fid, data: got[fid].append(data)))
    out = [b"".join(g) for g in got]
    if compression:
        import gzip
        out = [gzip.decompress(x) if x else b"" for x in out]
    return out, stats


def first_diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            lo = a.rfind(b"\n@", 0, i) + 1
            return "byte %d: want %r / got %r" % (i, a[lo:i + 80], b[lo:i + 80])
    return "lengths differ: want %d got %d; tail want %r got %r" % (len(a), len(b), a[n - 60:n + 60], b[n - 60:n + 60])
