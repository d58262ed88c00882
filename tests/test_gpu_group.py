"""-m gpu: a device group (dwgsim_gpu_create_group: one handle, one host thread per device, batches handed to the sink
in order) must write the bytes of a single device.  The box the driver tests on has one GPU, so the ranks of the group
share cuda:0 there; with two or more devices visible the same cases also run over distinct devices."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import gpu_harness as gh  # noqa: E402
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu


def device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets.append([0, 1])
    if n >= 4:
        sets.append([0, 1, 2, 3])
    return sets


def check(oracle, opts, fasta, tmp_path, devices, **kw):
    opts = make_golden.materialize(opts, str(tmp_path))
    sess, want = gh.oracle_expected(oracle, opts, fasta, str(tmp_path / "orc"))
    try:
        got, stats = gh.gpu_actual(sess, opts, orc_opt=sess.opt, devices=devices, **kw)
        for i, name in enumerate(gh.FILE_NAMES):
            assert got[i] == want[i], "%s on devices %s: %s" % (name, devices, gh.first_diff(want[i], got[i]))
        assert sum(s.n_pairs for s in stats) == sess.stats.n_pairs_total
        assert sum(s.n_random for s in stats) == sess.stats.n_random
    finally:
        sess.close()


@pytest.mark.parametrize("case", ["illumina_indel_heavy", "solid_2x50", "illumina_maxn_hap"])
def test_group_equals_oracle(oracle, synth_fa, tmp_path, case):
    for devices in device_sets():
        check(oracle, make_golden.MATRIX[case], synth_fa, tmp_path, devices, batch=500)


def test_group_per_contig_runs_and_uneven_batches(oracle, synth_fa, tmp_path):
    """run() once per contig (the host shell's pattern), batch counts that do not divide by the number of ranks"""
    opts = dict(seed=3, N=5000, length=(100, 100), mut_rate=0.02, indel_frac=0.5, rand_read=0.2)
    for devices in device_sets():
        check(oracle, opts, synth_fa, tmp_path, devices, batch=613, per_contig_runs=True)


def test_group_device_gzip_is_one_code_for_all_ranks(oracle, synth_fa, tmp_path):
    """with the device gzip writer every rank uses the code fitted to batch 0, so the members are those of one device"""
    opts = dict(seed=9, N=6000, length=(100, 100))
    opts = make_golden.materialize(opts, str(tmp_path))
    sess, want = gh.oracle_expected(oracle, opts, synth_fa, str(tmp_path / "orc"))
    try:
        from dwgsim_b200 import DwgsimGpu, params_from_options
        outs = []
        for devices in (None, [0, 0, 0]):
            got = [[], [], []]
            with DwgsimGpu(params_from_options(**{k: v for k, v in opts.items() if k in gh.GPU_KEYS}), devices=devices) as gpu:
                gpu.set_batch(700, 2)
                gpu.set_compression(1)
                for k in range(sess.n_contigs):
                    c = sess.contig(k)
                    gpu.add_contig(c["contig_i"], c["name"], c["seq"], c["len"], c["hap"][0], c["hap"][1], c["ins"][0],
                                   c["n_ins"][0], c["ins"][1], c["n_ins"][1], c["n_pairs"])
                gpu.run(lambda fid, data: got[fid].append(data))
            outs.append([b"".join(g) for g in got])
        assert outs[0] == outs[1]
        import gzip
        for i in range(3):
            assert gzip.decompress(outs[1][i]) == want[i]
    finally:
        sess.close()


def test_cli_gpus_option(oracle, synth_fa, tmp_path):
    """the drop-in with --gpus 2 (ranks sharing cuda:0 through DWGSIM_DEVICES on a one-GPU box) writes the files of --gpus 1"""
    from dwgsim_b200 import build
    exe = build.build_cli()
    outs = []
    for tag, env_extra, extra in (("one", {}, []), ("two", {"DWGSIM_DEVICES": "0,0"}, ["--gpus", "2"])):
        prefix = str(tmp_path / tag)
        env = dict(os.environ, **env_extra)
        r = subprocess.run([exe, "-z", "5", "-N", "8000", "-1", "100", "-2", "100", "--batch", "1000", "--uncompressed"] + extra +
                           [synth_fa, prefix], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        outs.append([open(prefix + "." + f, "rb").read() for f in gh.FILE_NAMES + ["mutations.txt", "mutations.vcf"]])
    assert outs[0] == outs[1]
    assert len(outs[0][0]) > 0
