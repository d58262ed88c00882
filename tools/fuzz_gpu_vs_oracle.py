"""Parity fuzz on the GPU (run under gpurun): random option sets through the oracle's Philox backend and through the C ABI
(tests/gpu_harness.py plays the reference host); the three FASTQ streams must be byte-identical.
    python tools/fuzz_gpu_vs_oracle.py SEED N [random-fasta]"""
import os, random, shutil, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as mg
import gpu_harness as gh
from oracle import pyoracle as po

WD = "/tmp/dwgsim_fuzz_gpu"; os.makedirs(WD, exist_ok=True)
fa = mg.synth_fasta(os.path.join(WD, "synth.fa"))
RANDOM_FASTA = len(sys.argv) > 3 and sys.argv[3] == "random-fasta"
if RANDOM_FASTA:                    # a new FASTA every 10 cases (tools/fuzz_oracle_vs_reference.py: random_fasta)
    src = open(os.path.join(ROOT, "tools", "fuzz_oracle_vs_reference.py")).read()
    ns = {}
    exec(src[src.index("def random_fasta"):src.index("fa = mg.synth_fasta")], ns)
rnd = random.Random(int(sys.argv[1])); bad = 0; done = 0; t0 = time.time()
for it in range(int(sys.argv[2])):
    if RANDOM_FASTA and it % 10 == 0:
        fa = ns["random_fasta"](os.path.join(WD, "rand.fa"), int(sys.argv[1]) * 100000 + it)
    o = dict(seed=rnd.randint(0, 10 ** 6))
    dt = rnd.choice([0, 0, 0, 1, 1, 2])
    o["data_type"] = dt
    l0 = rnd.choice([25, 36, 50, 70, 75, 100, 101, 125, 150, 151, 250, 300]); l1 = rnd.choice([0, 25, 50, 100, 150, l0, l0])
    o["length"] = (l0, l1)
    if rnd.random() < 0.6: o["N"] = rnd.randint(200, 6000)
    else: o["C"] = rnd.choice([0.5, 2, 5])
    if dt == 2:
        o["flow_order"] = rnd.choice([mg.FLOW, "TACG", "TCAGTCAG"]); e = rnd.choice([0.005, 0.02, 0.1]); o["e"] = e; o["E"] = rnd.choice([e, 0.01])
    else:
        if rnd.random() < 0.7: o["e"] = rnd.choice(["0.0", "0.02", "0.001-0.05", "0.1-0.0", "0.3", "0.001-0.01"])
        if rnd.random() < 0.7: o["E"] = rnd.choice(["0.0", "0.02", "0.001-0.05", "0.2", "0.001-0.01"])
    if rnd.random() < 0.6: o["mut_rate"] = rnd.choice([0, 0.001, 0.01, 0.05])
    if rnd.random() < 0.5: o["indel_frac"] = rnd.choice([0, 0.1, 0.5, 1])
    if rnd.random() < 0.4: o["indel_extend"] = rnd.choice([0, 0.3, 0.9, 0.98])
    if rnd.random() < 0.3: o["indel_min"] = rnd.choice([1, 2, 4])
    if rnd.random() < 0.2: o["is_hap"] = 1
    if rnd.random() < 0.3: o["mut_freq"] = rnd.choice([0, 0.2, 1])
    if rnd.random() < 0.4: o["rand_read"] = rnd.choice([0, 0.05, 0.5, 0.95])
    if rnd.random() < 0.3: o["max_n"] = rnd.choice([0, 2, 20])
    if rnd.random() < 0.4: o["dist"] = rnd.choice([150, 300, 500, 2000]); o["std_dev"] = rnd.choice([0, 10, 50, 200])
    if rnd.random() < 0.3: o["is_inner"] = 1
    if rnd.random() < 0.3: o["strandedness"] = rnd.choice([0, 1, 2])
    if rnd.random() < 0.3: o["read_one_strand"] = rnd.choice([0, 1, 2])
    if rnd.random() < 0.3: o["quality_std"] = rnd.choice([0, 1, 2, 7, 40])
    if rnd.random() < 0.15: o["fixed_quality"] = rnd.choice(["I", "5"])
    if rnd.random() < 0.15: o["read_prefix"] = "pf"
    if rnd.random() < 0.3: o["reads_output_type"] = rnd.choice([0, 1, 2])
    if rnd.random() < 0.1 and l1 > 0: o["amplicons"] = 1; o["max_n"] = 300
    kw = {}
    if rnd.random() < 0.3: kw["batch"] = rnd.choice([100, 257, 1000, 4096])
    if rnd.random() < 0.15: kw["compression"] = 1
    if rnd.random() < 0.15: kw["devices"] = rnd.choice([[0, 0], [0, 0, 0]])       # a device group whose ranks share cuda:0
    if rnd.random() < 0.15: kw["per_contig_runs"] = True
    sub = os.path.join(WD, "case"); shutil.rmtree(sub, ignore_errors=True); os.makedirs(sub)
    try:
        sess, want = gh.oracle_expected(po, o, fa, os.path.join(sub, "orc"))
    except ValueError:
        continue
    try:
        if sess.stats.error != 0 or sess.n_contigs == 0:        # nothing to hand to the C ABI (every contig skipped)
            continue
        try:
            got, stats = gh.gpu_actual(sess, o, orc_opt=sess.opt, **kw)
        except Exception as ex:
            bad += 1; print("GPU ERROR", it, o, kw, repr(ex)[:200]); continue
        done += 1
        for i, name in enumerate(gh.FILE_NAMES):
            if got[i] != want[i]:
                bad += 1; print("MISMATCH", it, name, o, kw, gh.first_diff(want[i], got[i])[:300]); break
    finally:
        sess.close()
    if bad > 4: break
print("done: %d cases compared, mismatches %d, %.0f s" % (done, bad, time.time() - t0))
