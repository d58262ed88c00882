"""Differential fuzz (build container only: needs oracle/_ref/dwgsim_ref): random option sets through the compiled reference and the
oracle's drand48 backend; all five output files must be byte-identical.
    python tools/fuzz_oracle_vs_reference.py SEED N [random-fasta]"""
import gzip, hashlib, os, random, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as mg
from oracle import pyoracle as po

WD = "/tmp/dwgsim_fuzz"; os.makedirs(WD, exist_ok=True)
def random_fasta(path, seed):
    """contigs of assorted lengths around the skip thresholds, N runs, lowercase, IUPAC codes, Windows line ends now and then"""
    import numpy as np
    rng = np.random.default_rng(seed)
    with open(path, "wb") as f:
        for k in range(int(rng.integers(1, 7))):
            n = int(rng.choice([60, 200, 500, 700, 1500, 5000, 20000]))
            s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
            for _ in range(int(rng.integers(0, 3))):
                a = int(rng.integers(0, n)); s[a:a + int(rng.integers(1, max(2, n // 4)))] = ord("N")
            if rng.random() < 0.3:
                a = int(rng.integers(0, n)); s[a:a + 50] += 32
            if rng.random() < 0.3:
                s[int(rng.integers(0, n))] = ord(rng.choice(list("RYKMSWBDHV")))
            f.write((">c%d %s\n" % (k, "desc" if rng.random() < 0.5 else "")).encode())
            w = int(rng.choice([50, 60, 70, 1000000]))
            eol = b"\r\n" if rng.random() < 0.15 else b"\n"
            b = s.tobytes()
            for i in range(0, len(b), w):
                f.write(b[i:i + w] + eol)
    return path


fa = mg.synth_fasta(os.path.join(WD, "synth.fa"))
if len(sys.argv) > 3 and sys.argv[3] == "random-fasta":
    fa = random_fasta(os.path.join(WD, "rand.fa"), int(sys.argv[1]))
FLOW = mg.FLOW


def md5(p):
    if not os.path.exists(p):
        return None
    op = gzip.open if p.endswith(".gz") else open
    with op(p, "rb") as f:
        return hashlib.md5(f.read()).hexdigest()


rnd = random.Random(int(sys.argv[1])); bad = 0; compared = 0
for it in range(int(sys.argv[2])):
    o = dict(seed=rnd.randint(0, 10 ** 6))
    dt = rnd.choice([0, 0, 0, 1, 2])
    o["data_type"] = dt
    l0 = rnd.choice([30, 50, 70, 100, 150]); l1 = rnd.choice([0, 30, 50, 100, l0])
    o["length"] = (l0, l1)
    if rnd.random() < 0.5: o["N"] = rnd.randint(50, 1500)
    else: o["C"] = rnd.choice([0.05, 0.2, 1])
    if dt == 2:
        o["flow_order"] = rnd.choice([FLOW, "TACG", "TCAGTCAG"]); e = rnd.choice([0.005, 0.02, 0.1]); o["e"] = e; o["E"] = rnd.choice([e, 0.01])
        if rnd.random() < 0.2: o["use_base_error"] = 1
    else:
        if rnd.random() < 0.6: o["e"] = rnd.choice(["0.0", "0.02", "0.001-0.05", "0.1-0.0", "0.3"])
        if rnd.random() < 0.6: o["E"] = rnd.choice(["0.0", "0.02", "0.001-0.05", "0.2"])
    if rnd.random() < 0.5: o["mut_rate"] = rnd.choice([0, 0.001, 0.01, 0.05])
    if rnd.random() < 0.5: o["indel_frac"] = rnd.choice([0, 0.1, 0.5, 1])
    if rnd.random() < 0.4: o["indel_extend"] = rnd.choice([0, 0.3, 0.9, 0.98])
    if rnd.random() < 0.3: o["indel_min"] = rnd.choice([1, 2, 4])
    if rnd.random() < 0.3: o["is_hap"] = 1
    if rnd.random() < 0.3: o["mut_freq"] = rnd.choice([0, 0.2, 1])
    if rnd.random() < 0.4: o["rand_read"] = rnd.choice([0, 0.05, 0.5])
    if rnd.random() < 0.3: o["max_n"] = rnd.choice([0, 2, 20])
    if rnd.random() < 0.4: o["dist"] = rnd.choice([150, 300, 500, 2000]); o["std_dev"] = rnd.choice([0, 10, 50, 200])
    if rnd.random() < 0.3: o["is_inner"] = 1
    if rnd.random() < 0.3: o["strandedness"] = rnd.choice([0, 1, 2])
    if rnd.random() < 0.3: o["read_one_strand"] = rnd.choice([0, 1, 2])
    if rnd.random() < 0.3: o["quality_std"] = rnd.choice([0, 1, 2, 7, 40])
    if rnd.random() < 0.15: o["fixed_quality"] = rnd.choice(["I", "5"])
    if rnd.random() < 0.15: o["read_prefix"] = "pf"
    if rnd.random() < 0.3: o["reads_output_type"] = rnd.choice([0, 1, 2])
    if rnd.random() < 0.1 and l1 > 0: o["amplicons"] = 1; o["max_n"] = 100
    sub = os.path.join(WD, "case"); shutil.rmtree(sub, ignore_errors=True); os.makedirs(sub)
    try:
        opt = po.make_opt(**o)
    except ValueError:
        r = subprocess.run([po.ref_binary()] + po.opt_to_ref_argv(**o) + [fa, os.path.join(sub, "ref")], capture_output=True)
        if r.returncode == 0:
            bad += 1; print("ORACLE REJECTS, REFERENCE ACCEPTS", o)
        continue
    r = subprocess.run([po.ref_binary()] + po.opt_to_ref_argv(**o) + [fa, os.path.join(sub, "ref")], capture_output=True, timeout=120)
    with po.Session(opt, fa, os.path.join(sub, "orc")) as s:
        err = s.stats.error
    if r.returncode != 0 or err != 0:
        if (r.returncode != 0) != (err != 0):
            bad += 1; print("EXIT MISMATCH", it, o, r.returncode, err, r.stderr.decode(errors="ignore")[-150:])
        continue
    compared += 1
    for f in mg.FILES:
        a = md5(os.path.join(sub, "ref." + f + (".gz" if f.endswith("fastq") else ""))); b = md5(os.path.join(sub, "orc." + f))
        if a != b:
            bad += 1; print("MISMATCH", it, f, o); break
    if bad > 4: break
print("done: %d cases compared file by file, mismatches %d" % (compared, bad))
