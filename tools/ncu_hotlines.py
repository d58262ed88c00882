#!/usr/bin/env python
"""tools/ncu_hotlines.py <file.ncu-rep> <kernel regex> [top]: CUDA source lines of kernels.cuh by executed instructions
and stall samples (needs a capture made with --import-source on and a -lineinfo build)."""
import csv, io, subprocess, sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + rx], capture_output=True, text=True).stdout
    # the output holds one table per (kernel launch, file); keep the first launch's kernels.cuh table
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            continue
        if row[0] == "Function Name":
            continue
        if row[0] in ("File Name", "File Path"):
            cur = {"file": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
            continue
        if cur is None:
            continue
        if cur["hdr"] is None:
            cur["hdr"] = row
        elif row[0] != "":                # source-line rows; SASS rows have an empty line number
            cur["rows"].append(row)
    blk = [b for b in blocks if b["file"].endswith("kernels.cuh")][0]
    h = {}
    for i, n in enumerate(blk["hdr"]):
        h.setdefault(n, i)
    def num(r, k):
        try:
            return float(r[h[k]].replace(",", ""))
        except (ValueError, KeyError, IndexError):
            return 0.0
    tot_i = sum(num(r, "Instructions Executed") for r in blk["rows"])
    tot_s = sum(num(r, "# Samples") for r in blk["rows"])
    stall_cols = [n for n in blk["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    rows = sorted(blk["rows"], key=lambda r: -(num(r, "Instructions Executed") if "--by-inst" in sys.argv else num(r, "# Samples")))
    print("instructions %d samples %d" % (tot_i, tot_s))
    for r in rows[:top]:
        st = sorted(((num(r, c), c[6:]) for c in stall_cols), reverse=True)[:3]
        print("%5s inst=%5.1f%% smp=%5.1f%% thr=%4.1f %-34s| %s" % (
            r[h["Line No"]], 100 * num(r, "Instructions Executed") / max(tot_i, 1), 100 * num(r, "# Samples") / max(tot_s, 1),
            num(r, "Avg. Threads Executed"), ",".join("%s:%d" % (n, v) for v, n in st if v > 0), r[1].strip()[:110]))


if __name__ == "__main__":
    main()
