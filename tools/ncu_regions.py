#!/usr/bin/env python
"""tools/ncu_regions.py <file.ncu-rep> <kernel regex> a-b[:label] ...: executed instructions and stall samples of
kernels.cuh summed over source-line ranges (first launch of the kernel; capture needs --import-source on, -lineinfo)."""
import csv, io, subprocess, sys


def table(rep, rx):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + rx], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if not row or row[0] in ("Kernel Name", "Function Name"):
            continue
        if row[0] in ("File Name", "File Path"):
            cur = {"file": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
            continue
        if cur is None:
            continue
        if cur["hdr"] is None:
            cur["hdr"] = row
        elif row[0] != "":
            cur["rows"].append(row)
    return [b for b in blocks if b["file"].endswith("kernels.cuh")][0]


def main():
    blk = table(sys.argv[1], sys.argv[2])
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
    print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
    for spec in sys.argv[3:]:
        rng, _, label = spec.partition(":")
        a, b = (int(x) for x in rng.split("-"))
        rows = [r for r in blk["rows"] if a <= int(r[h["Line No"]]) <= b]
        i = sum(num(r, "Instructions Executed") for r in rows)
        s = sum(num(r, "# Samples") for r in rows)
        print("%-28s lines %5d-%-5d inst %6.2f%% (%11d)  samples %6.2f%%" % (label or rng, a, b, 100 * i / max(tot_i, 1), i, 100 * s / max(tot_s, 1)))


if __name__ == "__main__":
    main()
