#!/usr/bin/env python
"""tools/ncu_read.py <file.ncu-rep>: key metrics per kernel from an ncu capture (runs without a GPU)."""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("----", r[col["Kernel Name"]][:60])
        for k in KEYS:
            if k in col:
                print("   %s = %s %s" % (k, r[col[k]], units[col[k]]))
        st = []
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or \
               h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio"):
                try:
                    st.append((float(r[i].replace(",", "")), h.split("stalled_")[1].split("_per_issue")[0].replace(".ratio", "")))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("   stalls: " + ", ".join("%s=%.2f" % (n, v) for v, n in st[:8]))


if __name__ == "__main__":
    main()
