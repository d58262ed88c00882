#!/usr/bin/env python
"""tools/ncu_read.py <file.ncu-rep> [--traffic out.json]: key metrics per kernel from an ncu capture (runs without a GPU).
--traffic: the capture holds the launches of ONE device batch (tools/ncu_capture.sh with SKIP/COUNT on a batch boundary);
their DRAM bytes are summed into out.json, which bench.py reports as roofline.traffic."""
import csv, io, json, subprocess, sys

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
    if "--traffic" in sys.argv:
        out_json = sys.argv[sys.argv.index("--traffic") + 1]
        num = lambda r, k: float(r[col[k]].replace(",", ""))
        scale = lambda k: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[col[k]]]
        per = []
        for r in rows[2:]:
            rd, wr = num(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum"), num(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum")
            per.append({"kernel": r[col["Kernel Name"]].split("(")[0], "us": num(r, "gpu__time_duration.sum"), "dram_read": rd, "dram_write": wr})
        rec = {"dram_bytes_per_batch": sum(k["dram_read"] + k["dram_write"] for k in per), "pairs_per_batch": 1 << 20, "kernels": per,
               "source": "ncu --set full capture of the launches of one 2^20-pair device batch of bench.py (%s), dram__bytes_read.sum + dram__bytes_write.sum" % rep}
        json.dump(rec, open(out_json, "w"), indent=1)
        print("wrote", out_json, rec["dram_bytes_per_batch"])
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
