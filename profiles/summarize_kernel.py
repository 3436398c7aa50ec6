#!/usr/bin/env python
"""Any kernel's `ncu --set full` capture -> a small tracked summary (markdown + json) under profiles/.

    python profiles/summarize_kernel.py gpurun_out/r02y/union_level.ncu-rep profiles/r02y_c2m_union_level_ncu "note"
"""
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
]


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
    launches = []
    for b in body:
        d = {"kernel": b[col["Kernel Name"]]}
        for k in KEEP + stalls:
            if k in col and b[col[k]] not in ("", "n/a"):
                d[k] = {"value": float(b[col[k]].replace(",", "")), "unit": units[col[k]]}
        launches.append(d)
    with open(out + ".json", "w") as f:
        json.dump({"report": rep, "note": note, "launches": launches}, f, indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full --clock-control none: {rep.split('/')[-1]}\n\n{note}\n\n")
        for i, d in enumerate(launches):
            f.write(f"## launch {i}: `{d['kernel']}`\n\n| metric | value |\n|---|---|\n")
            for k in KEEP:
                if k in d:
                    f.write(f"| {k} | {d[k]['value']:.6g} {d[k]['unit']} |\n")
            for k in stalls:
                if k in d and d[k]["value"] >= 0.3:
                    f.write(f"| {k.replace('smsp__average_warps_issue_stalled_', 'stall: ').replace('_per_issue_active.ratio', '')} | {d[k]['value']:.3g} per issue |\n")
            f.write("\n")
    print("written", out + ".md")


if __name__ == "__main__":
    main()
