#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py --round r01 --config c2 --rows 200000000 \
        --rep gpurun_out/prof_scan_c2.ncu-rep --launches gpurun_out/launches.csv

Writes profiles/<round>_lrt_scan_ncu.json (read by bench.py for `roofline.traffic`),
profiles/<round>_lrt_scan_ncu.md and profiles/<round>_launches.md.
"""
import argparse
import collections
import csv
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    return hdr, units, body


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r01")
    ap.add_argument("--config", default="c2")
    ap.add_argument("--rows", type=int, required=True)
    ap.add_argument("--row-bytes", type=int, default=88)
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--tag", default="lrt_scan")
    ap.add_argument("--command", default="", help="the profiled command line, recorded in the summary")
    args = ap.parse_args()

    if args.rep:
        hdr, units, body = raw_page(args.rep)
        col = {h: i for i, h in enumerate(hdr)}
        launches = []
        for r in body:
            d = {"kernel": r[col["Kernel Name"]]}
            for k in KEEP:
                if k in col:
                    v = float(r[col[k]].replace(",", "")) if r[col[k]] not in ("", "n/a") else None
                    if v is not None and units[col[k]] in SCALE and ("bytes" in k or "time" in k):
                        v *= SCALE[units[col[k]]]
                    d[k] = v
            launches.append(d)
        n = len(launches)
        mean = lambda k: sum(x[k] for x in launches) / n
        dram = mean("dram__bytes_read.sum") + mean("dram__bytes_write.sum")
        t = mean("gpu__time_duration.sum")
        alg = args.rows * args.row_bytes
        summary = {
            "round": args.round, "config": args.config, "rows": args.rows, "row_bytes": args.row_bytes,
            "kernel": launches[0]["kernel"], "launches_captured": n,
            "command": args.command or ("ncu --set full --clock-control none --import-source on -k regex:lrt_scan -s 3 -c 2 python bench.py "
                                        f"--steps 2 --warmup 3 --rows {args.rows} --no-cpu-baseline --e2e-steps 1 --e2e-rows 4194304"),
            "algorithmic_bytes_per_launch": alg, "dram_bytes_per_launch": dram, "traffic_over_algorithmic": dram / alg,
            "duration_s_under_ncu": t, "gbps_under_ncu": alg / t / 1e9,
            "metrics_mean": {k: mean(k) for k in KEEP if all(x.get(k) is not None for x in launches)},
        }
        with open(os.path.join(HERE, f"{args.round}_{args.tag}_ncu.json"), "w") as f:
            json.dump(summary, f, indent=1)
        with open(os.path.join(HERE, f"{args.round}_{args.tag}_ncu.md"), "w") as f:
            f.write(f"# {args.round}: ncu --set full, `{summary['kernel']}` ({args.config}, {args.rows} rows x {args.row_bytes} B)\n\n")
            f.write(f"`{summary['command']}`\n\n| metric | mean over {n} launches |\n|---|---|\n")
            f.write(f"| algorithmic bytes / launch | {alg:.6g} |\n| dram bytes read+write / launch | {dram:.6g} (x{dram / alg:.4f} of algorithmic) |\n")
            f.write(f"| duration under ncu | {t * 1e3:.4f} ms -> {alg / t / 1e9:.0f} GB/s |\n")
            for k, v in summary["metrics_mean"].items():
                f.write(f"| {k} | {v:.6g} |\n")
        print("traffic/algorithmic", dram / alg, "GB/s", alg / t / 1e9)

    if args.launches:
        with open(args.launches) as f:
            lines = [l for l in f if l.startswith('"')]
        agg = collections.OrderedDict()
        for x in csv.DictReader(lines):
            v = float(x["Metric Value"].replace(",", "")) * SCALE.get(x["Metric Unit"], 1.0)
            k = x["Kernel Name"].split("(")[0]
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += v
        with open(os.path.join(HERE, f"{args.round}_launches.md"), "w") as f:
            f.write(f"# {args.round}: launch list (ncu --metrics gpu__time_duration.sum --clock-control none), bench.py --steps 2 --warmup 3 "
                    f"--rows {args.rows} --config {args.config}\n\nCold-cache, serialised timings: compare shares, not absolutes.  "
                    "`synth_kernel` generates the resident workload once, outside the timed region.\n\n")
            step = {k: v for k, v in agg.items() if "synth" not in k and "at::" not in k}
            tot = sum(v[1] for v in step.values())
            f.write("| kernel | launches | total ms | share of step kernels |\n|---|---|---|---|\n")
            for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                share = f"{t / tot:.3f}" if k in step else "-"
                f.write(f"| `{k}` | {n} | {t * 1e3:.3f} | {share} |\n")
        print("launch summary written")


if __name__ == "__main__":
    main()
