#!/usr/bin/env python
"""GPU-side sweep of the lrt_scan geometry knobs (tile bytes, stages, consumer warps, lanes per row, ring vs
direct loads) on device-resident synthetic rows.  Prints one line per setting: GB/s of the scan kernel
(CUDA events around the launch, best of 5 after 2 warm-ups).  Usage: python profiles/sweep_scan.py [c2 c3 c5]"""
import itertools
import time
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import kmdiff_b200 as K  # noqa: E402

SHAPES = {"c2": (10, 10, 31, 4, 100_000_000), "c3": (100, 100, 31, 4, 10_000_000), "c5": (50, 50, 63, 2, 40_000_000),
          "c4": (200, 200, 31, 4, 5_000_000), "b1": (12, 20, 21, 1, 100_000_000)}


def measure(spec, totals, data, rows, env):
    for k in ("KMD_SCAN_TILE_BYTES", "KMD_SCAN_STAGES", "KMD_SCAN_WARPS", "KMD_SCAN_G", "KMD_SCAN_MODE", "KMD_SCAN_GROUPS"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    nc, nk = spec.nb_controls, spec.nb_cases
    try:
        ctx = K.DiffContext(nc, nk, int(totals[:nc].sum()), int(totals[nc:].sum()), kmer_size=spec.kmer_size,
                            count_bytes=spec.count_bytes, threshold=0.05 / 100000, max_survivors=1 << 21)
    except K.KmdError as e:
        return None, str(e)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    best = 1e9
    n_surv = 0
    for i in range(7):
        ctx.reset()
        ctx.scan_device(data.data_ptr(), rows)
        ms = ctx.scan_kernel_ms()
        if i >= 2:
            best = min(best, ms)
    n_surv = ctx.counters().survivors
    ctx.close()
    return rows * spec.row_bytes / (best * 1e-3) / 1e9, n_surv


T0 = time.time()


def main():
    names = sys.argv[1:] or ["c2", "c3", "c5"]
    out = []
    for name in names:
        nc, nk, k, b, rows = SHAPES[name]
        spec = K.SynthSpec(nc, nk, kmer_size=k, count_bytes=b, total_rows=rows, seed=42)
        totals = spec.totals()
        data = torch.empty(rows * spec.row_bytes, dtype=torch.uint8, device="cuda")
        spec.rows_device(data.data_ptr(), rows)
        torch.cuda.synchronize()
        base, ns = measure(spec, totals, data, rows, {})
        print(f"{name} R={spec.row_bytes} default: {base:.0f} GB/s survivors={ns}", flush=True)
        out.append({"config": name, "env": {}, "gbps": base})
        settings = []
        if os.environ.get("SWEEP_FULL"):
            settings.append({"KMD_SCAN_MODE": "direct"})
            for tile, stages in itertools.product((8192, 16384, 24576, 32768, 49152, 65536), (2, 3, 4, 6, 8)):
                settings.append({"KMD_SCAN_TILE_BYTES": tile, "KMD_SCAN_STAGES": stages})
            for g in (1, 2, 4, 8, 16, 32):
                settings.append({"KMD_SCAN_G": g})
        for (groups, warps), tile in itertools.product(((1, 24), (2, 12), (2, 15), (3, 8), (3, 10), (4, 7), (5, 6)),
                                                       (16384, 32768)):
            settings.append({"KMD_SCAN_GROUPS": groups, "KMD_SCAN_WARPS": warps, "KMD_SCAN_TILE_BYTES": tile})
        if os.environ.get("SWEEP_COMBO"):
            settings = []
            for (groups, warps), tile, g in itertools.product(((3, 8), (3, 10), (5, 6), (6, 5), (2, 15)), (16384, 24576, 32768), (0, 4, 8, 16, 32)):
                env = {"KMD_SCAN_GROUPS": groups, "KMD_SCAN_WARPS": warps, "KMD_SCAN_TILE_BYTES": tile}
                if g:
                    env["KMD_SCAN_G"] = g
                settings.append(env)
        for env in settings:
            gbps, ns2 = measure(spec, totals, data, rows, env)
            if gbps is None:
                print(f"{name} {env}: rejected ({ns2})", flush=True)
                continue
            assert ns2 == ns, (ns2, ns)
            print(f"{name} {env}: {gbps:.0f} GB/s  [t={time.time() - T0:.1f}s]", flush=True)
            out.append({"config": name, "env": env, "gbps": gbps})
        del data
        torch.cuda.empty_cache()
    with open(os.path.join(ROOT, "gpurun_out", os.environ.get("SWEEP_OUT", "sweep_scan.json")), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
