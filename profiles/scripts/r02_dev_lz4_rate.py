#!/usr/bin/env python
"""Decode rate of the device LZ4-frame decoder (kmd_lz4_decode_host: one warp per stream, size pass + copy pass) against
the number of streams in flight, on c2-like count-file streams (sorted 12-byte records, LZ4 frames of linked 64 KiB
blocks as kmtricks writes them).  Writes one JSON line."""
import json
import os
import sys
import time

import numpy as np
import pyarrow as pa

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kmdiff_b200 as K
from kmdiff_b200 import kmtricks_io as kio


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    spec = K.SynthSpec(10, 10, kmer_size=31, count_bytes=4, total_rows=rows, seed=42, enriched_ppm=1000)
    dense = spec.rows_host(rows)
    base = kio.rows_to_streams(dense, 20, 1, 4)
    frames = [np.frombuffer(pa.compress(s.tobytes(), codec="lz4", asbytes=True), dtype=np.uint8) for s in base]
    out = {"rows": rows, "record_bytes": 12, "decoded_bytes_per_20_streams": int(sum(s.size for s in base)),
           "compressed_bytes_per_20_streams": int(sum(f.size for f in frames)), "runs": []}
    for n_streams in (20, 200, 640, 2000, 6000):
        reps = (n_streams + 19) // 20
        fr = (frames * reps)[:n_streams]
        dec = sum((base * reps)[i].size for i in range(n_streams))
        with K.DiffContext(n_streams // 2, n_streams - n_streams // 2, 10**9, 10**9, kmer_size=31, count_bytes=4, threshold=1e-9) as ctx:
            got, ms = ctx.lz4_decode_host(fr, dec // 12)      # warm-up (buffers)
            t0 = time.perf_counter()
            got, ms = ctx.lz4_decode_host(fr, dec // 12)
            wall = time.perf_counter() - t0
        ok = bool(np.array_equal(got[:base[0].size], base[0])) and got.size == dec
        out["runs"].append({"streams": n_streams, "decoded_bytes": int(dec), "device_ms_two_passes": ms,
                            "decoded_gbs": dec / (ms * 1e-3) / 1e9, "per_stream_mbs": dec / n_streams / (ms * 1e-3) / 1e6,
                            "wall_s_incl_copies": wall, "bytes_equal": ok})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
