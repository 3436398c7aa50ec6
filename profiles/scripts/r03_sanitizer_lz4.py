"""small device-LZ4 decode for compute-sanitizer (racecheck is ~100x slower than a plain run)"""
import os, sys
import numpy as np
import pyarrow as pa
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import kmdiff_b200 as K

rng = np.random.default_rng(2)
streams = []
for s in range(4):
    n = 6000 + 1000 * s
    keys = np.sort(rng.integers(0, 2**62, size=n, dtype=np.uint64))
    rec = np.zeros((n, 12), np.uint8)
    rec[:, :8] = keys.view(np.uint8).reshape(-1, 8)
    rec[:, 8:] = (rng.poisson(6, size=n).astype(np.uint32) + 1).view(np.uint8).reshape(-1, 4)
    if s == 3:
        rec[2000:2400] = rec[2000]          # a long run: length extensions and short-period matches
    streams.append(rec.ravel())
frames = [np.frombuffer(pa.compress(x.tobytes(), codec="lz4", asbytes=True), dtype=np.uint8) for x in streams]
with K.DiffContext(2, 2, 10**9, 10**9, kmer_size=31, count_bytes=4, threshold=1e-9) as ctx:
    got, ms = ctx.lz4_decode_host(frames, sum(x.size for x in streams) // 12)
assert np.array_equal(got, np.concatenate(streams))
print("ok", got.size, "bytes", ms, "ms")
