import sys, numpy as np
import os; R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import kmdiff_b200 as K, oracle
from test_gpu_merge import _random_streams, ctx_for
rng = np.random.default_rng(3)
for (n, k, b, m, dens) in [(20, 31, 4, 40000, 0.9), (7, 63, 2, 20000, 0.9), (33, 20, 1, 15000, 0.3), (200, 31, 4, 4000, 0.05)]:
    streams, kw = _random_streams(rng, n, k, b, m, lambda s: dens)
    want = oracle.kmer_merge(streams, kw, b)
    with ctx_for(n // 2, n - n // 2, k, b) as ctx:
        got = ctx.merge_rows_host(streams)
        info = ctx.merge_info()
    assert np.array_equal(got, want), (n, k, b)
    print("ok", n, k, b, info["algo"], info["rows"], len(want), flush=True)
