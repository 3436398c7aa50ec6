import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")
    # native pieces: built here when stale (no-op on the GPU box, where the prebuilt .so travel)
    from kmdiff_b200 import build as kb
    kb.build_lib()
    import oracle
    if not os.path.exists(oracle.ORACLE_SO):
        oracle.build(ref=False)


def _has_gpu() -> bool:
    from kmdiff_b200 import _lib
    return _lib.load().kmd_device_count() > 0


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# Cohort shapes exercised by the parity tests: (nb_controls, nb_cases, k, count_bytes)
SHAPES = {
    "c2": (10, 10, 31, 4),      # BASELINE config 2 (R = 88)
    "c3": (100, 100, 31, 4),    # config 3 (R = 808)
    "c5": (50, 50, 63, 2),      # config 5 (R = 216, 128-bit k-mers, 2-byte counts)
    "c4": (200, 200, 31, 4),    # config 4 row shape (R = 1608)
    "b1": (12, 20, 21, 1),      # 1-byte counts, unequal cohorts
    "odd": (7, 6, 31, 4),       # N*B % 8 != 0: generic (non word-aligned) path, R = 60
    "straddle": (5, 7, 31, 4),  # cohort boundary inside an 8-byte word
    "tiny": (1, 1, 20, 4),      # the reference's own unit fixture shape (1 control + 1 case)
    "r128": (15, 15, 31, 4),    # R = 128: worst case for shared-memory bank conflicts
    "b2odd": (3, 5, 40, 2),     # 2-byte counts, k-mer in two words, straddling boundary
}


@pytest.fixture(scope="session")
def shapes():
    return SHAPES


def make_cohort(shape: str, n_rows: int, seed: int = 1, enriched_ppm: int = 2000):
    """(spec, payload bytes, totals) of a seeded synthetic cohort."""
    import kmdiff_b200 as K
    nc, nk, k, b = SHAPES[shape]
    spec = K.SynthSpec(nc, nk, kmer_size=k, count_bytes=b, total_rows=max(n_rows, 1), seed=seed,
                       enriched_ppm=enriched_ppm)
    rows = spec.rows_host(n_rows) if n_rows else np.empty(0, np.uint8)
    return spec, rows, spec.totals()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.where(b == 0, 1.0, np.abs(b))
    return np.abs(a - b) / denom


def materialize_km_out_dir(dst: str) -> str:
    """The reference's merge_test.cpp run directory (tests/golden/km_out_dir.json: its files as hex) written under dst."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "km_out_dir.json")) as f:
        files = json.load(f)["files"]
    for rel, hx in files.items():
        path = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            f.write(bytes.fromhex(hx))
    os.makedirs(os.path.join(dst, "matrices"), exist_ok=True)      # kmtricks leaves an empty matrices/ directory
    return dst


@pytest.fixture
def km_out_dir(tmp_path):
    return materialize_km_out_dir(str(tmp_path / "km_out_dir"))
