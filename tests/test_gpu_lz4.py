"""Device-side LZ4-frame decoder (kmdiff_b200/csrc/kmd_lz4.cuh) against the host decoder and liblz4:
the reference's own `.kmer.lz4` files (config-1 fixture written by kmtricks, and the merge_test.cpp fixture), frames
written by liblz4 through pyarrow (linked 64 KiB blocks, km::lz4_stream's format), frames of this repo's own writer
(independent blocks, stored blocks), and adversarial contents (long runs, short-period matches, far matches, empty)."""
import os
import subprocess
import tarfile

import numpy as np
import pytest

import kmdiff_b200 as K
import oracle
from kmdiff_b200 import build as kb
from kmdiff_b200 import kmtricks_io as kio

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def liblz4_frame(payload: bytes) -> bytes:
    import pyarrow as pa
    return pa.compress(payload, codec="lz4", asbytes=True)


def host_decode(cli, frame: bytes, tmp_path) -> bytes:
    p = tmp_path / "frame.lz4"
    p.write_bytes(frame)
    return subprocess.run([cli, "dev-cat", str(p), "0", "lz4"], capture_output=True, timeout=120).stdout


@pytest.fixture(scope="module")
def cli():
    return kb.build_cli()


def ctx_for(nc, nk, k=31, b=4):
    return K.DiffContext(nc, nk, 10**9 * nc, 10**9 * nk, kmer_size=k, count_bytes=b)


def test_decodes_the_reference_fixture_files(cli, tmp_path):
    """counts/partition_p/*.kmer.lz4 of examples/run.sh, compressed by kmtricks itself"""
    with tarfile.open(os.path.join(HERE, "golden", "config1_run.tar.xz")) as t:
        t.extractall(tmp_path, filter="data")
    total = 0
    with ctx_for(10, 10) as ctx:
        for p in range(4):
            paths = kio.get_partition_paths(str(tmp_path), p)
            frames = [np.fromfile(f, dtype=np.uint8)[kio.KMER_HEADER_BYTES:] for f in paths]
            want = [kio.read_kmer_file(f)[1] for f in paths]       # liblz4 (pyarrow)
            got, ms = ctx.lz4_decode_host(frames, sum(w.size for w in want) // 12)
            assert got.tobytes() == b"".join(w.tobytes() for w in want)
            assert host_decode(cli, frames[3].tobytes(), tmp_path) == want[3].tobytes()
            total += got.size
            # ... and straight into the merge + scan: same rows as from the decoded records
            rows, recs = ctx.merge_scan_lz4_host(frames, ab_min=2, row_base=total, partition=p)
            assert recs == got.size // 12
            assert rows == len(oracle.kmer_merge(want, 1, 4, 2))
    assert total > 1_000_000


@pytest.mark.parametrize("name", ["records", "zeros", "period3", "far", "random", "empty", "tiny", "mixed"])
def test_decodes_liblz4_and_own_frames(cli, tmp_path, name):
    rng = np.random.default_rng(5)
    spec = K.SynthSpec(10, 10, total_rows=400_000, seed=9)
    recs = kio.rows_to_streams(spec.rows_host(400_000), 20, 1, 4)
    payloads = {
        "records": [r.tobytes() for r in recs],
        "zeros": [bytes(12 * (1000 + 37 * i)) for i in range(20)],
        "period3": [bytes([1, 2, 3]) * (4 * (5000 + i)) for i in range(20)],
        # matches far beyond the 2 KiB ring and across 64 KiB block borders
        "far": [(rng.integers(0, 256, 6000, dtype=np.uint8).tobytes() * 40)[:12 * 19000] for _ in range(20)],
        "random": [rng.integers(0, 256, 12 * 7001, dtype=np.uint8).tobytes() for _ in range(20)],
        "empty": [b""] * 20,
        "tiny": [bytes(range(12)) * (i % 3) for i in range(20)],
    }
    payloads["mixed"] = [payloads[k][i] for i, k in enumerate(["records", "zeros", "period3", "far", "random"] * 4)]
    want = payloads[name]
    frames = [np.frombuffer(liblz4_frame(w), dtype=np.uint8) for w in want]
    with ctx_for(10, 10) as ctx:
        got, ms = ctx.lz4_decode_host(frames, sum(len(w) for w in want) // 12 + 1)
        assert got.tobytes() == b"".join(want)
        # frames of this repo's writer: independent blocks, stored blocks for incompressible data
        own = []
        for w in want:
            src, dst = tmp_path / "in.bin", tmp_path / "out.lz4"
            src.write_bytes(w)
            assert subprocess.run([cli, "dev-pack", str(src), str(dst)], timeout=60).returncode == 0
            own.append(np.fromfile(dst, dtype=np.uint8))
        got, _ = ctx.lz4_decode_host(own, sum(len(w) for w in want) // 12 + 1)
        assert got.tobytes() == b"".join(want)


def test_malformed_frames_are_reported():
    good = np.frombuffer(liblz4_frame(bytes(1200)), dtype=np.uint8)
    with ctx_for(1, 1) as ctx:
        for bad in (good[:len(good) // 2], np.concatenate([np.array([1, 2, 3, 4], np.uint8), good[4:]])):
            with pytest.raises(K.KmdError):
                ctx.lz4_decode_host([good, bad], 1000)
        # a payload that is not a whole number of records
        with pytest.raises(K.KmdError):
            ctx.lz4_decode_host([good, np.frombuffer(liblz4_frame(bytes(13)), dtype=np.uint8)], 1000)
        got, _ = ctx.lz4_decode_host([good, good], 1000)
        assert got.size == 2400


def test_partition_too_large_for_one_call_is_refused():
    spec = K.SynthSpec(10, 10, total_rows=50_000, seed=3)
    recs = kio.rows_to_streams(spec.rows_host(50_000), 20, 1, 4)
    frames = [np.frombuffer(liblz4_frame(r.tobytes()), dtype=np.uint8) for r in recs]
    with ctx_for(10, 10) as ctx:
        with pytest.raises(K.KmdError) as e:
            ctx.merge_scan_lz4_host(frames, max_records=1000)
        assert e.value.code == K._lib.ERR_OVERFLOW if hasattr(K, "_lib") else True
        rows, n = ctx.merge_scan_lz4_host(frames)
        assert rows == 50_000 and n == sum(r.size for r in recs) // 12
        assert ctx.counters().rows == 50_000
