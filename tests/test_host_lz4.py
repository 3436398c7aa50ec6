"""The host program's LZ4 frame reader / writer (kmdiff_b200/host/lz4_frame.hpp) against liblz4 (through pyarrow), on the
CPU: kmtricks writes every matrix, count file and spill file as one LZ4 frame of linked 64 KiB blocks
(kmtricks/io/lz4_stream.hpp:42, 128-140), and the reader's fast paths (wild copies, the bounds-check-free zone for short
sequences) must reproduce liblz4's output byte for byte whatever the sequence shapes and the sizes it is read in."""
import os
import subprocess

import numpy as np
import pyarrow as pa
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("lz4") / "lz4_host_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "kmdiff_b200", "host"), "-o", exe,
                           os.path.join(ROOT, "tests", "tools", "lz4_host_check.cpp")])
    return exe


def _cases():
    rng = np.random.default_rng(7)
    cases = {"empty": b"", "tiny": b"abc", "zeros": bytes(500_000),
             "random": rng.integers(0, 256, size=300_000, dtype=np.uint8).tobytes()}
    for per in (1, 2, 3, 5, 7, 8, 9, 12, 15, 16, 17, 31, 33):          # overlapping matches of every short period
        cases[f"period{per}"] = (bytes(rng.integers(0, 256, size=per, dtype=np.uint8)) * (400_000 // per + 1))[:400_000]
    text = b"the quick brown fox jumps over the lazy dog " * 20000
    cases["text"] = text
    mix = bytearray()
    for _ in range(3000):                                              # literal runs, byte runs, far and near copies
        k = rng.integers(0, 4)
        if k == 0:
            mix += rng.integers(0, 256, size=rng.integers(1, 300), dtype=np.uint8).tobytes()
        elif k == 1:
            mix += bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 600))
        elif k == 2:
            mix += mix[max(0, len(mix) - int(rng.integers(1, 70000))):][:int(rng.integers(1, 400))]
        else:
            mix += text[:int(rng.integers(1, 200))]
    cases["mix"] = bytes(mix)
    n = 150_000                                                        # what a .kmer.lz4 file holds: sorted 12-byte records
    rec = np.zeros((n, 12), np.uint8)
    rec[:, :8] = np.sort(rng.integers(0, 2**62, size=n, dtype=np.uint64)).view(np.uint8).reshape(-1, 8)
    rec[:, 8:] = (rng.poisson(6, size=n).astype(np.uint32) + 1).view(np.uint8).reshape(-1, 4)
    cases["records12"] = rec.tobytes()
    rows = np.zeros((20000, 88), np.uint8)                             # ... and a matrix: 88-byte rows
    rows[:, :8] = np.sort(rng.integers(0, 2**62, size=20000, dtype=np.uint64)).view(np.uint8).reshape(-1, 8)
    rows[:, 8:] = rng.poisson(3, size=(20000, 20)).astype(np.uint32).view(np.uint8).reshape(20000, 80)
    cases["rows88"] = rows.tobytes()
    return cases


def test_reader_equals_liblz4_on_every_sequence_shape(harness, tmp_path):
    for name, data in _cases().items():
        src = tmp_path / f"{name}.lz4"
        src.write_bytes(b"H" * 41 + pa.compress(data, codec="lz4", asbytes=True))     # a kmtricks k-mer file header's worth of bytes first
        for chunk in (4 << 20, 65536, 1_000_003, 777):
            out = tmp_path / "out.bin"
            subprocess.check_call([harness, "decode", "41", str(chunk), str(src), str(out)])
            assert out.read_bytes() == data, (name, chunk)


def test_reader_rejects_damaged_frames(harness, tmp_path):
    data = _cases()["records12"]
    good = pa.compress(data, codec="lz4", asbytes=True)
    for what, blob in (("magic", b"\x00" + good[1:]), ("truncated", good[:len(good) // 2]),
                       ("offset before the start", good[:11] + b"\x10A\xff\xff" + good[15:])):
        src = tmp_path / "bad.lz4"
        src.write_bytes(blob)
        r = subprocess.run([harness, "decode", "0", "65536", str(src), str(tmp_path / "o.bin")], capture_output=True)
        assert r.returncode == 3, what


def test_writer_is_read_by_liblz4_and_by_the_reader(harness, tmp_path):
    for name in ("records12", "rows88", "random", "zeros", "mix", "empty"):
        data = _cases()[name]
        raw, enc, dec = tmp_path / "raw.bin", tmp_path / "enc.lz4", tmp_path / "dec.bin"
        raw.write_bytes(data)
        subprocess.check_call([harness, "encode", str(raw), str(enc)])
        got = pa.CompressedInputStream(pa.BufferReader(enc.read_bytes()), "lz4").read() if data else b""
        assert bytes(got) == data, name
        subprocess.check_call([harness, "decode", "0", "100000", str(enc), str(dec)])
        assert dec.read_bytes() == data, name
