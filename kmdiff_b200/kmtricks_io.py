"""kmtricks on-disk formats touched by `kmdiff diff` (reader + writer) and synthetic run dirs.

Formats (reference: thirdparty/kmtricks/include/kmtricks/, "km/" below):
  matrix  km/io/matrix_file.hpp:26-91 + km/io/io_common.hpp:125-158 — 45-byte header
          (u64 km_magic, u32 version, u8 compressed | u64 matrix_magic, u32 kmer_size,
          u32 kmer_slots, u32 count_slots, u32 nb_counts, u32 id, u32 partition) then AoS rows
          (raw, or one LZ4 frame when compressed=1).
  hist    km/io/hist_file.hpp:25-117 — 13-byte base header | u64 magic, u32 k, u32 id, u64 lower,
          upper, uniq, total, oob_ln, oob_lu, oob_un, oob_uu | u64 hist_u[upper-lower+1], hist_n[...]
  fof     km/io/fof.hpp:39-40 — "ID : path [! min_abundance]" per line, controls first.
  run dir SURVEY.md §8c: kmtricks.fof, options.txt ("kmer_size=K, c_ab_min=1,"), counts/partition_i/,
          histograms/<ID>.hist, matrices/matrix_<p>.count

The Python side only frames bytes; all arithmetic on the rows happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from dataclasses import dataclass

import numpy as np

from . import _lib

KM_MAGIC = 0x736B636972746D6B
MATRIX_MAGIC = 0x6B5F78697274616D
HIST_MAGIC = 0x747369686B
MATRIX_HEADER_BYTES = 45


@dataclass
class MatrixHeader:
    compressed: int
    kmer_size: int
    kmer_slots: int
    count_slots: int
    nb_counts: int
    id: int
    partition: int

    def pack(self) -> bytes:
        return struct.pack("<QIBQIIIIII", KM_MAGIC, 0, self.compressed, MATRIX_MAGIC, self.kmer_size,
                           self.kmer_slots, self.count_slots, self.nb_counts, self.id, self.partition)

    @staticmethod
    def unpack(b: bytes) -> "MatrixHeader":
        km, _ver, comp, mm, k, ks, cs, nb, i, p = struct.unpack("<QIBQIIIIII", b[:MATRIX_HEADER_BYTES])
        if km != KM_MAGIC or mm != MATRIX_MAGIC:
            raise IOError("Invalid file format.")  # km/io/io_common.hpp:150, matrix_file.hpp:80
        return MatrixHeader(comp, k, ks, cs, nb, i, p)


def write_matrix(path: str, payload, kmer_size: int, count_bytes: int, nb_counts: int,
                 partition: int = 0, lz4: bool = False) -> None:
    payload = np.ascontiguousarray(payload).view(np.uint8).ravel()
    hdr = MatrixHeader(int(lz4), kmer_size, (kmer_size + 31) // 32, count_bytes, nb_counts, 0, partition)
    with open(path, "wb") as f:
        f.write(hdr.pack())
        if lz4:
            import pyarrow as pa
            f.write(pa.compress(payload.tobytes(), codec="lz4", asbytes=True))
        else:
            payload.tofile(f)


def read_matrix(path: str) -> tuple[MatrixHeader, np.ndarray]:
    """-> (header, payload bytes).  The header's count_slots is ignored by the reference readers
    (km/io/matrix_file.hpp:149-150): the count width is the build's MAX_C, i.e. the caller's."""
    with open(path, "rb") as f:
        hdr = MatrixHeader.unpack(f.read(MATRIX_HEADER_BYTES))
        body = f.read()
    if hdr.compressed:
        import pyarrow as pa
        with pa.CompressedInputStream(pa.BufferReader(body), "lz4") as s:
            body = s.read()
    return hdr, np.frombuffer(body, dtype=np.uint8)


# ---- per-sample count files: counts/partition_{p}/<sample>.kmer[.lz4] (km/io/kmer_file.hpp) --------------

KMER_MAGIC = 0x72656D6B
KMER_HEADER_BYTES = 41


@dataclass
class KmerFileHeader:
    """km/io/kmer_file.hpp:26-67 (13-byte base header + 28 bytes)"""
    compressed: int
    kmer_size: int
    kmer_slots: int
    count_slots: int
    id: int
    partition: int

    def pack(self) -> bytes:
        return struct.pack("<QIBQIIIII", KM_MAGIC, 0, self.compressed, KMER_MAGIC, self.kmer_size,
                           self.kmer_slots, self.count_slots, self.id, self.partition)

    @staticmethod
    def unpack(b: bytes) -> "KmerFileHeader":
        km, _ver, comp, mm, k, ks, cs, i, p = struct.unpack("<QIBQIIIII", b[:KMER_HEADER_BYTES])
        if km != KM_MAGIC or mm != KMER_MAGIC:
            raise IOError("Invalid file format.")
        return KmerFileHeader(comp, k, ks, cs, i, p)


def write_kmer_file(path: str, records, kmer_size: int, count_bytes: int, sample_id: int = 0,
                    partition: int = 0, lz4: bool = False) -> None:
    """records: bytes of (kmer_slots*8 + count_bytes)-byte records, ascending k-mers (KmerWriter::write,
    km/io/kmer_file.hpp:94-101)."""
    payload = np.ascontiguousarray(records).view(np.uint8).ravel()
    hdr = KmerFileHeader(int(lz4), kmer_size, (kmer_size + 31) // 32, count_bytes, sample_id, partition)
    with open(path, "wb") as f:
        f.write(hdr.pack())
        if lz4:
            import pyarrow as pa
            f.write(pa.compress(payload.tobytes(), codec="lz4", asbytes=True))
        else:
            payload.tofile(f)


def read_kmer_file(path: str) -> tuple[KmerFileHeader, np.ndarray]:
    """-> (header, record bytes); the record's count width is the header's count_slots
    (KmerReader::read, km/io/kmer_file.hpp:127-137)."""
    with open(path, "rb") as f:
        hdr = KmerFileHeader.unpack(f.read(KMER_HEADER_BYTES))
        body = f.read()
    if hdr.compressed:
        import pyarrow as pa
        with pa.CompressedInputStream(pa.BufferReader(body), "lz4") as s:
            body = s.read()
    return hdr, np.frombuffer(body, dtype=np.uint8)


def get_partition_paths(run_dir: str, partition: int) -> list[str]:
    """KmDir::get_files_to_merge(p, true, KMER) (km/kmdir.hpp:48-73): one `.kmer.lz4` per fof sample, fof order."""
    fof = parse_fof(os.path.join(run_dir, "kmtricks.fof"))
    return [os.path.join(run_dir, "counts", f"partition_{partition}", f"{ident}.kmer.lz4") for ident, _ in fof]


def rows_to_streams(payload, nb_samples: int, kmer_words: int, count_bytes: int) -> list[np.ndarray]:
    """Split dense matrix rows into per-sample (k-mer, count) record streams, zero counts dropped —
    the inverse of the merge (what `kmtricks count` leaves per sample for these rows)."""
    R = 8 * kmer_words + nb_samples * count_bytes
    rows = np.ascontiguousarray(payload).view(np.uint8).reshape(-1, R)
    dt = {1: np.uint8, 2: "<u2", 4: "<u4"}[count_bytes]
    counts = np.ascontiguousarray(rows[:, 8 * kmer_words:]).view(dt).reshape(rows.shape[0], nb_samples)
    out = []
    for s in range(nb_samples):
        sel = counts[:, s] != 0
        rec = np.concatenate([rows[sel, :8 * kmer_words],
                              np.ascontiguousarray(counts[sel, s:s + 1]).view(np.uint8).reshape(-1, count_bytes)], axis=1)
        out.append(np.ascontiguousarray(rec).ravel())
    return out


def iter_kmer_ranges(streams, kmer_words: int, count_bytes: int, max_records: int):
    """Cut N sorted per-sample record streams into complete k-mer ranges of at most ~max_records records (all
    samples together), the way `kmdiff-b200 diff` feeds kmd_merge_scan_host (GpuWorker::merge_loop): every stream
    contributes up to max_records/N buffered records, the smallest "last buffered k-mer" of the streams that still
    have more bounds the range, records above it wait for the next one.  Yields lists of N uint8 views."""
    rec = 8 * kmer_words + count_bytes
    recs = [np.ascontiguousarray(s).view(np.uint8).reshape(-1, rec) for s in streams]
    n = len(recs)
    per = max(1, max_records // max(1, n))

    def keys(a):   # ascending sort key per record: (high word, low word)
        lo = np.ascontiguousarray(a[:, :8]).view("<u8").ravel()
        if kmer_words == 1:
            return lo, None
        return lo, np.ascontiguousarray(a[:, 8:16]).view("<u8").ravel()

    ks = [keys(a) for a in recs]
    pos = [0] * n
    while True:
        end = [min(len(recs[s]), pos[s] + per) for s in range(n)]
        if all(end[s] == pos[s] for s in range(n)):
            return
        bound = None
        for s in range(n):
            if end[s] < len(recs[s]):                 # the stream has more: nothing above its last buffered k-mer is complete
                lo, hi = ks[s]
                k = (int(hi[end[s] - 1]) if hi is not None else 0, int(lo[end[s] - 1]))
                if bound is None or k < bound:
                    bound = k
        cut = []
        for s in range(n):
            if bound is None:
                cut.append(end[s])
                continue
            lo, hi = ks[s]
            a, b = pos[s], end[s]
            if hi is None:
                c = a + int(np.searchsorted(lo[a:b], np.uint64(bound[1]), side="right"))
            else:
                c = a + int(np.searchsorted(hi[a:b], np.uint64(bound[0]), side="left"))
                c2 = a + int(np.searchsorted(hi[a:b], np.uint64(bound[0]), side="right"))
                c = c + int(np.searchsorted(lo[c:c2], np.uint64(bound[1]), side="right"))
            cut.append(c)
        yield [recs[s][pos[s]:cut[s]].ravel() for s in range(n)]
        pos = cut


def write_hist(path: str, kmer_size: int, sample_id: int, total: int, lower: int = 1, upper: int = 255,
               hist_u: np.ndarray | None = None) -> None:
    n = upper - lower + 1
    hu = np.zeros(n, dtype=np.uint64) if hist_u is None else np.asarray(hist_u, dtype=np.uint64)
    with open(path, "wb") as f:
        f.write(struct.pack("<QIB", KM_MAGIC, 0, 0))
        f.write(struct.pack("<QIIQQQQQQQQ", HIST_MAGIC, kmer_size, sample_id, lower, upper, 0, total, 0, 0, 0, 0))
        hu.tofile(f)
        np.zeros(n, dtype=np.uint64).tofile(f)


def read_hist_total(path: str, ab_min: int = 1) -> int:
    """get_total_kmer for one sample (reference src/kmtricks_utils.cpp:93-107):
    total - sum_{j=1}^{ab_min-1} j * hist_u[j-1]."""
    with open(path, "rb") as f:
        km, _v, comp = struct.unpack("<QIB", f.read(13))
        magic, _k, _id, lower, upper, _uniq, total, *_ = struct.unpack("<QIIQQQQQQQQ", f.read(80))
        if km != KM_MAGIC or magic != HIST_MAGIC:
            raise IOError("Invalid file format.")
        body = f.read()
    if comp:
        import pyarrow as pa
        with pa.CompressedInputStream(pa.BufferReader(body), "lz4") as s:
            body = s.read()
    hu = np.frombuffer(body, dtype=np.uint64, count=upper - lower + 1)
    for j in range(1, ab_min):
        total -= j * int(hu[j - 1])
    return total


def parse_fof(path: str) -> list[tuple[str, int]]:
    """[(sample id, min abundance or 0)], in file order (km/io/fof.hpp)."""
    out = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            ident, rest = line.split(":", 1)
            ab = 0
            if "!" in rest:
                ab = int(rest.split("!")[1].strip() or 0)
            out.append((ident.strip(), ab))
    return out


@dataclass
class RunConfig:
    kmer_size: int
    abundance_min: int
    nb_partitions: int


def get_kmtricks_config(run_dir: str) -> RunConfig:
    """reference src/kmtricks_utils.cpp:29-76"""
    path = os.path.join(run_dir, "kmdiff-count.opt")
    key = "abundance_min"
    if not os.path.exists(path):
        path = os.path.join(run_dir, "options.txt")
        key = "c_ab_min"
    k = ab = 0
    with open(path) as f:
        for line in f:
            if "kmer_size" in line:
                for o in line.split(","):
                    o = o.strip()
                    if "kmer_size" in o:
                        k = int(o.split("=")[1])
                    if key in o:
                        ab = int(o.split("=")[1])
    counts = os.path.join(run_dir, "counts")
    npart = len(os.listdir(counts)) if os.path.isdir(counts) else 0
    if not k or not npart:
        raise RuntimeError(f"ConfigError - Unable to load config from {path}.")
    return RunConfig(k, ab, npart)


def get_total_kmer(run_dir: str, nb_controls: int, nb_cases: int, abundance_min: int):
    """reference src/kmtricks_utils.cpp:78-139 -> (total_controls, total_cases)"""
    fof = parse_fof(os.path.join(run_dir, "kmtricks.fof"))
    totals = []
    for ident, ab in fof[:nb_controls + nb_cases]:
        totals.append(read_hist_total(os.path.join(run_dir, "histograms", f"{ident}.hist"), ab or abundance_min))
    return totals[:nb_controls], totals[nb_controls:]


def list_matrices(run_dir: str) -> list[str]:
    d = os.path.join(run_dir, "matrices")
    if not os.path.isdir(d):
        return []
    return sorted(os.path.join(d, f) for f in os.listdir(d))


# ---- synthetic cohorts ------------------------------------------------------------------------------

@dataclass
class SynthSpec:
    nb_controls: int
    nb_cases: int
    kmer_size: int = 31
    count_bytes: int = 4
    total_rows: int = 1 << 20
    seed: int = 42
    enriched_ppm: int = 1000

    @property
    def kmer_words(self) -> int:
        return (self.kmer_size + 31) // 32

    @property
    def row_bytes(self) -> int:
        return 8 * self.kmer_words + (self.nb_controls + self.nb_cases) * self.count_bytes

    def c_struct(self) -> "_lib.Synth":
        return _lib.Synth(self.seed, self.total_rows, self.nb_controls, self.nb_cases, self.kmer_size,
                          self.kmer_words, self.count_bytes, self.enriched_ppm)

    def totals(self) -> np.ndarray:
        t = np.empty(self.nb_controls + self.nb_cases, dtype=np.uint64)
        s = self.c_struct()
        _lib.load().kmd_synth_totals(C.byref(s), t.ctypes.data)
        return t

    def rows_host(self, n_rows: int, row_base: int = 0, threads: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        """rows [row_base, row_base+n_rows) of the synthetic matrix as a flat uint8 payload"""
        if out is None:
            out = np.empty(n_rows * self.row_bytes, dtype=np.uint8)
        s = self.c_struct()
        rc = _lib.load().kmd_synth_fill_host(C.byref(s), out.ctypes.data, n_rows, row_base,
                                             threads or (os.cpu_count() or 1))
        if rc:
            raise RuntimeError(f"kmd_synth_fill_host failed: {rc}")
        return out

    def rows_device(self, dev_ptr: int, n_rows: int, row_base: int = 0, device: int = 0, stream: int = 0) -> None:
        s = self.c_struct()
        rc = _lib.load().kmd_synth_fill_device(C.byref(s), device, C.c_void_p(dev_ptr), n_rows, row_base,
                                               C.c_void_p(stream))
        if rc:
            raise RuntimeError(f"kmd_synth_fill_device failed: {rc}")


def write_run_dir(run_dir: str, spec: SynthSpec, n_rows: int, nb_partitions: int, lz4: bool = False,
                  payload: np.ndarray | None = None, totals: np.ndarray | None = None,
                  count_files: bool = False) -> list[str]:
    """A kmtricks run directory `kmdiff diff` (reference and ours) accepts in matrix mode, holding
    rows [0, n_rows) of the synthetic matrix split into nb_partitions contiguous partitions."""
    os.makedirs(os.path.join(run_dir, "matrices"), exist_ok=True)
    os.makedirs(os.path.join(run_dir, "histograms"), exist_ok=True)
    n = spec.nb_controls + spec.nb_cases
    ids = [f"C{i}" for i in range(spec.nb_controls)] + [f"K{i}" for i in range(spec.nb_cases)]
    with open(os.path.join(run_dir, "kmtricks.fof"), "w") as f:
        for ident in ids:
            f.write(f"{ident} : /dev/null/{ident}.fa\n")
    with open(os.path.join(run_dir, "options.txt"), "w") as f:
        f.write(f"kmer_size={spec.kmer_size}, c_ab_min=1,\n")
    if totals is None:
        totals = spec.totals()
    for j, ident in enumerate(ids):
        write_hist(os.path.join(run_dir, "histograms", f"{ident}.hist"), spec.kmer_size, j, int(totals[j]))
    stale = os.path.join(run_dir, "kmdiff-count.opt")
    if os.path.exists(stale):
        os.remove(stale)
    paths = []
    per = (n_rows + nb_partitions - 1) // nb_partitions
    for p in range(nb_partitions):
        os.makedirs(os.path.join(run_dir, "counts", f"partition_{p}"), exist_ok=True)
        lo, hi = min(n_rows, p * per), min(n_rows, (p + 1) * per)
        if payload is None:
            rows = spec.rows_host(hi - lo, lo)
        else:
            rows = payload[lo * spec.row_bytes:hi * spec.row_bytes]
        if count_files:
            # what `kmtricks pipeline --until count` leaves: one sorted (k-mer, count) file per sample
            # and partition, no matrices (the reference then merges them itself, cmd/diff.hpp:88-95)
            streams = rows_to_streams(rows, n, spec.kmer_words, spec.count_bytes)
            for j, ident in enumerate(ids):
                path = os.path.join(run_dir, "counts", f"partition_{p}", f"{ident}.kmer.lz4")
                write_kmer_file(path, streams[j], spec.kmer_size, spec.count_bytes, j, p, lz4)
                paths.append(path)
            continue
        path = os.path.join(run_dir, "matrices", f"matrix_{p}.count" + (".lz4" if lz4 else ""))
        write_matrix(path, rows, spec.kmer_size, spec.count_bytes, n, p, lz4)
        paths.append(path)
    return paths
