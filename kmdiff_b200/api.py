"""Python mirror of the reference's operator interface for the `kmdiff diff` path, over the C-ABI.

Names and argument meaning follow the reference (file:line relative to the reference root):

    LogFactorialTable      include/kmdiff/log_factorial_table.hpp:9-26
    PoissonLikelihood      include/kmdiff/model.hpp:93-192   (IModel::process, imodel.hpp:37)
    Significance           include/kmdiff/kmer.hpp:33-38
    CorrectionType         include/kmdiff/correction.hpp
    bonferroni / benjamini / sidak / holm / basic_threshold, make_corrector
                           include/kmdiff/corrector.hpp, src/corrector.cpp
    DiffContext            global_merge + diff_observer + accumulators + aggregators
                           (include/kmdiff/merge.hpp, accumulator.hpp, aggregator.hpp)

Everything that computes runs in the CUDA library; nothing here falls back to numpy.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import SURVIVOR_DTYPE, Config, Counters, PopstratConfig


class KmdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"kmdiff_b200 error {code}: {msg}")
        self.code = code


class Significance(enum.IntEnum):
    CONTROL = 0
    CASE = 1
    NO = 2


class CorrectionType(enum.IntEnum):
    NOTHING = 0
    BONFERRONI = 1
    BENJAMINI = 2
    SIDAK = 3
    HOLM = 4


def correction_from_str(s: str) -> CorrectionType:
    """-c/--correction values of the reference CLI (src/cli.cpp)."""
    table = {"bonferroni": CorrectionType.BONFERRONI, "benjamini": CorrectionType.BENJAMINI,
             "sidak": CorrectionType.SIDAK, "holm": CorrectionType.HOLM,
             "disabled": CorrectionType.NOTHING}
    if s not in table:
        raise ValueError(f"unknown correction '{s}'")
    return table[s]


class LogFactorialTable:
    """table[i] = log(i) + log(i-1) + ... + log(2); beyond the table the value is recomputed."""

    def __init__(self, size: int):
        self.size = int(size)
        self.table = np.empty(self.size, dtype=np.float64)
        _lib.load().kmd_lf_table_fill(self.table.ctypes.data, self.size)

    def __getitem__(self, i: int) -> float:
        if i < self.size:
            return float(self.table[i])
        t = np.empty(i + 1, dtype=np.float64)
        _lib.load().kmd_lf_table_fill(t.ctypes.data, i + 1)
        return float(t[i])


# ---- correctors (host predicates, stateful like the reference) -----------------------------------

class ICorrector:
    kind = CorrectionType.NOTHING

    def __init__(self, threshold: float, total: int = 0):
        self.threshold = float(threshold)
        self.total = int(total)
        self._applied = 0   # benjamini: accepted so far; holm: applications so far

    def type(self) -> CorrectionType:
        return self.kind

    def str_type(self) -> str:
        return {0: "threshold", 1: "bonferroni", 2: "benjamini", 3: "sidak", 4: "holm"}[int(self.kind)]

    def apply(self, pvalue: float) -> bool:
        L = _lib.load()
        rank = self._applied + 1 if self.kind == CorrectionType.BENJAMINI else self._applied
        ok = bool(L.kmd_corrector_apply(int(self.kind), self.threshold, self.total, rank, pvalue))
        if self.kind == CorrectionType.BENJAMINI:
            self._applied += 1 if ok else 0
        elif self.kind == CorrectionType.HOLM:
            self._applied += 1
        return ok


class basic_threshold(ICorrector):
    kind = CorrectionType.NOTHING


class bonferroni(ICorrector):
    kind = CorrectionType.BONFERRONI


class benjamini(ICorrector):
    kind = CorrectionType.BENJAMINI


class sidak(ICorrector):
    kind = CorrectionType.SIDAK


class holm(ICorrector):
    kind = CorrectionType.HOLM


def make_corrector(kind: CorrectionType, threshold: float, kmers: int) -> ICorrector:
    return {CorrectionType.BONFERRONI: bonferroni, CorrectionType.SIDAK: sidak,
            CorrectionType.BENJAMINI: benjamini, CorrectionType.HOLM: holm,
            CorrectionType.NOTHING: basic_threshold}[CorrectionType(kind)](threshold, kmers)


# ---- the device context ---------------------------------------------------------------------------

def _as_u8(buf) -> np.ndarray:
    a = np.asarray(buf)
    if a.dtype != np.uint8:
        a = a.view(np.uint8)
    return np.ascontiguousarray(a).ravel()


class DiffContext:
    """One GPU's share of `kmdiff diff`: the partition loop (scan), the survivor accumulators and
    the correction, all device-side behind the C-ABI."""

    def __init__(self, nb_controls: int, nb_cases: int, sum_controls: int, sum_cases: int,
                 kmer_size: int = 31, count_bytes: int = 4, threshold: float = 0.05 / 100000,
                 lf_size: int = 10000, lf_table: np.ndarray | None = None, device: int = 0,
                 max_survivors: int = 0, keep_counts: bool = False):
        self._L = _lib.load()
        self.nb_controls, self.nb_cases = int(nb_controls), int(nb_cases)
        self.kmer_size = int(kmer_size)
        self.kmer_words = (self.kmer_size + 31) // 32
        self.count_bytes = int(count_bytes)
        self.row_bytes = 8 * self.kmer_words + (self.nb_controls + self.nb_cases) * self.count_bytes
        self.threshold = float(threshold)
        self.device = int(device)
        self._table = None if lf_table is None else np.ascontiguousarray(lf_table, dtype=np.float64)
        cfg = Config(_lib.KMD_ABI_VERSION, self.device, self.nb_controls, self.nb_cases,
                     self.kmer_size, self.kmer_words, self.count_bytes,
                     _lib.FLAG_KEEP_COUNTS if keep_counts else 0, int(sum_controls),
                     int(sum_cases), self.threshold,
                     None if self._table is None else self._table.ctypes.data_as(C.POINTER(C.c_double)),
                     int(lf_size if self._table is None else self._table.size), int(max_survivors))
        h = C.c_void_p()
        rc = self._L.kmd_create(C.byref(cfg), C.byref(h))
        if rc:
            raise KmdError(rc, self._L.kmd_last_error(None).decode())
        self._h = h

    # -- plumbing --
    def _ck(self, rc: int) -> None:
        if rc:
            raise KmdError(rc, self._L.kmd_last_error(self._h).decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.kmd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream: int | None) -> None:
        self._ck(self._L.kmd_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def sync(self) -> None:
        self._ck(self._L.kmd_sync(self._h))

    def reset(self) -> None:
        self._ck(self._L.kmd_reset(self._h))

    # -- partition loop --
    def scan_device(self, dev_ptr: int, n_rows: int, row_base: int = 0, partition: int = 0) -> None:
        self._ck(self._L.kmd_scan_device(self._h, C.c_void_p(dev_ptr), n_rows, row_base, partition))

    def scan_host(self, rows, n_rows: int | None = None, row_base: int = 0, partition: int = 0) -> None:
        """rows: bytes-like / uint8 array holding a matrix payload, or an integer host address."""
        if isinstance(rows, int):
            assert n_rows is not None
            ptr = rows
        else:
            a = _as_u8(rows)
            if n_rows is None:
                assert a.size % self.row_bytes == 0, "payload is not a whole number of rows"
                n_rows = a.size // self.row_bytes
            ptr = a.ctypes.data
        self._ck(self._L.kmd_scan_host(self._h, C.c_void_p(ptr), n_rows, row_base, partition))

    def scan_host_async(self, host_ptr: int, n_rows: int, row_base: int = 0, partition: int = 0) -> int:
        """kmd_scan_host without waiting for the copies; host_ptr (pinned) stays untouched until host_wait(ticket)"""
        t = C.c_uint64()
        self._ck(self._L.kmd_scan_host_async(self._h, C.c_void_p(host_ptr), n_rows, row_base, partition, C.byref(t)))
        return t.value

    def host_wait(self, ticket: int) -> None:
        rc = self._L.kmd_host_wait(self._h, int(ticket))
        if rc:
            raise KmdError(rc, "kmd_host_wait failed")

    # -- per-sample count files: km::KmerMerger (kmtricks/merge.hpp:102-361) --
    def _merge_args(self, streams, ab_min):
        n = self.nb_samples
        assert len(streams) == n, f"{n} streams expected (controls first)"
        rec = 8 * self.kmer_words + self.count_bytes
        arrs = [_as_u8(s) for s in streams]
        for a in arrs:
            assert a.size % rec == 0, "stream is not a whole number of records"
        ptrs = (C.c_void_p * n)(*[a.ctypes.data if a.size else None for a in arrs])
        cnt = (C.c_uint64 * n)(*[a.size // rec for a in arrs])
        ab = np.ascontiguousarray(np.broadcast_to(np.asarray(ab_min, dtype=np.uint32), (n,)))
        return arrs, ptrs, cnt, ab

    def merge_scan_host(self, streams, ab_min=1, row_base: int = 0, partition: int = 0) -> int:
        """global_merge::merge() for one partition (merge.hpp:259-301): N-way merge of the samples'
        sorted (k-mer, count) records on the device + the LRT scan of the merged rows; -> rows."""
        arrs, ptrs, cnt, ab = self._merge_args(streams, ab_min)
        n = C.c_uint64()
        self._ck(self._L.kmd_merge_scan_host(self._h, ptrs, cnt, ab.ctypes.data, row_base, partition, C.byref(n)))
        return n.value

    def merge_scan_device(self, d_records: int, n_records, ab_min=1, row_base: int = 0, partition: int = 0) -> int:
        """merge_scan_host for records that already lie in device memory: d_records = device address of the samples'
        records back to back, n_records[s] = records of sample s; -> rows"""
        n = self.nb_samples
        assert len(n_records) == n, f"{n} streams expected (controls first)"
        cnt = (C.c_uint64 * n)(*[int(c) for c in n_records])
        ab = np.ascontiguousarray(np.broadcast_to(np.asarray(ab_min, dtype=np.uint32), (n,)))
        rows = C.c_uint64()
        self._ck(self._L.kmd_merge_scan_device(self._h, C.c_void_p(int(d_records)), cnt, ab.ctypes.data, row_base, partition,
                                               C.byref(rows)))
        return rows.value

    def merge_info(self) -> dict:
        """the last merge call: 'algo' ('union' = k-mers merged with duplicates dropped + counts scattered,
        'pairwise' = (k-mer, payload) merge tree), 'rows' ('assembled' in shared memory / 'scattered' into zeroed rows),
        keys read / written by the union tree's levels"""
        algo, mode, kin, kout = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint64()
        self._ck(self._L.kmd_merge_info(self._h, C.byref(algo), C.byref(mode), C.byref(kin), C.byref(kout)))
        return {"algo": "union" if algo.value == 0 else "pairwise", "rows": {0: None, 1: "assembled", 2: "scattered"}[mode.value],
                "keys_in": kin.value, "keys_out": kout.value}

    def _frame_args(self, frames):
        n = self.nb_samples
        assert len(frames) == n, f"{n} frames expected (controls first)"
        arrs = [_as_u8(f) for f in frames]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data if a.size else None for a in arrs])
        sizes = (C.c_uint64 * n)(*[a.size for a in arrs])
        return arrs, ptrs, sizes

    def merge_scan_lz4_host(self, frames, compressed=None, ab_min=1, row_base: int = 0, partition: int = 0,
                            max_records: int = 0) -> tuple[int, int]:
        """one partition's count files as they lie on disk (payloads after the 41-byte header; LZ4 frames are decoded
        on the device) -> (rows, records)"""
        arrs, ptrs, sizes = self._frame_args(frames)
        ab = np.ascontiguousarray(np.broadcast_to(np.asarray(ab_min, dtype=np.uint32), (self.nb_samples,)))
        comp = None if compressed is None else np.ascontiguousarray(compressed, dtype=np.uint8)
        rows, recs = C.c_uint64(), C.c_uint64()
        self._ck(self._L.kmd_merge_scan_lz4_host(self._h, ptrs, sizes, None if comp is None else comp.ctypes.data, ab.ctypes.data,
                                                 row_base, partition, int(max_records), C.byref(rows), C.byref(recs)))
        return rows.value, recs.value

    def lz4_decode_host(self, frames, cap_records: int) -> tuple[np.ndarray, float]:
        """the device LZ4-frame decoder alone: -> (decoded records of all streams back to back as bytes, device ms)"""
        arrs, ptrs, sizes = self._frame_args(frames)
        rec = 8 * self.kmer_words + self.count_bytes
        out = np.zeros(max(1, cap_records) * rec, dtype=np.uint8)
        n, ms = C.c_uint64(), C.c_float()
        self._ck(self._L.kmd_lz4_decode_host(self._h, ptrs, sizes, out.ctypes.data, cap_records, C.byref(n), C.byref(ms)))
        return out[:n.value * rec], ms.value

    def merge_prefetch(self, blobs, frames: bool = False):
        """start the H2D copies of the next merge_scan_host (record streams) / merge_scan_lz4_host (frames) call on the
        copy stream; returns the arrays whose memory must stay untouched (and be passed again) until that call"""
        if frames:
            arrs, ptrs, sizes = self._frame_args(blobs)
        else:
            arrs, ptrs, cnt, _ = self._merge_args(blobs, 1)
            rec = 8 * self.kmer_words + self.count_bytes
            sizes = (C.c_uint64 * self.nb_samples)(*[c * rec for c in cnt])
        self._ck(self._L.kmd_merge_prefetch(self._h, ptrs, sizes, int(bool(frames))))
        return arrs

    def merge_ms(self) -> dict:
        """device times of the last merge call: {'unpack', 'passes', 'rows'} in ms, 'n_passes', 'n_records'"""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        npass, nrec = C.c_uint32(), C.c_uint64()
        self._ck(self._L.kmd_merge_ms(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(npass), C.byref(nrec)))
        return {"unpack": a.value, "passes": b.value, "rows": c.value, "n_passes": npass.value, "n_records": nrec.value}

    def merge_rows_host(self, streams, ab_min=1) -> np.ndarray:
        """the merged dense rows (KmerMerger::write_as_bin payload), [n_rows, row_bytes] uint8"""
        arrs, ptrs, cnt, ab = self._merge_args(streams, ab_min)
        cap = int(sum(cnt))
        out = np.zeros((max(cap, 1), self.row_bytes), dtype=np.uint8)
        n = C.c_uint64()
        self._ck(self._L.kmd_merge_rows_host(self._h, ptrs, cnt, ab.ctypes.data, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value]

    def counters(self) -> Counters:
        c = Counters()
        self._ck(self._L.kmd_counters_get(self._h, C.byref(c)))
        return c

    def launch_count(self) -> int:
        return int(self._L.kmd_launch_count(self._h))

    def scan_kernel_ms(self) -> float:
        ms = C.c_float()
        self._ck(self._L.kmd_scan_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def survivors(self) -> np.ndarray:
        n = C.c_uint64()
        self._ck(self._L.kmd_survivors(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=SURVIVOR_DTYPE)
        if n.value:
            self._ck(self._L.kmd_survivors(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def push_survivors(self, records: np.ndarray, rows_tested: int = 0, counts: np.ndarray | None = None) -> None:
        """append host-side survivor records (plugin models, resumed spill files), optionally with their counts"""
        rec = np.ascontiguousarray(records, dtype=SURVIVOR_DTYPE)
        if counts is None:
            self._ck(self._L.kmd_push_survivors(self._h, rec.ctypes.data, rec.size, int(rows_tested)))
            return
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[self.count_bytes]
        c = np.ascontiguousarray(counts, dtype=dt).reshape(rec.size, self.nb_controls + self.nb_cases)
        self._ck(self._L.kmd_push_survivors2(self._h, rec.ctypes.data, c.ctypes.data, rec.size, int(rows_tested)))

    # -- IModel::process, batched --
    def model_rows(self, rows):
        a = _as_u8(rows)
        assert a.size % self.row_bytes == 0
        n = a.size // self.row_bytes
        p = np.empty(n, np.float64)
        sign = np.empty(n, np.int32)
        mc = np.empty(n, np.float64)
        kk = np.empty(n, np.float64)
        self._ck(self._L.kmd_model_rows_host(self._h, a.ctypes.data, n, p.ctypes.data, sign.ctypes.data,
                                             mc.ctypes.data, kk.ctypes.data))
        return p, sign, mc, kk

    # -- correction --
    def correct(self, correction: int, significance: float, total_rows: int) -> tuple[int, int]:
        nc, nk = C.c_uint64(), C.c_uint64()
        self._ck(self._L.kmd_correct(self._h, int(correction), significance, total_rows, C.byref(nc), C.byref(nk)))
        return nc.value, nk.value

    def significant(self) -> np.ndarray:
        n = C.c_uint64()
        self._ck(self._L.kmd_significant(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=SURVIVOR_DTYPE)
        if n.value:
            self._ck(self._L.kmd_significant(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def sorted_pvalues(self) -> np.ndarray:
        n = C.c_uint64()
        self._ck(self._L.kmd_sorted_pvalues(self._h, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.float64)
        if n.value:
            self._ck(self._L.kmd_sorted_pvalues(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def survivor_pvalues_device(self, dev_ptr: int, cap: int) -> int:
        """survivors' p-values (unsorted) copied device-to-device into dev_ptr on the context's stream; -> count"""
        n = C.c_uint64()
        self._ck(self._L.kmd_survivor_pvalues_device(self._h, C.c_void_p(dev_ptr) if dev_ptr else None, cap, C.byref(n)))
        return n.value

    def first_reject_device(self, dev_ptr: int, n: int, correction: int, significance: float, total_rows: int):
        """sorted_aggregator::run's walk over a device array of merged ascending p-values; -> (n_accepted, p_cut)"""
        acc, cut = C.c_uint64(), C.c_double()
        self._ck(self._L.kmd_first_reject_device(self._h, C.c_void_p(dev_ptr), n, int(correction), significance, total_rows,
                                                 C.byref(acc), C.byref(cut)))
        return acc.value, cut.value

    # -- multi-GPU correction, stream-ordered on the device (kmd_exchange_*) --
    def stream_ptr(self) -> int:
        """the cudaStream_t (as an integer) all device work of this context is ordered on"""
        return int(self._L.kmd_get_stream(self._h) or 0)

    def exchange_export(self, slot_dev_ptr: int, cap: int) -> None:
        self._ck(self._L.kmd_exchange_export(self._h, C.c_void_p(slot_dev_ptr), int(cap)))

    def exchange_correct(self, all_dev_ptr: int, world: int, cap: int, correction: int, significance: float) -> None:
        self._ck(self._L.kmd_exchange_correct(self._h, C.c_void_p(all_dev_ptr), int(world), int(cap), int(correction),
                                              float(significance)))

    def exchange_result(self) -> _lib.ExchangeDecision:
        out = _lib.ExchangeDecision()
        self._ck(self._L.kmd_exchange_result(self._h, C.byref(out)))
        return out

    def keep_below(self, p_cut: float) -> tuple[int, int]:
        nc, nk = C.c_uint64(), C.c_uint64()
        self._ck(self._L.kmd_keep_below(self._h, p_cut, C.byref(nc), C.byref(nk)))
        return nc.value, nk.value


    # -- population stratification (--pop-correction) --
    @property
    def nb_samples(self) -> int:
        return self.nb_controls + self.nb_cases

    def popstrat_enable(self, kmer_pca: float, seed: int = 0, is_diploid: bool = True, n_pc: int = 2,
                        max_sampled: int = 0) -> None:
        """Sampler(geno, snp, kmer_pca, seed) (popstrat.hpp:107-118): every later scan also samples."""
        cfg = PopstratConfig(float(kmer_pca), int(seed), int(bool(is_diploid)), int(n_pc), int(max_sampled))
        self._ck(self._L.kmd_popstrat_enable(self._h, C.byref(cfg)))
        self.n_pc = int(n_pc)

    def popstrat_set_draw_index(self, next_draw: int) -> None:
        """The next scanned row takes Sampler draw `next_draw`, whatever row_base the scans are given."""
        self._ck(self._L.kmd_popstrat_set_draw_index(self._h, int(next_draw)))

    def popstrat_sampled(self) -> int:
        n = C.c_uint64()
        self._ck(self._L.kmd_popstrat_sampled(self._h, C.byref(n)))
        return n.value

    def popstrat_geno(self) -> tuple[np.ndarray, np.ndarray]:
        """(geno [m, N] 0/1, global row index of each sampled k-mer), in row order (the .geno file)."""
        m = self.popstrat_sampled()
        geno = np.zeros((m, self.nb_samples), dtype=np.uint8)
        rows = np.zeros(m, dtype=np.uint64)
        n = C.c_uint64()
        if m:
            self._ck(self._L.kmd_popstrat_geno(self._h, geno.ctypes.data, rows.ctypes.data, m, C.byref(n)))
        return geno, rows

    def popstrat_gram(self, normalize: bool = True) -> tuple[np.ndarray, int]:
        """smartpca's XTX of the sampled k-mers (int8 tensor-core Gram + fp64 fix-up); -> (xtx, n_used)."""
        xtx = np.zeros((self.nb_samples, self.nb_samples), dtype=np.float64)
        used = C.c_uint64()
        self._ck(self._L.kmd_popstrat_gram(self._h, xtx.ctypes.data, int(normalize), C.byref(used)))
        return xtx, used.value

    def popstrat_gram_partial(self, g_dev_ptr: int, hist_dev_ptr: int) -> None:
        """integer Gram matrices by carrier count ([N, N, N] int32) + carrier histogram ([N + 1] u64) into DEVICE memory"""
        self._ck(self._L.kmd_popstrat_gram_partial(self._h, C.c_void_p(g_dev_ptr), C.c_void_p(hist_dev_ptr)))

    def popstrat_gram_combine(self, g_dev_ptr: int, hist_dev_ptr: int, normalize: bool = True) -> tuple[np.ndarray, int]:
        """XTX from (globally summed) integer Gram matrices and histogram; -> (xtx, n_used)"""
        xtx = np.zeros((self.nb_samples, self.nb_samples), dtype=np.float64)
        used = C.c_uint64()
        self._ck(self._L.kmd_popstrat_gram_combine(self._h, C.c_void_p(g_dev_ptr), C.c_void_p(hist_dev_ptr), xtx.ctypes.data,
                                                   int(normalize), C.byref(used)))
        return xtx, used.value

    def popstrat_gram_ms(self) -> float:
        ms = C.c_float()
        self._ck(self._L.kmd_popstrat_gram_ms(self._h, C.byref(ms)))
        return ms.value

    def popstrat_init(self, Z: np.ndarray, totals: np.ndarray) -> np.ndarray:
        """pop_strat_corrector::init_global_features; Z: [N, 10]; returns the null model weights."""
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        assert Z.shape == (self.nb_samples, 10)
        t = np.ascontiguousarray(totals, dtype=np.uint64)
        assert t.size == self.nb_samples
        null_model = np.zeros(self.n_pc + 2, dtype=np.float64)
        self._ck(self._L.kmd_popstrat_init(self._h, Z.ctypes.data, t.ctypes.data, null_model.ctypes.data))
        return null_model

    def popstrat_apply(self) -> None:
        """pop_strat_corrector::apply over the accumulators: survivors' p-values become GLM p-values."""
        self._ck(self._L.kmd_popstrat_apply(self._h))

    def popstrat_fit(self, counts: np.ndarray):
        """pop_strat_corrector::apply(KmerSign&) on [n_kmers, N] counts -> (p, model, llr, iters)."""
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[self.count_bytes]
        c = np.ascontiguousarray(counts, dtype=dt).reshape(-1, self.nb_samples)
        m = c.shape[0]
        p = np.zeros(m, np.float64)
        model = np.zeros((m, self.n_pc + 3), np.float64)
        llr = np.zeros(m, np.float64)
        iters = np.zeros(m, np.int32)
        self._ck(self._L.kmd_popstrat_fit(self._h, c.ctypes.data, m, p.ctypes.data, model.ctypes.data,
                                          llr.ctypes.data, iters.ctypes.data))
        return p, model, llr, iters

    def survivor_counts(self) -> np.ndarray:
        """[n_survivors, N] counts of the prefilter survivors, row order like survivors()."""
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[self.count_bytes]
        n = C.c_uint64()
        self._ck(self._L.kmd_survivor_counts(self._h, None, 0, C.byref(n)))
        out = np.zeros((n.value, self.nb_samples), dtype=dt)
        if n.value:
            self._ck(self._L.kmd_survivor_counts(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out


def sym_eig(a: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """eigenvalues (descending) and eigenvectors (rows) of a symmetric matrix (host helper)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    ev = np.zeros(n)
    vec = np.zeros((n, n))
    rc = _lib.load().kmd_sym_eig(a.ctypes.data, n, ev.ctypes.data, vec.ctypes.data)
    if rc:
        raise KmdError(rc, "eigensolver did not converge")
    return ev, vec


def smartpca_pcs(xtx: np.ndarray, numeigs: int = 10) -> tuple[np.ndarray, np.ndarray]:
    """(eigenvalues, pcs [N, numeigs]) as pcs.evec holds them after "%10.4f" (host helper)."""
    x = np.ascontiguousarray(xtx, dtype=np.float64)
    n = x.shape[0]
    ev = np.zeros(n)
    pcs = np.zeros((n, numeigs))
    rc = _lib.load().kmd_smartpca_pcs(x.ctypes.data, n, numeigs, ev.ctypes.data, pcs.ctypes.data)
    if rc:
        raise KmdError(rc, "smartpca_pcs failed")
    return ev, pcs


def sampler_rows(seed: int, kmer_pca: float, first_row: int, n: int) -> np.ndarray:
    """Sampler decisions of rows [first_row, first_row + n) (host helper, jump-ahead LCG)."""
    out = np.zeros(n, dtype=np.uint8)
    _lib.load().kmd_sampler_rows(int(seed), float(kmer_pca), int(first_row), int(n), out.ctypes.data)
    return out


def first_reject(p_sorted: np.ndarray, correction: int, significance: float, total_rows: int):
    """Walk merged ascending p-values like sorted_aggregator::run; -> (n_accepted, p_cut)."""
    p = np.ascontiguousarray(p_sorted, dtype=np.float64)
    cut = C.c_double()
    n = _lib.load().kmd_first_reject(p.ctypes.data, p.size, int(correction), significance, total_rows, C.byref(cut))
    return int(n), cut.value


class PoissonLikelihood:
    """IModel implementation of the reference (model.hpp:93-192) evaluated on the GPU.

    process(controls, cases) keeps the reference's per-row signature; process_rows() is the batched
    form the GPU is meant for."""

    def __init__(self, nb_controls: int, nb_cases: int, total_controls: Sequence[int],
                 total_cases: Sequence[int], preload: int = 10000, count_bytes: int = 4,
                 kmer_size: int = 31, device: int = 0):
        assert len(total_controls) == nb_controls and len(total_cases) == nb_cases
        self.nb_controls, self.nb_cases = nb_controls, nb_cases
        self.count_bytes = count_bytes
        self._dtype = {1: np.uint8, 2: np.uint16, 4: np.uint32}[count_bytes]
        self._ctx = DiffContext(nb_controls, nb_cases, int(sum(int(t) for t in total_controls)),
                                int(sum(int(t) for t in total_cases)), kmer_size=kmer_size,
                                count_bytes=count_bytes, lf_size=preload, device=device,
                                max_survivors=1024)

    def configure(self, config: str) -> None:  # IModel::configure (model.hpp:121)
        pass

    def process(self, controls, cases):
        c = np.concatenate([np.asarray(controls, dtype=self._dtype), np.asarray(cases, dtype=self._dtype)])
        assert c.size == self.nb_controls + self.nb_cases
        row = np.zeros(self._ctx.row_bytes, dtype=np.uint8)
        row[8 * self._ctx.kmer_words:] = c.view(np.uint8)
        p, sign, mc, kk = self._ctx.model_rows(row)
        return float(p[0]), Significance(int(sign[0])), float(mc[0]), float(kk[0])

    def process_rows(self, payload):
        return self._ctx.model_rows(payload)


def kmer_to_string(words, k: int) -> str:
    w = np.ascontiguousarray(words, dtype=np.uint64).ravel()
    buf = C.create_string_buffer(k + 1)
    _lib.load().kmd_kmer_to_string(w.ctypes.data, w.size, k, buf)
    return buf.value.decode()


def format_fasta(index: int, record: np.void, kmer_size: int) -> str:
    """One FASTA record as the reference's writer prints it (aggregator.hpp:55-62, 60-col wrap)."""
    rec = np.array([record], dtype=SURVIVOR_DTYPE)
    buf = C.create_string_buffer(512 + 2 * kmer_size)
    n = _lib.load().kmd_format_fasta(index, rec.ctypes.data, (kmer_size + 31) // 32, kmer_size, buf, len(buf))
    assert n < len(buf)
    return buf.value.decode()
