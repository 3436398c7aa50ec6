"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes bindings for the parity oracle:

* ``libkmd_oracle.so``      — our plain-C restatement of the reference's `kmdiff diff` arithmetic
                              (oracle/kmd_oracle.c, each function cites reference file:line);
* ``_ref/libkmdiff_ref.so`` — the UNMODIFIED reference (PoissonLikelihood, LogFactorialTable,
                              correctors, alglib chi-square) compiled from /root/reference behind
                              oracle/ref_shim.cpp.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs may import this package — as the checker, never as the thing measured or shipped.  The
product package ``kmdiff_b200`` never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libkmd_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libkmdiff_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "kmdiff")
REF_BIN16 = os.path.join(HERE, "_ref", "kmdiff_c16")
REF_SMARTPCA = os.path.join(HERE, "_ref", "smartpca")
REF_PLUGIN_OUR_HEADERS = os.path.join(HERE, "_ref", "libref_ex_model.so")     # reference plugin source x this repo's headers
OUR_PLUGIN_REF_HEADERS = os.path.join(HERE, "_ref", "libex_model_refhdr.so")   # this repo's plugin x the reference's headers
REFERENCE_ROOT = "/root/reference"

CONTROL, CASE, NO = 0, 1, 2
NOTHING, BONFERRONI, BENJAMINI, SIDAK, HOLM = 0, 1, 2, 3, 4
CORRECTIONS = {"disabled": NOTHING, "nothing": NOTHING, "bonferroni": BONFERRONI,
               "benjamini": BENJAMINI, "sidak": SIDAK, "holm": HOLM}


def build(ref: bool | None = None, refbin: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is present, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref is None:
        ref = os.path.isdir(REFERENCE_ROOT)
    if ref:
        targets = ["ref"] + (["refbin", "refbin16"] if refbin else [])
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE] + targets)
        if refbin:
            # HAWK's smartpca next to the reference binary: lets the reference's --pop-correction pipeline run here
            # (needs scipy's bundled OpenBLAS for LAPACK); the parity tests that use it skip when it is absent
            subprocess.call(["make", "-s", "-j8", "-C", HERE, "smartpca"])
            subprocess.call(["make", "-s", "-C", HERE, "plugins"])


class _Model(C.Structure):
    _fields_ = [("nb_controls", C.c_uint32), ("nb_cases", C.c_uint32),
                ("sum_controls", C.c_uint64), ("sum_cases", C.c_uint64),
                ("lf_table", C.POINTER(C.c_double)), ("lf_size", C.c_uint64)]


class _Corrector(C.Structure):
    _fields_ = [("type", C.c_int), ("threshold", C.c_double), ("total", C.c_uint64),
                ("rank", C.c_uint64)]


class _PopStrat(C.Structure):
    _fields_ = [("n", C.c_int), ("nb_controls", C.c_int), ("nb_cases", C.c_int), ("npc", C.c_int),
                ("null_p", C.c_int), ("alt_p", C.c_int), ("max_iter", C.c_int), ("epsilon", C.c_double),
                ("null_features", C.POINTER(C.c_double)), ("alt_features", C.POINTER(C.c_double)),
                ("Y", C.POINTER(C.c_double)), ("totals", C.POINTER(C.c_double)),
                ("null_model", C.POINTER(C.c_double))]


SURVIVOR_DTYPE = np.dtype([("kmer", "<u8", (2,)), ("p", "<f8"), ("sign", "<i4"), ("pad_", "<i4"),
                           ("mean_control", "<f8"), ("mean_case", "<f8"), ("row", "<u8")])

_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        d, u64, vp = C.c_double, C.c_uint64, C.c_void_p
        L.kmo_log_factorial.restype = d
        L.kmo_log_factorial.argtypes = [u64]
        L.kmo_lf_table_fill.argtypes = [vp, u64]
        L.kmo_lf_lookup.restype = d
        L.kmo_lf_lookup.argtypes = [vp, u64, u64]
        L.kmo_lngamma.restype = d
        L.kmo_lngamma.argtypes = [d, C.POINTER(d)]
        for f in (L.kmo_igam, L.kmo_igamc, L.kmo_chisqc):
            f.restype = d
            f.argtypes = [d, d]
        L.kmo_model_rows.argtypes = [C.POINTER(_Model), vp, u64, C.c_int, C.c_int, vp, vp, vp, vp]
        L.kmo_diff_rows.restype = u64
        L.kmo_diff_rows.argtypes = [C.POINTER(_Model), vp, u64, C.c_int, C.c_int, d, u64, vp, u64,
                                    C.POINTER(u64), C.POINTER(u64)]
        L.kmo_corrector_init.argtypes = [C.POINTER(_Corrector), C.c_int, d, u64]
        L.kmo_corrector_apply.restype = C.c_int
        L.kmo_corrector_apply.argtypes = [C.POINTER(_Corrector), d]
        L.kmo_aggregate.restype = u64
        L.kmo_aggregate.argtypes = [vp, u64, C.c_int, d, u64, vp]
        L.kmo_kmer_to_string.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
        # population stratification (kmd_popstrat_oracle.h)
        i32 = C.c_int
        L.kmo_sampler_rows.restype = u64
        L.kmo_sampler_rows.argtypes = [u64, d, u64, u64, vp]
        L.kmo_sigmoid.restype = d
        L.kmo_sigmoid.argtypes = [d]
        L.kmo_inverse.argtypes = [vp, i32, vp, C.POINTER(i32), C.POINTER(i32)]
        L.kmo_glm_irls.argtypes = [vp, vp, i32, i32, i32, vp, C.POINTER(i32), C.POINTER(i32),
                                   C.POINTER(d), C.POINTER(i32)]
        L.kmo_popstrat_init.restype = i32
        L.kmo_popstrat_init.argtypes = [C.POINTER(_PopStrat), i32, i32, vp, vp, i32]
        L.kmo_popstrat_free.argtypes = [C.POINTER(_PopStrat)]
        L.kmo_popstrat_apply.restype = d
        L.kmo_popstrat_apply.argtypes = [C.POINTER(_PopStrat), vp, vp, C.POINTER(i32), C.POINTER(d)]
        L.kmo_smartpca_xtx.restype = u64
        L.kmo_smartpca_xtx.argtypes = [vp, u64, i32, i32, vp]
        L.kmo_sym_eig.argtypes = [vp, i32, vp, vp]
        L.kmo_smartpca_pcs.argtypes = [vp, u64, i32, i32, vp, i32, vp]
        # per-sample count files (kmd_merge_oracle.h)
        L.kmo_kmer_merge.restype = u64
        L.kmo_kmer_merge.argtypes = [vp, vp, C.c_uint32, i32, i32, vp, C.c_uint32, C.c_uint32, vp, u64]
        _lib = L
    return _lib


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def ref() -> C.CDLL:
    """The unmodified reference behind ref_shim.cpp (raises if oracle/_ref was never built)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle ref` where "
                                    f"{REFERENCE_ROOT} exists")
        L = C.CDLL(REF_SO)
        d, u64, vp, sz = C.c_double, C.c_uint64, C.c_void_p, C.c_size_t
        L.ref_lf_table.argtypes = [sz, sz, vp, vp]
        L.ref_chisqc.restype = d
        L.ref_chisqc.argtypes = [d, d]
        L.ref_model_create.restype = vp
        L.ref_model_create.argtypes = [C.c_int, sz, sz, vp, sz]
        L.ref_model_destroy.argtypes = [C.c_int, vp]
        L.ref_model_rows.argtypes = [C.c_int, vp, vp, sz, sz, vp, vp, vp, vp, C.c_int]
        L.ref_corrector_create.restype = vp
        L.ref_corrector_create.argtypes = [C.c_int, d, sz]
        L.ref_corrector_apply.restype = C.c_int
        L.ref_corrector_apply.argtypes = [vp, d]
        L.ref_corrector_destroy.argtypes = [vp]
        if hasattr(L, "ref_glm_irls"):
            i32 = C.c_int
            L.ref_sampler_rows.restype = sz
            L.ref_sampler_rows.argtypes = [sz, d, sz, vp]
            L.ref_sigmoid.restype = d
            L.ref_sigmoid.argtypes = [d]
            L.ref_inverse.argtypes = [vp, i32, vp, C.POINTER(i32), C.POINTER(i32)]
            L.ref_glm_irls.argtypes = [vp, vp, i32, i32, i32, vp, C.POINTER(i32), C.POINTER(i32),
                                       C.POINTER(d), C.POINTER(i32)]
            L.ref_popstrat_create.restype = vp
            L.ref_popstrat_create.argtypes = [sz, sz, vp, sz, C.c_char_p, C.c_char_p]
            L.ref_popstrat_destroy.argtypes = [vp]
            L.ref_popstrat_state.argtypes = [vp, vp, vp, vp]
            L.ref_popstrat_apply.restype = d
            L.ref_popstrat_apply.argtypes = [vp, vp, sz]
        if hasattr(L, "ref_kmer_merge"):
            L.ref_kmer_merge.restype = sz
            L.ref_kmer_merge.argtypes = [vp, sz, C.c_uint32, C.c_int, vp, vp, sz]
        _ref = L
    return _ref


# ------------------------------------------------------------------------------------------------
# numpy-level helpers around the C restatement
# ------------------------------------------------------------------------------------------------

def lf_table(size: int) -> np.ndarray:
    t = np.empty(size, dtype=np.float64)
    lib().kmo_lf_table_fill(t.ctypes.data, size)
    return t


def lf_lookup(table: np.ndarray, i: int) -> float:
    return lib().kmo_lf_lookup(table.ctypes.data, table.size, int(i))


def chisqc(v: float, x: float) -> float:
    return lib().kmo_chisqc(v, x)


@dataclass
class Model:
    """PoissonLikelihood state (reference include/kmdiff/model.hpp:105-118)."""
    nb_controls: int
    nb_cases: int
    totals: np.ndarray          # per-sample totals, controls first (get_total_kmer)
    lf_size: int = 10000        # --log-factorial default (src/cli.cpp:354-357)

    def __post_init__(self):
        self.totals = np.ascontiguousarray(self.totals, dtype=np.uint64)
        assert self.totals.size == self.nb_controls + self.nb_cases
        self.sum_controls = int(self.totals[:self.nb_controls].sum())
        self.sum_cases = int(self.totals[self.nb_controls:].sum())
        self.table = lf_table(self.lf_size)
        self._c = _Model(self.nb_controls, self.nb_cases, self.sum_controls, self.sum_cases,
                         self.table.ctypes.data_as(C.POINTER(C.c_double)), self.lf_size)

    def row_bytes(self, kmer_words: int, count_bytes: int) -> int:
        return 8 * kmer_words + (self.nb_controls + self.nb_cases) * count_bytes

    def rows(self, payload: np.ndarray, kmer_words: int, count_bytes: int):
        """(p, sign, mean_control, mean_case) for every row of an AoS matrix payload."""
        payload = np.ascontiguousarray(payload, dtype=np.uint8).ravel()
        R = self.row_bytes(kmer_words, count_bytes)
        n = payload.size // R
        p = np.empty(n, np.float64)
        sign = np.empty(n, np.int32)
        mc = np.empty(n, np.float64)
        kk = np.empty(n, np.float64)
        lib().kmo_model_rows(C.byref(self._c), payload.ctypes.data, n, kmer_words, count_bytes,
                             p.ctypes.data, sign.ctypes.data, mc.ctypes.data, kk.ctypes.data)
        return p, sign, mc, kk

    def diff(self, payload: np.ndarray, kmer_words: int, count_bytes: int, thr: float,
             row_base: int = 0):
        """Survivors of diff_observer::process (p <= thr) + (n_control, n_case)."""
        payload = np.ascontiguousarray(payload, dtype=np.uint8).ravel()
        R = self.row_bytes(kmer_words, count_bytes)
        n = payload.size // R
        cap = max(1024, n // 16)
        while True:
            out = np.zeros(cap, dtype=SURVIVOR_DTYPE)
            nc, nk = C.c_uint64(), C.c_uint64()
            found = lib().kmo_diff_rows(C.byref(self._c), payload.ctypes.data, n, kmer_words,
                                        count_bytes, thr, row_base, out.ctypes.data, cap,
                                        C.byref(nc), C.byref(nk))
            if found <= cap:
                return out[:found], nc.value, nk.value
            cap = found


class Corrector:
    """Stateful corrector (reference src/corrector.cpp)."""

    def __init__(self, kind: int, threshold: float, total: int):
        self._c = _Corrector()
        lib().kmo_corrector_init(C.byref(self._c), kind, threshold, total)

    def apply(self, p: float) -> bool:
        return bool(lib().kmo_corrector_apply(C.byref(self._c), p))


def aggregate(p: np.ndarray, kind: int, threshold: float, total: int) -> np.ndarray:
    """Indices of kept survivors in output order (reference aggregator.hpp run())."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    keep = np.empty(max(1, p.size), dtype=np.uint64)
    n = lib().kmo_aggregate(p.ctypes.data, p.size, kind, threshold, total, keep.ctypes.data)
    return keep[:n].astype(np.int64)


def kmer_merge(streams, kmer_words: int, count_bytes: int, ab_min=1, recurrence_min: int = 1,
               save_if: int = 0) -> np.ndarray:
    """km::KmerMerger over in-memory record streams (kmd_merge_oracle.c) -> dense rows [n, R] uint8."""
    n = len(streams)
    arrs = [np.ascontiguousarray(np.asarray(s).view(np.uint8)).ravel() for s in streams]
    rec = 8 * kmer_words + count_bytes
    ptrs = (C.c_void_p * n)(*[a.ctypes.data if a.size else None for a in arrs])
    cnt = (C.c_uint64 * n)(*[a.size // rec for a in arrs])
    ab = np.ascontiguousarray(np.broadcast_to(np.asarray(ab_min, dtype=np.uint32), (n,)))
    cap = int(sum(cnt))
    R = 8 * kmer_words + n * count_bytes
    out = np.zeros((max(cap, 1), R), dtype=np.uint8)
    kept = lib().kmo_kmer_merge(ptrs, cnt, n, kmer_words, count_bytes, ab.ctypes.data, recurrence_min, save_if,
                                out.ctypes.data, cap)
    return out[:kept]


def ref_kmer_merge(paths, kmer_size: int, count_bytes: int, ab_min=1, cap_rows: int = 1 << 20) -> np.ndarray:
    """The real km::KmerMerger (oracle/_ref) over .kmer[.lz4] files -> dense rows [n, R] uint8."""
    n = len(paths)
    arr = (C.c_char_p * n)(*[p.encode() for p in paths])
    ab = np.ascontiguousarray(np.broadcast_to(np.asarray(ab_min, dtype=np.uint32), (n,)))
    R = 8 * ((kmer_size + 31) // 32) + n * count_bytes
    out = np.zeros((cap_rows, R), dtype=np.uint8)
    kept = ref().ref_kmer_merge(arr, n, kmer_size, count_bytes, ab.ctypes.data, out.ctypes.data, cap_rows)
    assert kept <= cap_rows
    return out[:kept]


def kmer_to_string(words, k: int) -> str:
    w = np.ascontiguousarray(words, dtype=np.uint64).ravel()
    buf = C.create_string_buffer(k + 1)
    lib().kmo_kmer_to_string(w.ctypes.data, w.size, k, buf)
    return buf.value.decode()


# ------------------------------------------------------------------------------------------------
# numpy-level helpers around the real reference (oracle/_ref)
# ------------------------------------------------------------------------------------------------

def ref_lf(size: int, idx) -> np.ndarray:
    idx = np.ascontiguousarray(idx, dtype=np.uint64)
    out = np.empty(idx.size, np.float64)
    ref().ref_lf_table(size, idx.size, idx.ctypes.data, out.ctypes.data)
    return out


def ref_chisqc(v: float, x: float) -> float:
    return ref().ref_chisqc(v, x)


def ref_model_rows(nb_controls: int, nb_cases: int, totals, lf_size: int, payload: np.ndarray,
                   kmer_words: int, count_bytes: int, threads: int = 1):
    """kmdiff::PoissonLikelihood<MAX_C>::process on every row, through the real reference."""
    totals = np.ascontiguousarray(totals, dtype=np.uint64)
    payload = np.ascontiguousarray(payload, dtype=np.uint8).ravel()
    R = 8 * kmer_words + (nb_controls + nb_cases) * count_bytes
    n = payload.size // R
    h = ref().ref_model_create(count_bytes, nb_controls, nb_cases, totals.ctypes.data, lf_size)
    try:
        p = np.empty(n, np.float64)
        sign = np.empty(n, np.int32)
        mc = np.empty(n, np.float64)
        kk = np.empty(n, np.float64)
        ref().ref_model_rows(count_bytes, h, payload.ctypes.data, n, 8 * kmer_words,
                             p.ctypes.data, sign.ctypes.data, mc.ctypes.data, kk.ctypes.data,
                             threads)
    finally:
        ref().ref_model_destroy(count_bytes, h)
    return p, sign, mc, kk


class RefCorrector:
    def __init__(self, kind: int, threshold: float, total: int):
        self._h = ref().ref_corrector_create(kind, threshold, total)

    def apply(self, p: float) -> bool:
        return bool(ref().ref_corrector_apply(self._h, p))

    def __del__(self):
        try:
            ref().ref_corrector_destroy(self._h)
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# Population stratification (kmd_popstrat_oracle.c; reference popstrat.hpp / linear_model.cpp / smartpca)
# ------------------------------------------------------------------------------------------------

def sampler_rows(seed: int, kmer_pca: float, n: int, first_row: int = 0) -> np.ndarray:
    """Sampler::sample() decisions for rows first_row..first_row+n-1 (popstrat.hpp:107-146)."""
    out = np.zeros(n, dtype=np.uint8)
    lib().kmo_sampler_rows(seed, kmer_pca, first_row, n, out.ctypes.data)
    return out.astype(bool)


def glm_irls(x: np.ndarray, y: np.ndarray, max_iters: int = 100, impl: str = "oracle"):
    """glm_irls (linear_model.cpp:297-410) -> (weights, singular, nan, error, iterations)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n, p = x.shape
    w = np.empty(p, np.float64)
    s, q, it = C.c_int(), C.c_int(), C.c_int()
    e = C.c_double()
    f = lib().kmo_glm_irls if impl == "oracle" else ref().ref_glm_irls
    f(x.ctypes.data, y.ctypes.data, n, p, max_iters, w.ctypes.data, C.byref(s), C.byref(q), C.byref(e),
      C.byref(it))
    return w, bool(s.value), bool(q.value), e.value, it.value


def inverse(a: np.ndarray, impl: str = "oracle"):
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    inv = np.empty((n, n), np.float64)
    s, q = C.c_int(), C.c_int()
    f = lib().kmo_inverse if impl == "oracle" else ref().ref_inverse
    f(a.ctypes.data, n, inv.ctypes.data, C.byref(s), C.byref(q))
    return inv, bool(s.value), bool(q.value)


class PopStrat:
    """pop_strat_corrector (popstrat.hpp:148-372): features, standardize, null model, apply."""

    def __init__(self, nb_controls: int, nb_cases: int, totals, Z: np.ndarray, npc: int):
        self.totals = np.ascontiguousarray(totals, dtype=np.uint64)
        Z10 = np.zeros((nb_controls + nb_cases, 10), np.float64)
        Z = np.asarray(Z, dtype=np.float64)
        Z10[:, :Z.shape[1]] = Z
        self.Z = Z10
        self._c = _PopStrat()
        rc = lib().kmo_popstrat_init(C.byref(self._c), nb_controls, nb_cases, self.totals.ctypes.data,
                                     Z10.ctypes.data, npc)
        if rc:
            raise ValueError("kmo_popstrat_init failed")
        n, pc, pa = self._c.n, self._c.null_p, self._c.alt_p
        self.n, self.null_p, self.alt_p = n, pc, pa
        self.null_features = np.ctypeslib.as_array(self._c.null_features, shape=(n, pc)).copy()
        self.alt_features = np.ctypeslib.as_array(self._c.alt_features, shape=(n, pa)).copy()
        self.null_model = np.ctypeslib.as_array(self._c.null_model, shape=(pc,)).copy()
        self.Y = np.ctypeslib.as_array(self._c.Y, shape=(n,)).copy()

    def apply(self, counts, full: bool = False):
        counts = np.ascontiguousarray(counts, dtype=np.float64)
        model = np.empty(self.alt_p, np.float64)
        it = C.c_int()
        llr = C.c_double()
        p = lib().kmo_popstrat_apply(C.byref(self._c), counts.ctypes.data, model.ctypes.data,
                                     C.byref(it), C.byref(llr))
        return (p, model, it.value, llr.value) if full else p

    def __del__(self):
        try:
            lib().kmo_popstrat_free(C.byref(self._c))
        except Exception:
            pass


class RefPopStrat:
    """The unmodified pop_strat_corrector behind ref_shim.cpp, fed through the files do_pop reads."""

    def __init__(self, nb_controls: int, nb_cases: int, totals, Z: np.ndarray, npc: int, workdir: str):
        totals = np.ascontiguousarray(totals, dtype=np.uint64)
        n = nb_controls + nb_cases
        Z10 = np.zeros((n, 10), np.float64)
        Z = np.asarray(Z, dtype=np.float64)
        Z10[:, :Z.shape[1]] = Z
        pcs = os.path.join(workdir, "pcs.evec")
        ind = os.path.join(workdir, "gwas_eigenstratX.ind")
        with open(pcs, "w") as f:
            for row in Z10:     # evec2pca.perl layout: " %.04f" per component
                f.write("".join((" " if v <= 0 else "  ") + "%.04f" % v for v in row) + "\n")
        with open(ind, "w") as f:
            for i in range(n):
                f.write(f"S{i}\tU\t{'Control' if i < nb_controls else 'Case'}\n")
        self.n, self.null_p, self.alt_p = n, npc + 2, npc + 3
        self._h = ref().ref_popstrat_create(nb_controls, nb_cases, totals.ctypes.data, npc,
                                            pcs.encode(), ind.encode())
        self.null_features = np.empty((n, self.null_p), np.float64)
        self.alt_features = np.empty((n, self.alt_p), np.float64)
        self.null_model = np.empty(self.null_p, np.float64)
        ref().ref_popstrat_state(self._h, self.null_features.ctypes.data, self.alt_features.ctypes.data,
                                 self.null_model.ctypes.data)

    def apply(self, counts) -> float:
        counts = np.ascontiguousarray(counts, dtype=np.float64)
        return ref().ref_popstrat_apply(self._h, counts.ctypes.data, counts.size)

    def __del__(self):
        try:
            ref().ref_popstrat_destroy(self._h)
        except Exception:
            pass


def smartpca_xtx(geno: np.ndarray, is_diploid: bool = True):
    """(XTX normalised like smartpca.c:1011-1023, SNPs used).  geno: m x n of 0/1."""
    geno = np.ascontiguousarray(geno, dtype=np.uint8)
    m, n = geno.shape
    xtx = np.empty((n, n), np.float64)
    used = lib().kmo_smartpca_xtx(geno.ctypes.data, m, n, int(is_diploid), xtx.ctypes.data)
    return xtx, used


def sym_eig(a: np.ndarray):
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    evals = np.empty(n, np.float64)
    evecs = np.empty((n, n), np.float64)
    lib().kmo_sym_eig(a.ctypes.data, n, evals.ctypes.data, evecs.ctypes.data)
    return evals, evecs


def smartpca_pcs(geno: np.ndarray, evecs: np.ndarray, numeigs: int = 10, is_diploid: bool = True):
    geno = np.ascontiguousarray(geno, dtype=np.uint8)
    evecs = np.ascontiguousarray(evecs, dtype=np.float64)
    m, n = geno.shape
    pcs = np.empty((n, numeigs), np.float64)
    lib().kmo_smartpca_pcs(geno.ctypes.data, m, n, int(is_diploid), evecs.ctypes.data, numeigs,
                           pcs.ctypes.data)
    return pcs


# ------------------------------------------------------------------------------------------------
# Output formatting (reference include/kmdiff/aggregator.hpp:55-62, fmt 8.1.1, kseq++ 60-col wrap)
# ------------------------------------------------------------------------------------------------

def fmt_g(x: float) -> str:
    """fmt `{:g}` == printf %g (6 significant digits, trailing zeros stripped)."""
    return "%g" % x


def fmt_shortest(x: float) -> str:
    """fmt `{}` for double: shortest round-trip digits, fixed for 1e-4 <= |x| < 1e16."""
    r = repr(float(x))
    if r.endswith(".0"):
        r = r[:-2]
    if "e" in r:
        mant, exp = r.split("e")
        if mant.endswith(".0"):
            mant = mant[:-2]
        sign = "-" if exp.startswith("-") else "+"
        digits = exp.lstrip("+-").rjust(2, "0")
        r = f"{mant}e{sign}{digits}"
    return r


def fasta_record(index: int, p: float, mean_control: float, mean_case: float, seq: str) -> str:
    head = f">{index}_pval={fmt_g(p)}_control={int(mean_control)}_case={fmt_shortest(mean_case)}"
    lines = [seq[i:i + 60] for i in range(0, len(seq), 60)]
    return head + "\n" + "\n".join(lines) + "\n"
