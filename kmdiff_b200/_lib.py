"""ctypes binding of include/kmdiff_b200.h.  Fails loudly when the CUDA library is missing:
there is no Python or CPU fallback for any compute entry point."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libkmdiff_b200.so")

KMD_ABI_VERSION = 1
OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_OVERFLOW, ERR_STATE, ERR_NO_DEVICE = range(7)
SIGN_CONTROL, SIGN_CASE, SIGN_NO = 0, 1, 2
CORR_NOTHING, CORR_BONFERRONI, CORR_BENJAMINI, CORR_SIDAK, CORR_HOLM = range(5)

# every symbol include/kmdiff_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "kmd_abi_version", "kmd_device_count", "kmd_create", "kmd_destroy", "kmd_last_error",
    "kmd_set_stream", "kmd_sync", "kmd_reset", "kmd_scan_device", "kmd_scan_host",
    "kmd_counters_get", "kmd_survivors", "kmd_push_survivors", "kmd_push_survivors2", "kmd_launch_count", "kmd_scan_kernel_ms", "kmd_model_rows_device", "kmd_model_rows_host",
    "kmd_correct", "kmd_significant", "kmd_sorted_pvalues", "kmd_keep_below", "kmd_first_reject",
    "kmd_lf_table_fill", "kmd_chisqc", "kmd_corrector_apply", "kmd_kmer_to_string",
    "kmd_format_fasta", "kmd_host_alloc", "kmd_host_free", "kmd_synth_fill_host",
    "kmd_synth_fill_device", "kmd_synth_totals",
    "kmd_popstrat_enable", "kmd_popstrat_set_draw_index", "kmd_popstrat_sampled", "kmd_popstrat_geno", "kmd_popstrat_gram",
    "kmd_popstrat_gram_ms", "kmd_popstrat_gram_partial", "kmd_popstrat_gram_combine", "kmd_popstrat_gram_local", "kmd_popstrat_init", "kmd_popstrat_apply", "kmd_popstrat_fit",
    "kmd_survivor_counts", "kmd_smartpca_normalize", "kmd_sym_eig", "kmd_smartpca_pcs",
    "kmd_sampler_rows", "kmd_merge_scan_host", "kmd_merge_scan_lz4_host", "kmd_merge_prefetch", "kmd_lz4_decode_host", "kmd_merge_rows_host", "kmd_merge_ms", "kmd_merge_scan_device", "kmd_merge_info", "kmd_survivor_pvalues_device", "kmd_first_reject_device",
    "kmd_scan_host_async", "kmd_host_wait", "kmd_bind_host_thread", "kmd_warmup", "kmd_get_stream", "kmd_exchange_export", "kmd_exchange_correct", "kmd_exchange_result", "kmd_exchange_local",
]

EXCHANGE_HEADER = 8
FLAG_KEEP_COUNTS = 1


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("nb_controls", C.c_uint32),
                ("nb_cases", C.c_uint32), ("kmer_size", C.c_uint32), ("kmer_words", C.c_uint32),
                ("count_bytes", C.c_uint32), ("flags", C.c_uint32), ("sum_controls", C.c_uint64),
                ("sum_cases", C.c_uint64), ("threshold", C.c_double),
                ("lf_table", C.POINTER(C.c_double)), ("lf_size", C.c_uint64),
                ("max_survivors", C.c_uint64)]


class Counters(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("candidates", C.c_uint64), ("survivors", C.c_uint64),
                ("n_control", C.c_uint64), ("n_case", C.c_uint64)]


class ExchangeDecision(C.Structure):
    _fields_ = [("total_rows", C.c_uint64), ("total_survivors", C.c_uint64), ("max_rank_survivors", C.c_uint64),
                ("n_accepted", C.c_uint64), ("p_cut", C.c_double), ("kept_control", C.c_uint64),
                ("kept_case", C.c_uint64), ("overflow", C.c_uint32), ("reserved", C.c_uint32)]


class PopstratConfig(C.Structure):
    _fields_ = [("kmer_pca", C.c_double), ("seed", C.c_uint64), ("is_diploid", C.c_uint32),
                ("n_pc", C.c_uint32), ("max_sampled", C.c_uint64)]


class Synth(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("total_rows", C.c_uint64), ("nb_controls", C.c_uint32),
                ("nb_cases", C.c_uint32), ("kmer_size", C.c_uint32), ("kmer_words", C.c_uint32),
                ("count_bytes", C.c_uint32), ("enriched_ppm", C.c_uint32)]


SURVIVOR_DTYPE = np.dtype([("kmer", "<u8", (2,)), ("p_value", "<f8"), ("sign", "<i4"),
                           ("partition", "<u4"), ("mean_control", "<f8"), ("mean_case", "<f8"),
                           ("row", "<u8")])
assert SURVIVOR_DTYPE.itemsize == 56


class LibraryMissing(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen libkmdiff_b200.so (built in-tree by kmdiff_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -m kmdiff_b200.build` "
            f"(nvcc, sm_100a).  kmdiff_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32, d = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_double
    pu64 = C.POINTER(u64)
    L.kmd_abi_version.restype = C.c_int
    L.kmd_device_count.restype = C.c_int
    L.kmd_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.kmd_destroy.argtypes = [vp]
    L.kmd_destroy.restype = None
    L.kmd_last_error.argtypes = [vp]
    L.kmd_last_error.restype = C.c_char_p
    L.kmd_set_stream.argtypes = [vp, vp]
    L.kmd_sync.argtypes = [vp]
    L.kmd_reset.argtypes = [vp]
    L.kmd_scan_device.argtypes = [vp, vp, u64, u64, u32]
    L.kmd_scan_host.argtypes = [vp, vp, u64, u64, u32]
    L.kmd_merge_scan_host.argtypes = [vp, vp, vp, vp, u64, u32, pu64]
    L.kmd_merge_rows_host.argtypes = [vp, vp, vp, vp, vp, u64, pu64]
    L.kmd_merge_scan_device.argtypes = [vp, vp, vp, vp, u64, u32, pu64]
    L.kmd_merge_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), pu64, pu64]
    L.kmd_merge_scan_lz4_host.argtypes = [vp, vp, vp, vp, vp, u64, u32, u64, pu64, pu64]
    L.kmd_merge_prefetch.argtypes = [vp, vp, vp, C.c_int]
    L.kmd_lz4_decode_host.argtypes = [vp, vp, vp, vp, u64, pu64, C.POINTER(C.c_float)]
    L.kmd_merge_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(u32), pu64]
    L.kmd_counters_get.argtypes = [vp, C.POINTER(Counters)]
    L.kmd_survivors.argtypes = [vp, vp, u64, pu64]
    L.kmd_push_survivors.argtypes = [vp, vp, u64, u64]
    L.kmd_push_survivors2.argtypes = [vp, vp, vp, u64, u64]
    L.kmd_launch_count.argtypes = [vp]
    L.kmd_launch_count.restype = u64
    L.kmd_scan_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.kmd_model_rows_device.argtypes = [vp, vp, u64, vp, vp, vp, vp]
    L.kmd_model_rows_host.argtypes = [vp, vp, u64, vp, vp, vp, vp]
    L.kmd_correct.argtypes = [vp, C.c_int, d, u64, pu64, pu64]
    L.kmd_significant.argtypes = [vp, vp, u64, pu64]
    L.kmd_sorted_pvalues.argtypes = [vp, vp, u64, pu64]
    L.kmd_keep_below.argtypes = [vp, d, pu64, pu64]
    L.kmd_survivor_pvalues_device.argtypes = [vp, vp, u64, pu64]
    L.kmd_first_reject_device.argtypes = [vp, vp, u64, C.c_int, d, u64, pu64, C.POINTER(d)]
    L.kmd_scan_host_async.argtypes = [vp, vp, u64, u64, u32, pu64]
    L.kmd_host_wait.argtypes = [vp, u64]
    L.kmd_get_stream.argtypes = [vp]
    L.kmd_get_stream.restype = vp
    L.kmd_exchange_export.argtypes = [vp, vp, u64]
    L.kmd_exchange_correct.argtypes = [vp, vp, u32, u64, C.c_int, d]
    L.kmd_exchange_result.argtypes = [vp, C.POINTER(ExchangeDecision)]
    L.kmd_exchange_local.argtypes = [vp, u32, C.c_int, d, C.POINTER(ExchangeDecision)]
    L.kmd_first_reject.argtypes = [vp, u64, C.c_int, d, u64, C.POINTER(d)]
    L.kmd_first_reject.restype = u64
    L.kmd_lf_table_fill.argtypes = [vp, u64]
    L.kmd_lf_table_fill.restype = None
    L.kmd_chisqc.argtypes = [d, d]
    L.kmd_chisqc.restype = d
    L.kmd_corrector_apply.argtypes = [C.c_int, d, u64, u64, d]
    L.kmd_kmer_to_string.argtypes = [vp, u32, u32, C.c_char_p]
    L.kmd_kmer_to_string.restype = None
    L.kmd_format_fasta.argtypes = [u64, vp, u32, u32, C.c_char_p, C.c_size_t]
    L.kmd_format_fasta.restype = C.c_size_t
    L.kmd_bind_host_thread.argtypes = [C.c_int]
    L.kmd_warmup.argtypes = [C.c_int]
    L.kmd_host_alloc.argtypes = [C.c_size_t]
    L.kmd_host_alloc.restype = vp
    L.kmd_host_free.argtypes = [vp]
    L.kmd_host_free.restype = None
    L.kmd_synth_fill_host.argtypes = [C.POINTER(Synth), vp, u64, u64, C.c_int]
    L.kmd_synth_fill_device.argtypes = [C.POINTER(Synth), C.c_int, vp, u64, u64, vp]
    L.kmd_synth_totals.argtypes = [C.POINTER(Synth), vp]
    L.kmd_synth_totals.restype = None
    L.kmd_popstrat_enable.argtypes = [vp, C.POINTER(PopstratConfig)]
    L.kmd_popstrat_set_draw_index.argtypes = [vp, u64]
    L.kmd_popstrat_sampled.argtypes = [vp, pu64]
    L.kmd_popstrat_geno.argtypes = [vp, vp, vp, u64, pu64]
    L.kmd_popstrat_gram.argtypes = [vp, vp, C.c_int, pu64]
    L.kmd_popstrat_gram_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.kmd_popstrat_gram_partial.argtypes = [vp, vp, vp]
    L.kmd_popstrat_gram_local.argtypes = [vp, u32, vp, C.c_int, pu64]
    L.kmd_popstrat_gram_combine.argtypes = [vp, vp, vp, vp, C.c_int, pu64]
    L.kmd_popstrat_init.argtypes = [vp, vp, vp, vp]
    L.kmd_popstrat_apply.argtypes = [vp]
    L.kmd_popstrat_fit.argtypes = [vp, vp, u64, vp, vp, vp, vp]
    L.kmd_survivor_counts.argtypes = [vp, vp, u64, pu64]
    L.kmd_smartpca_normalize.argtypes = [vp, u32]
    L.kmd_smartpca_normalize.restype = None
    L.kmd_sym_eig.argtypes = [vp, u32, vp, vp]
    L.kmd_smartpca_pcs.argtypes = [vp, u32, u32, vp, vp]
    L.kmd_sampler_rows.argtypes = [u64, d, u64, u64, vp]
    L.kmd_sampler_rows.restype = u64
    _lib = L
    return L
