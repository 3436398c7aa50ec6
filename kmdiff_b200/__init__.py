"""kmdiff_b200 — B200-native implementation of the `kmdiff diff` hot path (tlemane/kmdiff).

The package holds only what that path needs: csrc/ (CUDA kernels + the C-ABI of
include/kmdiff_b200.h), host/ (C++ host program mirroring `kmdiff diff`), and thin Python
mirrors of the reference's operator interface used by the tests, bench.py and the multi-GPU
driver.  Importing the package does not load the CUDA library; the first compute call does, and
fails loudly if it is missing.
"""
from .api import (CorrectionType, DiffContext, KmdError, LogFactorialTable, PoissonLikelihood,  # noqa: F401
                  Significance, basic_threshold, benjamini, bonferroni, correction_from_str,
                  first_reject, format_fasta, holm, kmer_to_string, make_corrector, sampler_rows, sidak,
                  smartpca_pcs, sym_eig)
from .kmtricks_io import SynthSpec  # noqa: F401

__version__ = "0.1.0"
