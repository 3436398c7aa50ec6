#!/usr/bin/env python
"""`kmdiff-b200 diff --pop-correction` (GPU) against the reference binary's own run (oracle/_ref/kmdiff + smartpca +
evec2pca.perl) on the same run directory: .geno bytes, pcs.evec up to eigenvector sign, LRT / GLM p-values of the
spill files, FASTA k-mer sets.  Needs a GPU.  Prints one JSON line; tests/test_gpu_popstrat.py asserts on it
(first result: profiles/r01u/popstrat_vs_reference.json — p-values within 1e-12 relative, everything else identical)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kmdiff_b200 as K  # noqa: E402
import oracle  # noqa: E402
from kmdiff_b200 import build as kb  # noqa: E402
from kmdiff_b200 import kmtricks_io as kio  # noqa: E402
from test_smartpca_reference import _structured_cohort  # noqa: E402


def spill(cli, path, n):
    rec = 8 + 28 + 2 + 8 * n
    raw = subprocess.run([cli, "dev-cat", path, "0", "lz4"], capture_output=True).stdout
    a = np.frombuffer(raw, dtype=np.uint8).reshape(-1, rec)
    return np.ascontiguousarray(a[:, :8]).view("<u8").ravel(), np.ascontiguousarray(a[:, 8:16]).view("<f8").ravel()


def fasta_kmers(path):
    return sorted(l.strip() for l in open(path) if not l.startswith(">"))


def main():
    nc, nk, npc, m, seed, kmer_pca = 30, 30, 4, 40_000, 31, 0.05
    n = nc + nk
    counts = _structured_cohort(nc, nk, m, seed=35)
    rows = np.zeros((m, 8 + 4 * n), dtype=np.uint8)
    rows[:, :8] = (np.arange(m, dtype=np.uint64) * np.uint64(977) + np.uint64(5)).view(np.uint8).reshape(m, 8)
    rows[:, 8:] = counts.view(np.uint8).reshape(m, 4 * n)
    totals = counts.sum(0).astype(np.uint64) * np.uint64(1000)
    spec = K.SynthSpec(nc, nk, kmer_size=31, count_bytes=4, total_rows=m, seed=1)
    cli = kb.build_cli()
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        d = os.path.join(tmp, "run")
        kio.write_run_dir(d, spec, m, 1, payload=rows.ravel(), totals=totals)
        outs = {}
        for name, binary in (("ref", oracle.REF_BIN), ("new", cli)):
            opt = os.path.join(d, "kmdiff-count.opt")
            if os.path.exists(opt):
                os.remove(opt)
            outs[name] = os.path.join(tmp, name)
            r = subprocess.run([binary, "diff", "--km-run", d, "-1", str(nc), "-2", str(nk), "-o", outs[name], "--pop-correction",
                                "--kmer-pca", str(kmer_pca), "--random-seed", str(seed), "--n-pc", str(npc), "--keep-tmp", "-s", "0.5",
                                "-u", "100000", "-t", "1"], capture_output=True, text=True)
            assert "[error]" not in r.stderr, (name, r.stderr[-600:])
        pr, pn = (os.path.join(outs[x], "popstrat") for x in ("ref", "new"))
        res["geno_identical"] = open(os.path.join(pr, "gwas_eigenstratX.geno"), "rb").read() == open(os.path.join(pn, "gwas_eigenstratX.geno"), "rb").read()
        Zr, Zn = np.loadtxt(os.path.join(pr, "pcs.evec")), np.loadtxt(os.path.join(pn, "pcs.evec"))
        res["pcs_equal_up_to_sign"] = bool(all(np.array_equal(Zr[:, j], Zn[:, j]) or np.array_equal(Zr[:, j], -Zn[:, j]) for j in range(10)))
        for suffix in ("_uncorrected", "_popstrat_uncorrected"):
            kr, p_r = spill(cli, os.path.join(outs["ref"], "partitions", "p0" + suffix), n)
            kn, p_n = spill(cli, os.path.join(outs["new"], "partitions", "p0" + suffix), n)
            res["kmers_identical" + suffix] = bool(np.array_equal(kr, kn))
            nz = p_r > 0
            res["max_rel_p" + suffix] = float((np.abs(p_n[nz] - p_r[nz]) / p_r[nz]).max()) if len(kr) == len(kn) and nz.any() else None
        for f in ("control_kmers.fasta", "case_kmers.fasta"):
            res["fasta_kmers_identical_" + f] = fasta_kmers(os.path.join(outs["ref"], f)) == fasta_kmers(os.path.join(outs["new"], f))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
