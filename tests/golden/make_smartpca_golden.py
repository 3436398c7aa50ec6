#!/usr/bin/env python
"""tests/golden/smartpca_vectors.json: outputs of the UNMODIFIED HAWK smartpca + evec2pca.perl (oracle/_ref/smartpca,
built from the reference's vendored sources by `make -C oracle smartpca`) driven exactly as kmdiff drives them
(src/popstrat.cpp:44-61 parfile, :97-135 run_eigenstrat_smartpca), on seeded presence matrices.
Build container only:  python tests/golden/make_smartpca_golden.py

Per case: the .geno matrix (bit-packed, hex), whether smartpca ran diploid (`-p`) or haploid (`-V -p`), the
eigenvalues of gwas_eigenstrat.eval and the text of pcs.evec (what pop_strat_corrector::load_Z reads)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

PARFILE = ("genotypename: gwas_eigenstratX.geno\nsnpname: gwas_eigenstratX.snp\nindivname: gwas_eigenstratX.ind\n"
           "evecoutname: gwas_eigenstrat.evec\nevaloutname: gwas_eigenstrat.eval\nusenorm: YES\nnumoutlieriter: 0\n"
           "numoutevec: 10\n")


def geno_matrix(rng, m, n):
    f = rng.beta(0.5, 0.5, size=(m, 1))
    g = (rng.random((m, n)) < f).astype(np.uint8)
    g[0] = 0                                   # a k-mer nobody carries: smartpca drops it
    g[1] = 1                                   # carried by everybody
    g[:, :n // 3] |= (rng.random((m, 1)) < 0.15).astype(np.uint8)   # a sub-population
    return g


def run_smartpca(geno, nc, diploid):
    """-> (evals, pcs.evec text) from the real programs, in a scratch directory"""
    m, n = geno.shape
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "gwas_eigenstratX.geno"), "w") as f:       # popstrat.hpp:56-80
            for row in geno:
                f.write("".join("1\t" if x else "0\t" for x in row) + "\n")
        with open(os.path.join(d, "gwas_eigenstratX.snp"), "w") as f:
            for i in range(m):
                f.write(f"{i}\t1\t0.0\t0\n")
        with open(os.path.join(d, "gwas_eigenstratX.ind"), "w") as f:        # src/popstrat.cpp:63-84
            for i in range(n):
                f.write(f"S{i}\tU\t{'Control' if i < nc else 'Case'}\n")
        # HAWK's smartpca opens "gwas_eigenstratX.total" unconditionally (smartpca.c:589-610) and crashes without
        # it; the scaled totals it derives are not used (smartpca.c:1695-1697 are commented out)
        with open(os.path.join(d, "gwas_eigenstratX.total"), "w") as f:  # src/popstrat.cpp:86-95
            for i in range(n):
                f.write(f"{1000000 + 17 * i}\n")
        with open(os.path.join(d, "parfile.txt"), "w") as f:
            f.write(PARFILE)
        args = ["-p", "parfile.txt"] if diploid else ["-V", "-p", "parfile.txt"]
        r = subprocess.run([oracle.REF_SMARTPCA] + args, cwd=d, capture_output=True, text=True)
        assert os.path.exists(os.path.join(d, "gwas_eigenstrat.evec")), r.stdout[-800:] + r.stderr[-800:]
        subprocess.run([os.path.join(os.path.dirname(oracle.REF_SMARTPCA), "evec2pca.perl"), str(min(n, 10)), "gwas_eigenstrat.evec",
                        "gwas_eigenstratX.ind", "gwas_eigenstrat.pca"], cwd=d, check=True, capture_output=True)
        with open(os.path.join(d, "gwas_eigenstrat.pca")) as f:
            lines = f.read().split("\n")
        skip = int(lines[0])
        pcs_text = "\n".join(lines[1 + skip:])                                # src/popstrat.cpp:124-132
        evals = np.loadtxt(os.path.join(d, "gwas_eigenstrat.eval")).tolist()
    return evals, pcs_text


def main():
    rng = np.random.default_rng(20261018)
    cases = []
    for m, n, nc, diploid in ((3000, 36, 16, True), (1500, 80, 40, True), (800, 24, 12, False), (5000, 130, 60, True)):
        g = geno_matrix(rng, m, n)
        evals, pcs = run_smartpca(g, nc, diploid)
        cases.append({"m": m, "n": n, "nb_controls": nc, "diploid": diploid,
                      "geno_bits": np.packbits(g, axis=None).tobytes().hex(), "evals": evals, "pcs_evec": pcs})
    with open(os.path.join(HERE, "smartpca_vectors.json"), "w") as f:
        json.dump({"source": "HAWK EIG6.0.1 smartpca (version 13050) + evec2pca.perl from the reference's thirdparty/hawk, "
                             "LAPACK dspev from OpenBLAS; eigenvector signs are LAPACK's and carry no meaning", "cases": cases}, f)
    print([(c["m"], c["n"], len(c["evals"])) for c in cases])


if __name__ == "__main__":
    main()
