#!/usr/bin/env python
"""tests/golden/kff_vectors.json: {control,case}_kmers.kff files written by the UNMODIFIED reference binary
(oracle/_ref/kmdiff diff -f -t 1) next to the k-mers of the FASTA files of the same run without -f (same
order).  Build container only:  python tests/golden/make_kff_golden.py

oracle/_ref stamps kff-cpp-api's KFF_VERSION_MAJOR/MINOR as 0/0 (its CMake configure step is not run); the
reference's real build writes 1/0 (thirdparty/kff-cpp-api/CMakeLists.txt:5), which is what the fixtures store."""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import kmdiff_b200 as K  # noqa: E402
import oracle  # noqa: E402
from kmdiff_b200 import kmtricks_io as kio  # noqa: E402


def main():
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        for k in (31, 30, 21):
            spec = K.SynthSpec(10, 10, kmer_size=k, count_bytes=4, total_rows=20000, seed=99, enriched_ppm=3000)
            d = os.path.join(tmp, f"run{k}")
            kio.write_run_dir(d, spec, 20000, 3)
            outs = {}
            for tag, extra in (("fa", []), ("kff", ["-f"])):
                opt = os.path.join(d, "kmdiff-count.opt")
                if os.path.exists(opt):
                    os.remove(opt)
                outs[tag] = os.path.join(tmp, f"{tag}{k}")
                r = subprocess.run([oracle.REF_BIN, "diff", "--km-run", d, "-1", "10", "-2", "10", "-o", outs[tag], "-t", "1",
                                    "-s", "0.05", "-u", "100000"] + extra, capture_output=True, text=True)
                assert "[error]" not in r.stderr, r.stderr
            for name in ("control", "case"):
                seqs = [l.strip() for l in open(os.path.join(outs["fa"], f"{name}_kmers.fasta")) if not l.startswith(">")]
                raw = bytearray(open(os.path.join(outs["kff"], f"{name}_kmers.kff"), "rb").read())
                raw[3], raw[4] = 1, 0
                cases.append({"k": k, "file": f"{name}_kmers.kff", "kmers": seqs, "kff": bytes(raw).hex()})
    with open(os.path.join(HERE, "kff_vectors.json"), "w") as f:
        json.dump({"source": "oracle/_ref/kmdiff diff -f -t 1 (KffWriter, include/kmdiff/kff_utils.hpp:32-107)", "cases": cases}, f)
    print([(c["k"], c["file"], len(c["kmers"])) for c in cases])


if __name__ == "__main__":
    main()
