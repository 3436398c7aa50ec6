"""BASELINE.json configs[0]: the reference's own example (/root/reference/examples/run.sh) —
`kmdiff count --kmer-size 31 --hard-min 2` on examples/data (20 FASTA files), then
`kmdiff diff -1 10 -2 10 -s 0.01` on the run directory kmtricks left: per-sample `.kmer.lz4` count files in 4
partitions, NO matrices (the km::KmerMerger input path).  The run directory in tests/golden/config1_run.tar.xz was
written by the reference's kmtricks (tests/golden/make_config1_golden.sh); the goldens below are what the
unmodified reference `kmdiff diff` prints and writes for it.
"""
import hashlib
import os
import re
import subprocess
import tarfile

import numpy as np
import pytest

import oracle
from kmdiff_b200 import build as kb
from kmdiff_b200 import kmtricks_io as kio

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "config1_run.tar.xz")

ROWS, PREFILTER = 139876, 59014
BEFORE = (23985, 35029)          # "Before correction: 23985 (control), 35029 (case)."
AFTER = (23761, 34837)           # "Significant k-mers: 23761 (control), 34837 (case)."
MD5 = {"case": "85c677ff748c477b2571fab37a0d09a0", "control": "9437e5a6abb836fd38a777c805aa6e5d"}
FIRST = {"case": (">0_pval=1.11209e-10_control=0_case=30", "AAAAAAAATGAATGGAGAGCATATTATATGT"),
         "control": (">0_pval=1.0429e-17_control=52_case=0", "AAAAAAATCCCCCCGAGGGGGGGGATCTCAA")}
SIGNIFICANCE, CUTOFF = 0.01, 100000


@pytest.fixture(scope="module")
def run_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("config1")
    with tarfile.open(FIXTURE) as t:
        t.extractall(d, filter="data")
    os.makedirs(d / "matrices", exist_ok=True)
    assert os.listdir(d / "matrices") == []
    return str(d)


def md5_sorted(seqs):
    return hashlib.md5("".join(s + "\n" for s in sorted(seqs)).encode()).hexdigest()


def test_oracle_reproduces_the_reference_goldens(run_dir):
    """the CPU restatement (KmerMerger + PoissonLikelihood + bonferroni + FASTA naming) on the real kmtricks output"""
    cfg = kio.get_kmtricks_config(run_dir)
    assert (cfg.kmer_size, cfg.abundance_min, cfg.nb_partitions) == (31, 2, 4)
    tc, tk = kio.get_total_kmer(run_dir, 10, 10, cfg.abundance_min)
    model = oracle.Model(10, 10, np.array(tc + tk, dtype=np.uint64))
    total, n_ctl, n_case, surv = 0, 0, 0, []
    for p in range(cfg.nb_partitions):
        streams = [kio.read_kmer_file(f)[1] for f in kio.get_partition_paths(run_dir, p)]
        rows = oracle.kmer_merge(streams, 1, 4, cfg.abundance_min)
        total += len(rows)
        s, c, k = model.diff(rows.ravel(), 1, 4, SIGNIFICANCE / CUTOFF)
        n_ctl, n_case = n_ctl + c, n_case + k
        surv.append(s)
    surv = np.concatenate(surv)
    assert total == ROWS and len(surv) == PREFILTER and (n_ctl, n_case) == BEFORE
    kept = surv[oracle.aggregate(surv["p"], oracle.BONFERRONI, SIGNIFICANCE, total)]
    ctl = kept[kept["sign"] == 0]
    case = kept[kept["sign"] != 0]
    assert (len(ctl), len(case)) == AFTER
    assert md5_sorted(oracle.kmer_to_string(r["kmer"][:1], 31) for r in ctl) == MD5["control"]
    assert md5_sorted(oracle.kmer_to_string(r["kmer"][:1], 31) for r in case) == MD5["case"]
    # the reference's first records are among ours, formatted identically
    for name, recs in (("control", ctl), ("case", case)):
        head, seq = FIRST[name]
        hit = [r for r in recs if oracle.kmer_to_string(r["kmer"][:1], 31) == seq]
        assert len(hit) == 1
        rec = oracle.fasta_record(0, hit[0]["p"], hit[0]["mean_control"], hit[0]["mean_case"], seq)
        assert rec == f"{head}\n{seq}\n"


def read_seqs(path):
    heads, seqs = [], []
    with open(path) as f:
        for line in f:
            (heads if line.startswith(">") else seqs).append(line.rstrip("\n"))
    return heads, seqs


@pytest.mark.gpu
@pytest.mark.parametrize("threads,extra", [(1, []), (8, []), (4, ["--device-lz4"])])
def test_cli_reproduces_the_reference_goldens(run_dir, tmp_path, threads, extra):
    """`kmdiff-b200 diff --km-run kc_dir -1 10 -2 10 --output-dir out -s 0.01` (examples/run.sh:5); the `.kmer.lz4` files
    decoded by the host threads (default for 20 samples) and on the device (--device-lz4)"""
    cli = kb.build_cli()
    out = str(tmp_path / "out")
    r = subprocess.run([cli, "diff", "--km-run", run_dir, "-1", "10", "-2", "10", "--output-dir", out, "-s", "0.01",
                        "-t", str(threads), "-v", "debug"] + extra, capture_output=True, text=True, timeout=600)
    log = r.stdout + r.stderr
    assert r.returncode == 0 and "[error]" not in log, log
    assert f"{PREFILTER}/{ROWS} significant k-mers" in log, log
    assert ("records decoded on the device" in log) == bool(extra)
    assert f"Before correction: {BEFORE[0]} (control), {BEFORE[1]} (case)." in log
    assert f"Significant k-mers: {AFTER[0]} (control), {AFTER[1]} (case)." in log
    for name, want in (("control", AFTER[0]), ("case", AFTER[1])):
        heads, seqs = read_seqs(os.path.join(out, f"{name}_kmers.fasta"))
        assert len(seqs) == want and len(heads) == want
        assert md5_sorted(seqs) == MD5[name]
        head, seq = FIRST[name]
        i = seqs.index(seq)
        assert re.sub(r"^>\d+_", ">0_", heads[i]) == head
