"""The C++ host program `kmdiff-b200 diff` (kmdiff_b200/host): kmtricks format code on the CPU, and —
on a GPU — whole runs compared with the oracle and, where oracle/_ref/kmdiff travelled along, with the
unmodified reference binary on the same run directory (same FASTA records as sets; the reference's
record order for the stateless correctors is thread-interleaving dependent).
"""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import kmdiff_b200 as K
import oracle
from kmdiff_b200 import build as kb
from kmdiff_b200 import kmtricks_io as kio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
THR_ARGS = ["-s", "0.05", "-u", "100000"]


@pytest.fixture(scope="module")
def cli():
    path = kb.build_cli()
    assert path and os.path.exists(path)
    return path


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, **kw)


def make_run_dir(tmp_path, shape=(10, 10, 31, 4), n_rows=120_000, parts=5, lz4=False, seed=31, ppm=3000):
    nc, nk, k, b = shape
    spec = K.SynthSpec(nc, nk, kmer_size=k, count_bytes=b, total_rows=n_rows, seed=seed, enriched_ppm=ppm)
    d = str(tmp_path / "run")
    kio.write_run_dir(d, spec, n_rows, parts, lz4=lz4)
    return spec, d


def read_fasta(path):
    recs = []
    with open(path) as f:
        head, seq = None, []
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if head is not None:
                    recs.append((head, "".join(seq)))
                head, seq = line, []
            else:
                seq.append(line)
        if head is not None:
            recs.append((head, "".join(seq)))
    return recs


def strip_index(head):
    return head.split("_", 1)[1]


# ---- CPU: format code ------------------------------------------------------------------------------

def test_lz4_frame_reader_decodes_real_lz4_frames(cli, tmp_path):
    """a matrix compressed by liblz4 (via pyarrow: LZ4F default preferences, linked 64 KiB blocks — what
    kmtricks' lz4_stream writes) decodes to the original payload"""
    spec = K.SynthSpec(10, 10, total_rows=60_000, seed=5)
    rows = spec.rows_host(60_000)
    path = str(tmp_path / "matrix_0.count.lz4")
    kio.write_matrix(path, rows, 31, 4, 20, 0, lz4=True)
    assert os.path.getsize(path) < rows.nbytes * 0.8          # it really is compressed
    out = subprocess.run([cli, "dev-cat", path, "45"], capture_output=True, timeout=120)
    assert out.returncode == 0
    assert np.array_equal(np.frombuffer(out.stdout, np.uint8), rows)
    # highly repetitive input: long overlapping matches
    rep = np.tile(np.arange(7, dtype=np.uint8), 300_000)
    kio.write_matrix(path, rep, 31, 4, 20, 0, lz4=True)
    out = subprocess.run([cli, "dev-cat", path, "45"], capture_output=True, timeout=120)
    assert np.array_equal(np.frombuffer(out.stdout, np.uint8), rep)


def test_lz4_frame_writer_is_readable_by_liblz4(cli, tmp_path):
    import pyarrow as pa
    data = np.random.default_rng(0).integers(0, 256, 300_001, dtype=np.uint8)
    src, dst = tmp_path / "in.bin", tmp_path / "out.lz4"
    data.tofile(src)
    assert run([cli, "dev-pack", str(src), str(dst)]).returncode == 0
    with pa.CompressedInputStream(pa.OSFile(str(dst)), "lz4") as s:
        back = s.read()
    assert back == data.tobytes()
    out = subprocess.run([cli, "dev-cat", str(dst), "0", "lz4"], capture_output=True, timeout=60)
    assert out.stdout == data.tobytes()
    # compressible inputs go through the block compressor: matrix rows (short matches, long literal runs), long
    # overlapping matches (period 7), a run of zeros, and sizes around the 64 KiB block and the 12/5-byte end rules
    spec = K.SynthSpec(10, 10, total_rows=30_000, seed=5)
    cases = [spec.rows_host(30_000), np.tile(np.arange(7, dtype=np.uint8), 100_003), np.zeros(65536 * 2 + 11, np.uint8)]
    cases += [np.frombuffer(bytes(range(256)) * 300, np.uint8)[:n] for n in (0, 1, 4, 5, 11, 12, 13, 14, 65535, 65536, 65537)]
    for data in cases:
        data.tofile(src)
        assert run([cli, "dev-pack", str(src), str(dst)]).returncode == 0
        with pa.CompressedInputStream(pa.OSFile(str(dst)), "lz4") as s:
            back = s.read()
        assert back == data.tobytes(), data.size
        out = subprocess.run([cli, "dev-cat", str(dst), "0", "lz4"], capture_output=True, timeout=60)
        assert out.stdout == data.tobytes()
    assert os.path.getsize(dst) < 1000                       # 65537 bytes of a 256-byte period
    cases[0].tofile(src)
    run([cli, "dev-pack", str(src), str(dst)])
    assert os.path.getsize(dst) < 0.8 * cases[0].nbytes      # synthetic count rows really shrink


def test_run_dir_config_and_totals(cli, tmp_path):
    spec, d = make_run_dir(tmp_path, n_rows=2000, parts=3)
    r = run([cli, "dev-config", d, "10", "10"])
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    assert lines[0] == "kmer_size=31 abundance_min=1 nb_partitions=3"
    totals = [int(x.split()[1]) for x in lines[1:21]]
    assert totals == [int(t) for t in spec.totals()]
    tc, tk = kio.get_total_kmer(d, 10, 10, 1)
    assert totals == tc + tk
    assert [x.split()[1] for x in lines[21:]] == kio.list_matrices(d)
    # abundance-min correction of the totals (kmtricks_utils.cpp:93-107): fof "! 3" subtracts 1*h[0] + 2*h[1]
    hu = np.arange(1, 256, dtype=np.uint64)
    kio.write_hist(os.path.join(d, "histograms", "C0.hist"), 31, 0, 10**6, hist_u=hu)
    with open(os.path.join(d, "kmtricks.fof")) as f:
        fof = f.read().replace("C0 : /dev/null/C0.fa", "C0 : /dev/null/C0.fa ! 3")
    with open(os.path.join(d, "kmtricks.fof"), "w") as f:
        f.write(fof)
    r = run([cli, "dev-config", d, "10", "10"])
    assert int(r.stdout.split("\n")[1].split()[1]) == 10**6 - 1 * 1 - 2 * 2


def test_cli_argument_checks(cli, tmp_path):
    r = run([cli, "diff", "-d", str(tmp_path), "-1", "1", "-2", "1"])
    assert r.returncode == 1 and "Not a kmtricks runtime directory" in r.stderr
    spec, d = make_run_dir(tmp_path, n_rows=100, parts=1)
    assert "Not in range [0.0, 0.5]" in run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-s", "0.7"]).stderr
    assert run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-c", "fdr"]).returncode == 1
    assert run([cli, "diff", "-d", d, "-1", "x", "-2", "10"]).returncode == 1
    assert run([cli, "--version"]).stdout.startswith("kmdiff-b200")


@pytest.mark.skipif(K._lib.load().kmd_device_count() > 0, reason="a GPU is present")
def test_cli_without_gpu_fails_loudly_with_exit_code_zero(cli, tmp_path):
    """like the reference, errors are logged and the exit code stays 0 (main.cc:93-102); and there is no CPU path"""
    spec, d = make_run_dir(tmp_path, n_rows=100, parts=1)
    r = run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-o", str(tmp_path / "out")])
    assert r.returncode == 0
    assert re.search(r"\[error\] GpuError - .*no CPU fallback", r.stderr)
    assert not os.path.exists(tmp_path / "out" / "case_kmers.fasta")


def test_example_plugin_builds_against_public_header(tmp_path):
    so = tmp_path / "libex_model.so"
    r = run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
             os.path.join(ROOT, "plugins", "ex_model.cpp"), "-o", str(so)])
    assert r.returncode == 0, r.stderr
    syms = run(["nm", "-D", str(so)]).stdout
    for s in ("plugin_name", "create8", "create16", "create32"):
        assert re.search(rf"\bT {s}\b", syms)


def test_our_plugin_built_against_reference_headers_loads_in_the_reference_binary(tmp_path):
    """plugin ABI, direction 1 (b2): plugins/ex_model.cpp compiled against the REFERENCE's include/kmdiff/imodel.hpp
    (oracle/_ref/libex_model_refhdr.so, `make -C oracle plugins`) is loaded by the unmodified reference binary with
    --cmodel and flags exactly the rows the model's rule flags"""
    if not (os.path.exists(oracle.REF_BIN) and os.path.exists(oracle.OUR_PLUGIN_REF_HEADERS)):
        pytest.skip("oracle/_ref/kmdiff or the plugin build is absent")
    n = 30_000
    spec, d = make_run_dir(tmp_path, n_rows=n, parts=3, ppm=5000)
    out = str(tmp_path / "out")
    r = run([oracle.REF_BIN, "diff", "--km-run", d, "-1", "10", "-2", "10", "-o", out, "--cmodel", oracle.OUR_PLUGIN_REF_HEADERS,
             "--config", "3.0", "-t", "2"])
    assert "[error]" not in r.stderr and "Plugin 'FoldChangeModel' loaded." in r.stderr, r.stderr
    counts = spec.rows_host(n).reshape(n, -1)[:, 8:].copy().view(np.uint32).reshape(n, 20).astype(np.float64)
    mc, mk = counts[:, :10].mean(axis=1), counts[:, 10:].mean(axis=1)
    up, down = mk > 3.0 * (mc + 1), mc > 3.0 * (mk + 1)
    assert len(read_fasta(os.path.join(out, "case_kmers.fasta"))) == int(up.sum()) > 0
    assert len(read_fasta(os.path.join(out, "control_kmers.fasta"))) == int((down & ~up).sum())


@pytest.mark.gpu
def test_reference_plugin_built_against_our_headers_loads_in_the_cli(cli, tmp_path):
    """plugin ABI, direction 2 (b2): the REFERENCE's plugins/ex_model.cpp compiled against this repo's
    include/kmdiff/imodel.hpp (oracle/_ref/libref_ex_model.so) is loaded by kmdiff-b200 with --cmodel.  That model
    calls every row CONTROL with p = 1e-12, so every row is written, with the plain means in the header."""
    if not os.path.exists(oracle.REF_PLUGIN_OUR_HEADERS):
        pytest.skip("oracle/_ref/libref_ex_model.so is absent (make -C oracle plugins)")
    n = 5_000
    spec, d = make_run_dir(tmp_path, n_rows=n, parts=2)
    out = str(tmp_path / "out")
    r = run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-o", out, "--cmodel", oracle.REF_PLUGIN_OUR_HEADERS, "-t", "2"])
    assert "[error]" not in r.stderr and "Plugin 'ExModel' loaded." in r.stderr, r.stderr
    assert f"{n}/{n} significant k-mers." in r.stderr
    recs = read_fasta(os.path.join(out, "control_kmers.fasta"))
    assert len(recs) == n and read_fasta(os.path.join(out, "case_kmers.fasta")) == []
    rows = spec.rows_host(n).reshape(n, -1)
    counts = rows[:, 8:].copy().view(np.uint32).reshape(n, 20).astype(np.float64)
    want = {}
    for i in range(n):
        seq = oracle.kmer_to_string(rows[i, :8].copy().view(np.uint64), 31)
        want[seq] = strip_index(oracle.fasta_record(0, 1e-12, counts[i, :10].mean(), counts[i, 10:].mean(), seq).split("\n")[0])
    assert {s: strip_index(h) for h, s in recs} == want
    if os.path.exists(oracle.REF_BIN):      # and the reference binary with its own plugin says the same
        out_ref = str(tmp_path / "ref")
        os.remove(os.path.join(d, "kmdiff-count.opt"))
        rr = run([oracle.REF_BIN, "diff", "--km-run", d, "-1", "10", "-2", "10", "-o", out_ref, "--cmodel", oracle.REF_PLUGIN_OUR_HEADERS, "-t", "2"])
        assert "[error]" not in rr.stderr, rr.stderr
        assert sorted((strip_index(h), s) for h, s in read_fasta(os.path.join(out_ref, "control_kmers.fasta"))) == \
            sorted((strip_index(h), s) for h, s in recs)


@pytest.mark.gpu
@pytest.mark.parametrize("lz4", [False, True])
def test_cli_reports_device_errors_and_input_errors_without_hanging(cli, tmp_path, lz4):
    """an error raised while the reader threads are running (here: the survivor buffers overflow; a truncated matrix)
    must end the run with the reference's `[error] Name - message` line and exit code 0 — every reader joined, no
    std::terminate from a joinable thread, no reader left waiting for a buffer"""
    n = 400_000
    spec, d = make_run_dir(tmp_path, n_rows=n, parts=6, lz4=lz4, ppm=20_000)
    r = run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-o", str(tmp_path / "o1"), "-t", "8", "--max-survivors", "64"] + THR_ARGS)
    assert r.returncode == 0 and re.search(r"\[error\] GpuError - .*overflow", r.stderr), r.stderr
    assert "terminate" not in r.stderr
    # a matrix cut in the middle of a row
    victim = sorted(os.listdir(os.path.join(d, "matrices")))[2]
    path = os.path.join(d, "matrices", victim)
    size = os.path.getsize(path)
    with open(path, "r+b") as f:
        f.truncate(size - 37)
    os.remove(os.path.join(d, "kmdiff-count.opt")) if os.path.exists(os.path.join(d, "kmdiff-count.opt")) else None
    r = run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-o", str(tmp_path / "o2"), "-t", "8"] + THR_ARGS)
    assert r.returncode == 0 and "[error]" in r.stderr and "terminate" not in r.stderr, r.stderr
    assert not os.path.exists(tmp_path / "o2" / "case_kmers.fasta")


@pytest.mark.gpu
@pytest.mark.parametrize("corr", ["benjamini", "bonferroni", "sidak"])
def test_cli_on_all_gpus_equals_one_gpu(cli, tmp_path, corr):
    """partitions sharded over every visible device: m and the BENJAMINI walk go through kmd_exchange_local (slots pulled
    over the peer links); the FASTA files must be the one-GPU run's, byte for byte"""
    from kmdiff_b200 import _lib
    n_dev = _lib.load().kmd_device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    n, parts = 300_000, 11
    spec, d = make_run_dir(tmp_path, n_rows=n, parts=parts, seed=17)
    outs = {}
    for g in (1, n_dev):
        out = str(tmp_path / f"out{g}")
        r = run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-o", out, "-c", corr, "--gpus", str(g), "-v", "debug"] + THR_ARGS)
        assert "[error]" not in r.stderr, r.stderr
        assert f"{g} device context(s) ready" in r.stderr
        outs[g] = [open(os.path.join(out, f)).read() for f in ("control_kmers.fasta", "case_kmers.fasta")]
    assert outs[1] == outs[n_dev] and len(outs[1][1]) > 1000


@pytest.mark.gpu
def test_cli_pop_correction_on_all_gpus_equals_one_gpu(cli, tmp_path):
    """--pop-correction with the partitions (hence the sampled k-mers) spread over every device: the integer Gram
    partials are summed exactly (kmd_popstrat_gram_local), so the .geno file, the PCs and the FASTA files are the
    one-GPU run's, byte for byte"""
    from kmdiff_b200 import _lib
    n_dev = _lib.load().kmd_device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    n, parts = 200_000, 9
    spec, d = make_run_dir(tmp_path, shape=(12, 12, 31, 4), n_rows=n, parts=parts, seed=23, ppm=4000)
    outs = {}
    for g in (1, n_dev):
        out = str(tmp_path / f"out{g}")
        r = run([cli, "diff", "-d", d, "-1", "12", "-2", "12", "-o", out, "--gpus", str(g), "--pop-correction", "--kmer-pca", "0.02",
                 "--n-pc", "3", "--random-seed", "7", "-t", "1", "-c", "disabled"] + THR_ARGS)
        assert "[error]" not in r.stderr, r.stderr
        outs[g] = [open(os.path.join(out, f)).read() for f in ("control_kmers.fasta", "case_kmers.fasta", "popstrat/pcs.evec",
                                                                "popstrat/gwas_eigenstratX.geno", "popstrat/gwas_eigenstrat.eval")]
    assert outs[1] == outs[n_dev]
    assert len(outs[1][3]) > 10_000 and len(outs[1][2]) > 100     # thousands of sampled k-mers, PCs for every sample


# ---- GPU: whole runs ---------------------------------------------------------------------------------

def expected_records(spec, n_rows, parts, corr, sig=0.05, cutoff=100000):
    """what `kmdiff diff` must write, from the oracle: (control set, case set) of 'header-without-index\\nseq'"""
    rows = spec.rows_host(n_rows)
    m = oracle.Model(spec.nb_controls, spec.nb_cases, spec.totals())
    want, *_ = m.diff(rows, spec.kmer_words, spec.count_bytes, sig / cutoff)
    keep = oracle.aggregate(want["p"], oracle.CORRECTIONS[corr], sig, n_rows)
    ctl, case = [], []
    for r in want[keep]:
        seq = oracle.kmer_to_string(r["kmer"][:spec.kmer_words], spec.kmer_size)
        rec = oracle.fasta_record(0, r["p"], r["mean_control"], r["mean_case"], seq)
        head, body = rec.split("\n", 1)
        (ctl if r["sign"] == oracle.CONTROL else case).append((strip_index(head), body.replace("\n", "")))
    return ctl, case, len(want)


@pytest.mark.gpu
@pytest.mark.parametrize("corr", ["bonferroni", "benjamini", "sidak", "holm", "disabled"])
@pytest.mark.parametrize("lz4", [False, True])
def test_cli_matches_oracle(cli, tmp_path, corr, lz4):
    n, parts = 120_000, 5
    spec, d = make_run_dir(tmp_path, n_rows=n, parts=parts, lz4=lz4)
    out = str(tmp_path / "out")
    sig = "1e-9" if corr == "disabled" else "0.05"
    r = run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-o", out, "-c", corr, "-s", sig, "-u", "100000" if corr != "disabled" else "1"])
    assert r.returncode == 0 and "[error]" not in r.stderr, r.stderr
    ctl, case, n_pref = expected_records(spec, n, parts, corr, float(sig), 100000 if corr != "disabled" else 1)
    got_ctl = [(strip_index(h), s) for h, s in read_fasta(os.path.join(out, "control_kmers.fasta"))]
    got_case = [(strip_index(h), s) for h, s in read_fasta(os.path.join(out, "case_kmers.fasta"))]
    assert len(ctl) + len(case) > 50
    if corr in ("benjamini", "holm"):
        assert got_ctl == ctl and got_case == case        # ascending p, ties by row: fully ordered
    else:
        assert sorted(got_ctl) == sorted(ctl) and sorted(got_case) == sorted(case)
    assert f"{n_pref}/{n} significant k-mers." in r.stderr
    assert f"Significant k-mers: {len(ctl)} (control), {len(case)} (case)." in r.stderr
    # running index per file (aggregator.hpp:55-62)
    heads = [h for h, _ in read_fasta(os.path.join(out, "case_kmers.fasta"))]
    assert [int(h[1:].split("_")[0]) for h in heads] == list(range(len(heads)))
    with open(os.path.join(out, "options.bin"), "rb") as f:
        ob = f.read()
    thr, cutoff, c, pop = struct.unpack("<ddiB", ob[:21])
    assert (thr, c, pop) == (float(sig), oracle.CORRECTIONS[corr], 0) and struct.unpack("<Q", ob[37:45])[0] == n


@pytest.mark.gpu
@pytest.mark.parametrize("shape,cb", [((50, 50, 63, 2), 2), ((12, 20, 21, 1), 1)])
def test_cli_other_count_widths_and_wide_kmers(cli, tmp_path, shape, cb):
    n, parts = 50_000, 3
    spec, d = make_run_dir(tmp_path, shape=shape, n_rows=n, parts=parts)
    out = str(tmp_path / "out")
    r = run([cli, "diff", "-d", d, "-1", str(shape[0]), "-2", str(shape[1]), "-o", out, "--count-bytes", str(cb), "--keep-tmp"])
    assert "[error]" not in r.stderr, r.stderr
    ctl, case, n_pref = expected_records(spec, n, parts, "bonferroni")
    got = [(strip_index(h), s) for f in ("control_kmers.fasta", "case_kmers.fasta") for h, s in read_fasta(os.path.join(out, f))]
    assert sorted(got) == sorted(ctl + case) and len(got) > 10
    if shape[2] > 60:
        assert all(len(s) == 63 for _, s in got)
        with open(os.path.join(out, "case_kmers.fasta")) as f:
            lines = f.read().split("\n")
        assert len(lines[1]) == 60 and len(lines[2]) == 3      # kseq++ wraps sequences at 60 columns
    # --keep-tmp: spill files in the record layout of a default (WITH_POPSTRAT) reference build:
    # kmer words | f64 p | i32 sign | f64 | f64 | u16 N | N x f64 counts
    nsamp = shape[0] + shape[1]
    rec = 8 * spec.kmer_words + 30 + 8 * nsamp
    total = 0
    for p in range(parts):
        raw = subprocess.run([cli, "dev-cat", os.path.join(out, "partitions", f"p{p}_uncorrected"), "0", "lz4"],
                             capture_output=True, timeout=60).stdout
        assert len(raw) % rec == 0
        total += len(raw) // rec
    assert total == n_pref


@pytest.mark.gpu
def test_cli_matches_reference_binary(cli, tmp_path):
    """the unmodified reference `kmdiff diff` and kmdiff-b200 on the same run directory"""
    if not os.path.exists(oracle.REF_BIN):
        pytest.skip("oracle/_ref/kmdiff not built")
    n, parts = 200_000, 8
    spec, d = make_run_dir(tmp_path, n_rows=n, parts=parts, seed=99)
    for corr in ("bonferroni", "benjamini"):
        out_ref, out_new = str(tmp_path / f"ref_{corr}"), str(tmp_path / f"new_{corr}")
        opt = os.path.join(d, "kmdiff-count.opt")
        if os.path.exists(opt):
            os.remove(opt)
        r = run([oracle.REF_BIN, "diff", "--km-run", d, "-1", "10", "-2", "10", "-o", out_ref, "-c", corr, "-t", "4"] + THR_ARGS)
        assert "[error]" not in r.stderr, r.stderr
        if os.path.exists(opt):
            os.remove(opt)
        r2 = run([cli, "diff", "--km-run", d, "-1", "10", "-2", "10", "-o", out_new, "-c", corr, "-t", "4"] + THR_ARGS)
        assert "[error]" not in r2.stderr, r2.stderr
        for f in ("control_kmers.fasta", "case_kmers.fasta"):
            a = sorted((strip_index(h), s) for h, s in read_fasta(os.path.join(out_ref, f)))
            b = sorted((strip_index(h), s) for h, s in read_fasta(os.path.join(out_new, f)))
            assert len(a) > 5 and a == b, (corr, f, len(a), len(b))
        # the summary lines the reference prints
        for pat in (r"\d+/\d+ significant k-mers\.", r"Before correction: \d+ \(control\), \d+ \(case\)\.",
                    r"Significant k-mers: \d+ \(control\), \d+ \(case\)\."):
            assert re.search(pat, r.stderr).group(0) == re.search(pat, r2.stderr).group(0)


@pytest.mark.gpu
def test_cli_plugin_model(cli, tmp_path):
    so = tmp_path / "libex_model.so"
    assert run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                os.path.join(ROOT, "plugins", "ex_model.cpp"), "-o", str(so)]).returncode == 0
    n = 40_000
    spec, d = make_run_dir(tmp_path, n_rows=n, parts=4, ppm=5000)
    out = str(tmp_path / "out")
    r = run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-o", out, "--cmodel", str(so), "--config", "3.0", "-t", "3"])
    assert "[error]" not in r.stderr and "Plugin 'FoldChangeModel' loaded." in r.stderr, r.stderr
    counts = spec.rows_host(n).reshape(n, -1)[:, 8:].copy().view(np.uint32).reshape(n, 20).astype(np.float64)
    mc, mk = counts[:, :10].mean(axis=1), counts[:, 10:].mean(axis=1)
    up, down = mk > 3.0 * (mc + 1), mc > 3.0 * (mk + 1)
    got_case = read_fasta(os.path.join(out, "case_kmers.fasta"))
    got_ctl = read_fasta(os.path.join(out, "control_kmers.fasta"))
    assert len(got_case) == int(up.sum()) > 0 and len(got_ctl) == int((down & ~up).sum())
    assert f"{int((up | down).sum())}/{n} significant k-mers." in r.stderr


# ---- run directories without matrices: per-sample count files merged on the device (row a3) ---------

@pytest.mark.gpu
@pytest.mark.parametrize("merge_records", [0, 20_000])
def test_cli_count_files_match_matrix_mode_and_reference(cli, tmp_path, merge_records):
    """`kmtricks pipeline --until count` layout (counts/partition_p/<sample>.kmer.lz4, no matrices): the k-mers,
    FASTA records and summary lines equal those of matrix mode on the same cohort and of the reference binary;
    a small --merge-records forces many k-mer-range chunks per partition."""
    n, parts = 150_000, 6
    spec = K.SynthSpec(10, 10, kmer_size=31, count_bytes=4, total_rows=n, seed=77, enriched_ppm=3000)
    d = str(tmp_path / "run_cf")
    kio.write_run_dir(d, spec, n, parts, lz4=True, count_files=True)
    out = str(tmp_path / "out_cf")
    extra = ["--merge-records", str(merge_records)] if merge_records else []
    for corr in ("bonferroni", "benjamini"):
        r = run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-o", out, "-c", corr, "-t", "4"] + THR_ARGS + extra)
        assert r.returncode == 0 and "[error]" not in r.stderr, r.stderr
        ctl, case, n_pref = expected_records(spec, n, parts, corr)
        got_ctl = [(strip_index(h), s) for h, s in read_fasta(os.path.join(out, "control_kmers.fasta"))]
        got_case = [(strip_index(h), s) for h, s in read_fasta(os.path.join(out, "case_kmers.fasta"))]
        assert len(ctl) + len(case) > 50
        assert sorted(got_ctl) == sorted(ctl) and sorted(got_case) == sorted(case)
        assert f"{n_pref}/{n} significant k-mers." in r.stderr
        if os.path.exists(oracle.REF_BIN) and not merge_records:
            out_ref = str(tmp_path / f"ref_{corr}")
            rr = run([oracle.REF_BIN, "diff", "--km-run", d, "-1", "10", "-2", "10", "-o", out_ref, "-c", corr, "-t", "4"] + THR_ARGS)
            assert "[error]" not in rr.stderr, rr.stderr
            for f in ("control_kmers.fasta", "case_kmers.fasta"):
                a = sorted((strip_index(h), s) for h, s in read_fasta(os.path.join(out_ref, f)))
                b = sorted((strip_index(h), s) for h, s in read_fasta(os.path.join(out, f)))
                assert a == b, (corr, f, len(a), len(b))
            for pat in (r"\d+/\d+ significant k-mers\.", r"Before correction: \d+ \(control\), \d+ \(case\)\.",
                        r"Significant k-mers: \d+ \(control\), \d+ \(case\)\."):
                assert re.search(pat, rr.stderr).group(0) == re.search(pat, r.stderr).group(0)
    assert not os.path.exists(os.path.join(d, "kmdiff-count.opt"))     # only matrix mode leaves the marker


@pytest.mark.gpu
def test_cli_reference_fixture_km_out_dir(cli, tmp_path, km_out_dir):
    """the reference's tests/data_test/km_out_dir (merge_test.cpp): 1 control + 1 case, k = 20, 4 partitions"""
    d = km_out_dir
    out = str(tmp_path / "out")
    r = run([cli, "diff", "-d", d, "-1", "1", "-2", "1", "-o", out, "-u", "10000", "--log-factorial", "100"])
    assert r.returncode == 0 and "[error]" not in r.stderr, r.stderr
    assert "0/320 significant k-mers." in r.stderr
    assert "Before correction: 0 (control), 0 (case)." in r.stderr
    if os.path.exists(oracle.REF_BIN):
        rr = run([oracle.REF_BIN, "diff", "--km-run", d, "-1", "1", "-2", "1", "-o", str(tmp_path / "ref"), "-u", "10000"])
        assert "0/320 significant k-mers." in rr.stderr, rr.stderr


@pytest.mark.gpu
def test_cli_count_files_wide_kmers_two_byte_counts(cli, tmp_path):
    n, parts = 40_000, 3
    spec = K.SynthSpec(6, 7, kmer_size=63, count_bytes=2, total_rows=n, seed=5, enriched_ppm=4000)
    d = str(tmp_path / "run_cf")
    kio.write_run_dir(d, spec, n, parts, lz4=False, count_files=True)
    out = str(tmp_path / "out")
    r = run([cli, "diff", "-d", d, "-1", "6", "-2", "7", "-o", out, "--count-bytes", "2", "--merge-records", "50000"])
    assert "[error]" not in r.stderr, r.stderr
    ctl, case, n_pref = expected_records(spec, n, parts, "bonferroni")
    got = [(strip_index(h), s) for f in ("control_kmers.fasta", "case_kmers.fasta") for h, s in read_fasta(os.path.join(out, f))]
    assert sorted(got) == sorted(ctl + case) and len(got) > 10
    assert f"{n_pref}/{n} significant k-mers." in r.stderr


# ---- resume from an earlier run's spill files (cmd/diff.hpp:277-303, 340-349) ---------------------------

@pytest.mark.gpu
def test_cli_resume_recorrects_from_spill_files(cli, tmp_path):
    n, parts = 120_000, 5
    spec, d = make_run_dir(tmp_path, n_rows=n, parts=parts)
    out = str(tmp_path / "out")
    base = [cli, "diff", "-d", d, "-1", "10", "-2", "10", "-o", out, "--keep-tmp"] + THR_ARGS
    r = run(base + ["-c", "bonferroni"])
    assert "[error]" not in r.stderr and "Process partitions" in r.stderr, r.stderr
    # same prefilter, other correction: the scan is skipped, the kept spill files are re-corrected with m = n
    r2 = run(base + ["-c", "benjamini"])
    assert "[error]" not in r2.stderr and "Process partitions" not in r2.stderr, r2.stderr
    ctl, case, _ = expected_records(spec, n, parts, "benjamini")
    got_ctl = [(strip_index(h), s) for h, s in read_fasta(os.path.join(out, "control_kmers.fasta"))]
    got_case = [(strip_index(h), s) for h, s in read_fasta(os.path.join(out, "case_kmers.fasta"))]
    assert len(ctl) + len(case) > 50
    assert sorted(got_ctl) == sorted(ctl) and sorted(got_case) == sorted(case)
    assert f"Significant k-mers: {len(ctl)} (control), {len(case)} (case)." in r2.stderr
    with open(os.path.join(out, "options.bin"), "rb") as f:
        ob = f.read()
    assert struct.unpack("<i", ob[16:20])[0] == oracle.CORRECTIONS["benjamini"] and struct.unpack("<Q", ob[37:45])[0] == n
    # nothing changed and the FASTA files exist: nothing is redone
    before = os.path.getmtime(os.path.join(out, "case_kmers.fasta"))
    r3 = run(base + ["-c", "benjamini"])
    assert "Aggregate" not in r3.stderr and "Done in" in r3.stderr, r3.stderr
    assert os.path.getmtime(os.path.join(out, "case_kmers.fasta")) == before
    # another cutoff: the prefilter changed, everything is redone
    r4 = run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-o", out, "--keep-tmp", "-s", "0.05", "-u", "1000", "-c", "benjamini"])
    assert "Process partitions" in r4.stderr and "[error]" not in r4.stderr, r4.stderr


def spill_records(cli, path, kmer_bytes, n_samples):
    """decoded p{i}_uncorrected file -> structured array (KmerSign::dump, kmer.hpp:113-127, WITH_POPSTRAT layout)"""
    raw = subprocess.run([cli, "dev-cat", path, "0", "lz4"], capture_output=True, timeout=120).stdout
    dt = np.dtype([("kmer", f"V{kmer_bytes}"), ("p", "<f8"), ("sign", "<i4"), ("mc", "<f8"), ("kk", "<f8"), ("n", "<u2"),
                   ("counts", "<f8", (n_samples,))])
    assert dt.itemsize == kmer_bytes + 30 + 8 * n_samples and len(raw) % dt.itemsize == 0
    return np.frombuffer(raw, dtype=dt)


@pytest.mark.gpu
def test_spill_files_interchange_with_the_reference(cli, tmp_path):
    """--keep-tmp spill files (a12): ours and the reference's hold the same records (k-mer, sign, means and counts bit
    for bit, p within 1e-9); an output directory written by the REFERENCE is re-corrected by kmdiff-b200 with another
    -c without a scan (f4) — m comes from the raw matrices' sizes, which the reference's options.bin forgets."""
    if not os.path.exists(oracle.REF_BIN):
        pytest.skip("oracle/_ref/kmdiff not built")
    n, parts = 160_000, 6
    spec, d = make_run_dir(tmp_path, n_rows=n, parts=parts, seed=123)
    out_ref, out_new = str(tmp_path / "ref"), str(tmp_path / "new")
    common = ["diff", "--km-run", d, "-1", "10", "-2", "10", "-c", "bonferroni", "-t", "4", "--keep-tmp"] + THR_ARGS
    r = run([oracle.REF_BIN] + common + ["-o", out_ref])
    assert "[error]" not in r.stderr, r.stderr
    os.remove(os.path.join(d, "kmdiff-count.opt"))
    r = run([cli] + common + ["-o", out_new])
    assert "[error]" not in r.stderr, r.stderr
    # (matrix mode: the reference numbers its accumulators by directory-iteration index, not by the header's partition
    # field, so the files are compared as one multiset)
    def all_records(base):
        recs = np.concatenate([spill_records(cli, os.path.join(base, "partitions", f"p{p}_uncorrected"), 8, 20) for p in range(parts)])
        return recs[np.argsort(recs["kmer"].view("<u8"), kind="stable")]

    a, b = all_records(out_ref), all_records(out_new)
    assert len(a) == len(b) > 100
    for f in ("kmer", "sign", "mc", "kk", "n", "counts"):
        assert np.array_equal(a[f], b[f]), f
    assert np.all(np.abs(a["p"] - b["p"]) <= 1e-9 * a["p"])
    n_rec = len(a)
    # re-correct the REFERENCE's output directory with another corrector: no scan, the spill files are read
    r2 = run([cli, "diff", "--km-run", d, "-1", "10", "-2", "10", "-c", "benjamini", "-t", "4", "--keep-tmp", "-o", out_ref] + THR_ARGS)
    assert "[error]" not in r2.stderr and "Process partitions" not in r2.stderr, r2.stderr
    ctl, case, _ = expected_records(spec, n, parts, "benjamini")
    got = sorted((strip_index(h), s) for f in ("control_kmers.fasta", "case_kmers.fasta") for h, s in read_fasta(os.path.join(out_ref, f)))
    assert got == sorted(ctl + case) and len(got) > 50
    # and the other way round: the reference resumes from OUR spill files (it corrects with m = 0 then, i.e. keeps
    # every prefilter survivor — SURVEY.md §5 — so only the record count can be checked)
    os.remove(os.path.join(d, "kmdiff-count.opt")) if os.path.exists(os.path.join(d, "kmdiff-count.opt")) else None
    r3 = run([oracle.REF_BIN, "diff", "--km-run", d, "-1", "10", "-2", "10", "-c", "benjamini", "-t", "4", "--keep-tmp", "-o", out_new] + THR_ARGS)
    assert "[error]" not in r3.stderr and "Process partitions" not in r3.stderr, r3.stderr
    n_ref = sum(len(read_fasta(os.path.join(out_new, f))) for f in ("control_kmers.fasta", "case_kmers.fasta"))
    assert n_ref == n_rec


@pytest.mark.gpu
@pytest.mark.parametrize("count_files", [False, True])
def test_save_sk_matches_the_reference(cli, tmp_path, count_files):
    """--save-sk (f3): positive_kmer_matrix/matrices/matrix_{p}.count.lz4 hold the rows that passed the prefilter, in
    km::MatrixWriter<65536>'s container (45-byte header with compressed = 1 and count_slots = 4, then one LZ4 frame);
    header fields and decoded payloads equal the reference's (merge.hpp:270-278, 83-86)"""
    n, parts = 90_000, 4
    spec = K.SynthSpec(10, 10, kmer_size=31, count_bytes=4, total_rows=n, seed=5, enriched_ppm=3000)
    d = str(tmp_path / "run")
    kio.write_run_dir(d, spec, n, parts, count_files=count_files)
    for sub in ("config_gatb", "repartition_gatb"):      # copy_kdir (cmd/diff.hpp:52-64) needs them; kmtricks writes them
        os.makedirs(os.path.join(d, sub), exist_ok=True)
        with open(os.path.join(d, sub, "marker.txt"), "w") as f:
            f.write(sub)
    out = str(tmp_path / "out")
    r = run([cli, "diff", "--km-run", d, "-1", "10", "-2", "10", "-o", out, "--save-sk", "-t", "4"] + THR_ARGS)
    assert "[error]" not in r.stderr, r.stderr
    root = os.path.join(out, "positive_kmer_matrix")
    assert os.path.exists(os.path.join(root, "kmtricks.fof")) and os.path.exists(os.path.join(root, "options.txt"))
    assert os.path.exists(os.path.join(root, "config_gatb", "marker.txt"))

    def load(base):
        res = {}
        for p in range(parts):
            path = os.path.join(base, "positive_kmer_matrix", "matrices", f"matrix_{p}.count.lz4")
            with open(path, "rb") as f:
                hdr = kio.MatrixHeader.unpack(f.read(45))
            body = subprocess.run([cli, "dev-cat", path, "45"], capture_output=True, timeout=120).stdout
            assert hdr.compressed == 1 and hdr.count_slots == 4 and hdr.nb_counts == 20 and hdr.kmer_size == 31 and hdr.id == 0
            res[hdr.partition] = body
        return res

    ours = load(out)
    # the rows are the oracle's prefilter survivors, partition by partition, in row order
    rows = spec.rows_host(n).reshape(n, spec.row_bytes)
    want, _, _ = oracle.Model(10, 10, spec.totals()).diff(rows.ravel(), 1, 4, 0.05 / 100000)
    per = (n + parts - 1) // parts
    total = 0
    for p in range(parts):
        idx = want["row"][(want["row"] >= p * per) & (want["row"] < (p + 1) * per)]
        assert ours[p] == rows[idx].tobytes(), p
        total += len(idx)
    assert total == len(want) > 50
    if os.path.exists(oracle.REF_BIN):
        out_ref = str(tmp_path / "ref")
        opt = os.path.join(d, "kmdiff-count.opt")
        if os.path.exists(opt) and not count_files:
            os.remove(opt)
        rr = run([oracle.REF_BIN, "diff", "--km-run", d, "-1", "10", "-2", "10", "-o", out_ref, "--save-sk", "-t", "4"] + THR_ARGS)
        assert "[error]" not in rr.stderr, rr.stderr
        ref = load(out_ref)
        # matrix mode names the reference's files by directory-iteration index, not by partition: compare as multisets
        assert sorted(ref.values()) == sorted(ours.values())
        if count_files:
            assert ref == ours


# ---- -f/--kff-output (include/kmdiff/kff_utils.hpp:32-107) -------------------------------------------------

def kff_kmers(raw: bytes, k: int):
    """minimal KFF reader for the files KffWriter leaves: -> list of k-mer strings"""
    assert raw[:3] == b"KFF" and raw[-3:] == b"KFF" and raw[5] == 0x1E
    pos = 8 + 4 + int.from_bytes(raw[8:12], "big")
    assert raw[pos:pos + 1] == b"v"
    nvars = int.from_bytes(raw[pos + 1:pos + 9], "big")
    pos += 9
    gv = {}
    for _ in range(nvars):
        end = raw.index(b"\0", pos)
        gv[raw[pos:end].decode()] = int.from_bytes(raw[end + 1:end + 9], "big")
        pos = end + 9
    assert gv == {"data_size": 0, "k": k, "max": 1}
    assert raw[pos:pos + 1] == b"r"
    nblocks = int.from_bytes(raw[pos + 1:pos + 9], "big")
    pos += 9
    nbytes, out = (k + 3) // 4, []
    for _ in range(nblocks):
        v = int.from_bytes(raw[pos:pos + nbytes], "big")
        out.append("".join("ACTG"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k)))
        pos += nbytes
    assert raw[pos:pos + 1] == b"i"
    return out


def test_kff_writer_matches_reference_files(cli, tmp_path):
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kff_vectors.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) == 6
    for c in cases:
        out = str(tmp_path / "o.kff")
        assert run([cli, "dev-kff", out, str(c["k"])] + c["kmers"]).returncode == 0
        with open(out, "rb") as f:
            got = f.read()
        assert got == bytes.fromhex(c["kff"]), (c["k"], c["file"])
        assert kff_kmers(got, c["k"]) == c["kmers"]
    # 128-bit k-mers (the reference truncates them, SURVEY.md §8 hazard 1): checked by decoding
    kmer = "ACGT" * 15 + "TGA"
    out = str(tmp_path / "w.kff")
    assert run([cli, "dev-kff", out, "63", kmer, kmer[::-1]]).returncode == 0
    with open(out, "rb") as f:
        assert kff_kmers(f.read(), 63) == [kmer, kmer[::-1]]


@pytest.mark.gpu
def test_cli_kff_output(cli, tmp_path):
    n, parts = 60_000, 3
    spec, d = make_run_dir(tmp_path, n_rows=n, parts=parts)
    out = str(tmp_path / "out")
    r = run([cli, "diff", "-d", d, "-1", "10", "-2", "10", "-o", out, "-f"] + THR_ARGS)
    assert "[error]" not in r.stderr, r.stderr
    ctl, case, _ = expected_records(spec, n, parts, "bonferroni")
    for name, want in (("control", ctl), ("case", case)):
        with open(os.path.join(out, f"{name}_kmers.kff"), "rb") as f:
            got = kff_kmers(f.read(), 31)
        assert sorted(got) == sorted(s for _, s in want) and len(got) > 3
    assert not os.path.exists(os.path.join(out, "case_kmers.fasta"))
