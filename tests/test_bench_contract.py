"""bench.py's contract on a box without a GPU: our arm refuses to run (no CPU fallback), the JSON keys the driver
reads are spelled in the source, and every config names a workload of BASELINE.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) and "no CPU fallback" in (r.stderr + r.stdout)
    assert not any(line.startswith("{") for line in r.stdout.splitlines())     # no bench line from a box without a GPU


def test_bench_source_carries_the_contract_keys():
    with open(os.path.join(ROOT, "bench.py")) as f:
        src = f.read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"',
                '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"', '"workload"', '"roofline"', '"bound"',
                '"achieved"', '"peak"', '"frac"', '"traffic"', '"cpu_baseline"', '"cores"', '"kind"', '"sample"', '"e2e"',
                '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"gpu_launches"', '"clocks"', '"impl"'):
        assert key in src, key
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert "rows tested/sec" in base["metric"]
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.CONFIGS) >= {"c2", "c3", "c4", "c5"}
    assert bench.CONFIGS["c2"][:4] == (10, 10, 31, 4) and bench.CONFIGS["c3"][:4] == (100, 100, 31, 4)
    assert bench.CONFIGS["c4"][:2] == (200, 200) and bench.CONFIGS["c5"][:4] == (50, 50, 63, 2)


def test_reference_arm_never_maps_the_product_library():
    """`bench.py --impl reference` times the unmodified reference binary; its run directories come from a `dev-synth`
    subprocess, so libkmdiff_b200.so must never be mapped into the arm's own process (round-1 review, item 4a)"""
    import pytest
    sys.path.insert(0, ROOT)
    import oracle
    if not os.path.exists(oracle.REF_BIN):
        pytest.skip("oracle/_ref/kmdiff not built")
    code = (
        "import sys\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--config', 'c2', '--steps', '1', '--warmup', '0']\n"
        "import bench\n"
        "bench.reference_sample_rows = lambda cfg, cores: 100000\n"
        "bench.main()\n"
        "print('MAPPED' if 'libkmdiff_b200' in open('/proc/self/maps').read() else 'CLEAN', file=sys.stderr)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    assert r.stderr.strip().endswith("CLEAN"), r.stderr[-300:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["config"]["hung"] == 0
    assert line["config"]["workload"].startswith("c2: synthetic 10+10 samples")
