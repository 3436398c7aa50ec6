O=gpurun_out/r02o; mkdir -p $O
KMD_MERGE_STAGE_OUT=0 timeout 600 python -m pytest tests/test_gpu_merge.py -m gpu -q > $O/pytest_merge_direct.log 2>&1; echo "pytest rc=$?" >> $O/pytest_merge_direct.log; tail -3 $O/pytest_merge_direct.log
for v in 1 0 1 0; do KMD_MERGE_STAGE_OUT=$v timeout 600 python bench.py --config c2m --no-cpu-baseline > $O/bench_c2m_stage$v.$RANDOM.json 2>> $O/bench.err; done
timeout 600 python -m pytest "tests/test_host_cli.py" -m gpu -q > $O/pytest_cli.log 2>&1; echo "pytest rc=$?" >> $O/pytest_cli.log; tail -3 $O/pytest_cli.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02o/bench_c2m_stage*.json")):
    j=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "value %.3e"%j["value"], "passes ms %.2f"%j["roofline"]["stage_ms_per_step"]["passes"], "frac %.3f"%j["roofline"]["frac"])
PY
