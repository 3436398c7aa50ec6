O=gpurun_out/r02k; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_popstrat.py tests/test_distributed.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
for c in c3 c2 c5 c4; do timeout 600 python profiles/sweep_scan.py $c 2>&1 | grep default >> $O/defaults.log; done
cat $O/defaults.log
( time python bench.py --no-cli ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
