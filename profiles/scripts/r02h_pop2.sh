O=gpurun_out/r02h; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_popstrat.py tests/test_gpu_merge.py tests/test_config1.py tests/test_gpu_lz4.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 600 python bench.py --config c4 --steps 5 > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?"
timeout 600 python bench.py --config c2m > $O/bench_c2m.json 2> $O/bench_c2m.err; echo "c2m rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c4.csv python bench.py --config c4 --steps 1 --warmup 3 --rows 6000000 --e2e-steps 1 --e2e-rows 262144 --no-cpu-baseline > $O/ncu_list_c4.log 2>&1
ls -la $O
