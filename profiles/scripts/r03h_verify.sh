# round 2, final tree: full verification + fresh ncu evidence of the union-tree kernels
O=gpurun_out/r03h; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time python bench.py ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -4 $O/bench.err
timeout 200 python bench.py --config c3m --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3m.json 2> $O/bench_c3m.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:union_level_kernel -c 2 -o $O/union_level python bench.py --config c2m --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:assemble_rows_kernel -c 1 -o $O/assemble python bench.py --config c2m --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c2m.csv python bench.py --config c2m --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu3.log 2>&1
ls -la $O
