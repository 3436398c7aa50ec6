# round 2: full GPU suite on the latest tree (no -x), smoke(), the default bench line, ncu of the GLM kernel
O=gpurun_out/r02f; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -8 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time python bench.py ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:glm_irls_kernel -s 4 -c 1 -o $O/prof_glm_c4 python bench.py --config c4 --steps 1 --warmup 3 --rows 20000000 --e2e-steps 1 --e2e-rows 262144 --no-cpu-baseline > $O/ncu_glm.log 2>&1; tail -2 $O/ncu_glm.log
ls -la $O
