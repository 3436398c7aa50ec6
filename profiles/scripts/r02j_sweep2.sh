O=gpurun_out/r02j; mkdir -p $O
for c in c3 c2 c5 c4; do SWEEP_COMBO=1 SWEEP_OUT=r02j/sweep_$c.json timeout 900 python profiles/sweep_scan.py $c > $O/sweep_$c.log 2>&1; tail -1 $O/sweep_$c.log; done
timeout 600 python -m pytest tests/test_gpu_popstrat.py -m gpu -q > $O/pytest_pop.log 2>&1; echo "pytest rc=$?" >> $O/pytest_pop.log; tail -3 $O/pytest_pop.log
timeout 600 python bench.py --config c4 --steps 5 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?"
