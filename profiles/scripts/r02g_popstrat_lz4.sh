# round 2: GLM / PCA after the rework (c4 bench + ncu of the alt-model fit), count-file CLI: device LZ4 vs host LZ4 vs reference
O=gpurun_out/r02g; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_popstrat.py tests/test_gpu_lz4.py tests/test_gpu_merge.py tests/test_smartpca_reference.py -m gpu -q > $O/pytest_pop.log 2>&1; echo "pytest rc=$?" >> $O/pytest_pop.log; tail -4 $O/pytest_pop.log
timeout 600 python bench.py --config c4 --steps 5 > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:glm_irls_kernel -s 5 -c 1 -o $O/prof_glm_c4 python bench.py --config c4 --steps 1 --warmup 3 --rows 20000000 --e2e-steps 1 --e2e-rows 262144 --no-cpu-baseline > $O/ncu_glm.log 2>&1; tail -2 $O/ncu_glm.log
# count files, LZ4-compressed (liblz4 frames, linked blocks), c2 cohort, 1.6e7 rows in 32 partitions
python - <<'PY' > $O/make_cf.log 2>&1
import sys, time
sys.path.insert(0, ".")
import kmdiff_b200 as K
from kmdiff_b200 import kmtricks_io as kio
spec = K.SynthSpec(10, 10, kmer_size=31, count_bytes=4, total_rows=10**9, seed=42)
t0 = time.time()
kio.write_run_dir("/dev/shm/cf_lz4", spec, 16_000_000, 32, lz4=True, count_files=True)
print("written in", time.time() - t0)
PY
du -sh /dev/shm/cf_lz4 >> $O/make_cf.log
D=/dev/shm/cf_lz4
for mode in device host; do for i in 1 2; do
  if [ $mode = host ]; then export KMD_HOST_LZ4=1; else unset KMD_HOST_LZ4; fi
  ( time ./kmdiff_b200/kmdiff-b200 diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/cf_out_$mode -v debug --gpus 1 ) 2>&1 | grep -v "partition " > $O/cli_cf_lz4_${mode}_run$i.log; sleep 1
done; done
unset KMD_HOST_LZ4
( time ./oracle/_ref/kmdiff diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/cf_out_ref ) > $O/cli_cf_lz4_reference.log 2>&1
cmp /dev/shm/cf_out_device/case_kmers.fasta /dev/shm/cf_out_host/case_kmers.fasta && echo "device-LZ4 FASTA == host-LZ4 FASTA" >> $O/cli_cf_lz4_device_run2.log
grep -c '>' /dev/shm/cf_out_device/*.fasta /dev/shm/cf_out_ref/*.fasta >> $O/cli_cf_lz4_reference.log
rm -rf /dev/shm/cf_lz4 /dev/shm/cf_out_*
timeout 600 python bench.py --config c2m > $O/bench_c2m.json 2> $O/bench_c2m.err; echo "c2m rc=$?"
ls -la $O
