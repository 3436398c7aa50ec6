# round 2: full verification on the final tree + count-file CLI timing with the two-lane pipeline
O=gpurun_out/r02l; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -8 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time python bench.py ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
# count files, LZ4-compressed, c2 cohort, 1.6e7 rows in 32 partitions: host decode, two lanes
python - <<'PY' > $O/make_cf.log 2>&1
import sys, time
sys.path.insert(0, ".")
import kmdiff_b200 as K
from kmdiff_b200 import kmtricks_io as kio
spec = K.SynthSpec(10, 10, kmer_size=31, count_bytes=4, total_rows=10**9, seed=42)
kio.write_run_dir("/dev/shm/cf_lz4", spec, 16_000_000, 32, lz4=True, count_files=True)
PY
D=/dev/shm/cf_lz4
for i in 1 2; do ( time env KMD_TIMING=1 ./kmdiff_b200/kmdiff-b200 diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/cf_out -v debug ) 2>&1 | grep -v "partition " > $O/cli_cf_lz4_run$i.log; sleep 1; done
( time ./oracle/_ref/kmdiff diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/cf_out_ref ) > $O/cli_cf_lz4_reference.log 2>&1
grep -c '>' /dev/shm/cf_out/*.fasta /dev/shm/cf_out_ref/*.fasta >> $O/cli_cf_lz4_reference.log
rm -rf /dev/shm/cf_lz4 /dev/shm/cf_out*
ls -la $O
