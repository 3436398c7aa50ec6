# host LZ4 decode rate after the fast-zone rework (16 host cores of the GPU box), then the CLI / LZ4 / config-1 tests on the final binaries
O=gpurun_out/r03k; mkdir -p $O
python - <<'PY' > $O/make.log 2>&1
import sys
sys.path.insert(0, ".")
import kmdiff_b200 as K
from kmdiff_b200 import kmtricks_io as kio
spec = K.SynthSpec(10, 10, kmer_size=31, count_bytes=4, total_rows=10**9, seed=42)
kio.write_run_dir("/dev/shm/l4c", spec, 8_000_000, 16, lz4=True, count_files=True)
kio.write_run_dir("/dev/shm/l4m", spec, 8_000_000, 16, lz4=True, count_files=False)
PY
C=$(find /dev/shm/l4c -name "*.kmer.lz4" | tr '\n' ' '); M=$(find /dev/shm/l4m -name "*.count.lz4" | tr '\n' ' ')
for t in 1 4 16; do ./kmdiff_b200/kmdiff-b200 dev-lz4bench $t $C $C | sed "s/^/count files: /" >> $O/host_lz4.jsonl; ./kmdiff_b200/kmdiff-b200 dev-lz4bench $t $M $M $M $M | sed "s/^/matrices: /" >> $O/host_lz4.jsonl; done
cat $O/host_lz4.jsonl
rm -rf /dev/shm/l4c /dev/shm/l4m
timeout 300 python -m pytest tests/test_host_cli.py tests/test_config1.py tests/test_gpu_lz4.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
