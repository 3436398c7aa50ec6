O=gpurun_out/r02m; mkdir -p $O
python - <<'PY' > $O/make_cf.log 2>&1
import sys
sys.path.insert(0, ".")
import kmdiff_b200 as K
from kmdiff_b200 import kmtricks_io as kio
spec = K.SynthSpec(10, 10, kmer_size=31, count_bytes=4, total_rows=10**9, seed=42)
kio.write_run_dir("/dev/shm/cf_lz4", spec, 16_000_000, 32, lz4=True, count_files=True)
PY
D=/dev/shm/cf_lz4
for lanes in 1 2 3 1 2; do ( time env KMD_MERGE_LANES=$lanes ./kmdiff_b200/kmdiff-b200 diff --km-run $D -1 10 -2 10 -t 16 -o /dev/shm/cf_out -v debug ) > $O/cli_lanes${lanes}_$RANDOM.log 2>&1; sleep 1; done
rm -rf /dev/shm/cf_lz4 /dev/shm/cf_out*
