#!/bin/bash
# Regenerates tests/golden/config1_run.tar.xz: BASELINE.json configs[0], the reference's own example
# (/root/reference/examples/run.sh:3 — `kmdiff count --file fof.txt --run-dir kc_dir --kmer-size 31 --hard-min 2`)
# run with the REFERENCE binaries: oracle/_ref/kmdiff (make -C oracle refbin) and oracle/_ref/kmtricks
# (make -C oracle kmtricks).  Only what `kmdiff diff` reads is kept: counts/partition_*/<sample>.kmer.lz4,
# histograms/*.hist, kmtricks.fof, options.txt, kmdiff-count.opt and an EMPTY matrices/ (count-file mode).
# Goldens of `kmdiff diff --km-run kc_dir -1 10 -2 10 -s 0.01` (run.sh:5), also in BASELINE.md §2:
#   59014/139876 significant k-mers; before correction 23985 (control), 35029 (case);
#   significant 23761 (control), 34837 (case);
#   md5 of the sorted sequence lines: case 85c677ff748c477b2571fab37a0d09a0, control 9437e5a6abb836fd38a777c805aa6e5d
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REF:-/root/reference}
export PATH="$HERE/../../oracle/_ref:$PATH"
W=$(mktemp -d)
cp -r "$REF/examples/data" "$REF/examples/fof.txt" "$W/"
(cd "$W" && kmdiff count --file fof.txt --run-dir kc_dir --kmer-size 31 --hard-min 2 -t 8)
mkdir -p "$W/min/matrices"
cp -r "$W/kc_dir/counts" "$W/kc_dir/histograms" "$W/kc_dir/options.txt" "$W/kc_dir/kmtricks.fof" \
      "$W/kc_dir/kmdiff-count.opt" "$W/min/"
(cd "$W/min" && tar -cJf "$HERE/config1_run.tar.xz" .)
(cd "$W" && kmdiff diff --km-run kc_dir -1 10 -2 10 --output-dir out -s 0.01 -t 8 2>&1 | grep -i "significant\|correction")
for f in case control; do grep -v '>' "$W/out/${f}_kmers.fasta" | sort | md5sum; done
rm -rf "$W"
