#!/bin/bash
# N-way comparison of library builds inside one gpurun call:  tools/abn.sh "point filters" lib1.so lib2.so ...
cd "$(dirname "$0")/.."
P="$1"; shift
fmt='import sys,json
for l in sys.stdin:
    r=json.loads(l); print("  %-34s %8.3f ms frac %.3f" % (r["point"], r["ms"], r["frac_of_measured_hbm_peak"]))'
for round in 1 2; do
for lib in "$@"; do
  echo "## $lib"; SRSCE_LIBRARY=$PWD/srsran_ce_pytorch_b200/$lib timeout 120 python tools/sweep_bench.py --only "$P" 2>/dev/null | python -c "$fmt"
done
done
