#!/bin/bash
# A/B of two library builds inside one gpurun call (box-to-box variance is larger than most effects):
#   tools/ab.sh "point filters" [lib_a] [lib_b] [rounds]
cd "$(dirname "$0")/.."
A=${2:-$PWD/srsran_ce_pytorch_b200/libsrsce.so}
B=${3:-$PWD/srsran_ce_pytorch_b200/libsrsce_old.so}
fmt='import sys,json
for l in sys.stdin:
    r=json.loads(l); print("  %-34s %8.3f ms frac %.3f" % (r["point"], r["ms"], r["frac_of_measured_hbm_peak"]))'
for i in $(seq 1 ${4:-2}); do
  echo "## A $(basename $A)"; SRSCE_LIBRARY=$A python tools/sweep_bench.py --only "$1" 2>/dev/null | python -c "$fmt"
  echo "## B $(basename $B)"; SRSCE_LIBRARY=$B python tools/sweep_bench.py --only "$1" 2>/dev/null | python -c "$fmt"
done
