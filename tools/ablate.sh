#!/bin/bash
# Ablation of the fused kernel with the PROFILING build (results are invalid by design; throughput only):
#   bit 0 skips phase A, bit 1 skips phase B (the dense write), bit 2 skips the staging gather.
# usage: tools/ablate.sh "cfg2 S=1,comb-2 S=1" [units]
cd "$(dirname "$0")/.."
export SRSCE_LIBRARY=$PWD/srsran_ce_pytorch_b200/libsrsce_prof.so
for mode in 0 4 1 5 2 6; do
  echo "## SRSCE_DEBUG_ABLATE=$mode"
  SRSCE_DEBUG_ABLATE=$mode python tools/sweep_bench.py --only "$1" --units "${2:-8192}" 2>/dev/null | \
    python -c 'import sys,json
for l in sys.stdin:
    r=json.loads(l); print("   %-34s %8.3f ms  %9d units/s  frac %.3f" % (r["point"], r["ms"], r["units_per_s"], r["frac_of_measured_hbm_peak"]))'
done
