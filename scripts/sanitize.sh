#!/usr/bin/env bash
# compute-sanitizer passes over a small GPU workload (the reference ships no sanitizer configuration,
# SURVEY 5.2).  memcheck + initcheck on the single-GPU virtual-rank tests; racecheck only sees shared
# memory hazards, which is what the LL dispatch / TMA copy kernels use.  Needs one GPU; slow (minutes).
#   scripts/sanitize.sh [memcheck|racecheck|initcheck|synccheck] [pytest -k expression]
set -euo pipefail
cd "$(dirname "$0")/.."
TOOL=${1:-memcheck}
EXPR=${2:-"allreduce_oneshot or test_dispatch_realistic or roundtrip_bit_exact"}
export UCCL_B200_TIMEOUT_MS=120000   # kernels run 10-100x slower under the sanitizer: no spurious peer timeouts
exec compute-sanitizer --tool "$TOOL" --target-processes all --error-exitcode 1 \
  python -m pytest tests -m gpu -x -q -k "$EXPR"
