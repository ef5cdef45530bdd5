#!/usr/bin/env bash
# nccl-tests against the NCCL-API drop-in (role of collective/rdma/run_nccl_test.sh and
# experimental/lite/scripts/run-nccl-tests.sh): same binaries, the shim is preloaded.
#   NCCL_TESTS=/path/to/nccl-tests/build examples/launchers/nccl_tests.sh all_reduce_perf 8
set -euo pipefail
REPO=$(cd "$(dirname "$0")/../.." && pwd)
BIN=${1:-all_reduce_perf}
N=${2:-8}
: "${NCCL_TESTS:?set NCCL_TESTS to the nccl-tests build directory}"
LD_PRELOAD="$REPO/uccl_b200/lib/libuccl_b200_nccl.so" "$NCCL_TESTS/$BIN" -b 1K -e 1G -f 2 -w 50 -n 50 -g "$N" -c 1
