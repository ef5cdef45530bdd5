#!/bin/bash
# nccl-tests (built by scripts/build_nccl_tests.sh) on N GPUs of this box: system NCCL vs the uccl_b200 NCCL-API
# drop-in preloaded over the SAME binaries, correctness checking on (-c 1).  One process, N GPUs (-g N), like
# experimental/lite/scripts/run-nccl-tests.sh.  Output: gpurun_out/nccl_tests_<N>/{nccl,uccl_b200}_<test>.txt
#   scripts/run_nccl_tests.sh 8 [tests...]
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
N="${1:-8}"; shift || true
TESTS=("$@"); [ ${#TESTS[@]} -eq 0 ] && TESTS=(all_reduce all_gather reduce_scatter alltoall broadcast sendrecv)
BIN="$ROOT/build/nccl-tests"
OUT="$ROOT/gpurun_out/nccl_tests_$N"; mkdir -p "$OUT"
SHIM="$ROOT/uccl_b200/lib/libuccl_b200_nccl.so"
export UCCL_B200_NCCL_HEAP_MB="${UCCL_B200_NCCL_HEAP_MB:-4608}"   # send + recv + expected buffers of 1 GiB each come from ncclMemAlloc
ARGS="-b ${NCCL_TESTS_MIN:-1K} -e ${NCCL_TESTS_MAX:-1G} -f 2 -g $N -c 1 -w ${NCCL_TESTS_WARMUP:-5} -n ${NCCL_TESTS_ITERS:-20}"
for t in "${TESTS[@]}"; do
  echo "== $t: system NCCL"
  timeout 300 "$BIN/${t}_perf" $ARGS > "$OUT/nccl_$t.txt" 2>&1; echo "rc=$?" >> "$OUT/nccl_$t.txt"
  tail -4 "$OUT/nccl_$t.txt"
  echo "== $t: uccl_b200 drop-in (LD_PRELOAD)"
  LD_PRELOAD="$SHIM" timeout 300 "$BIN/${t}_perf" $ARGS > "$OUT/uccl_b200_$t.txt" 2>&1; echo "rc=$?" >> "$OUT/uccl_b200_$t.txt"
  tail -4 "$OUT/uccl_b200_$t.txt"
done
python "$ROOT/scripts/nccl_tests_table.py" "$OUT" > "$OUT/table.md" 2>/dev/null || true
