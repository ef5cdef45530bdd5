#!/usr/bin/env bash
# The Python CPU test-suite against an AddressSanitizer build of the extension module: a scratch clone of the
# repository is built normally (kernel objects), every .cc -- bindings included -- is recompiled with
# -fsanitize=address and linked over the clone's _C module, and pytest runs there with libasan (and libstdc++, so that
# ASan can intercept __cxa_throw inside an interpreter that is not linked against it) preloaded.  No GPU needed.
#   scripts/sanitize_python.sh [scratch dir] [pytest args...]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SCRATCH=${1:-/tmp/uccl_b200_asan_py}; shift || true
rm -rf "$SCRATCH"; git clone -q "$ROOT" "$SCRATCH"; cd "$SCRATCH"
git -C "$ROOT" diff HEAD | git apply --allow-empty 2>/dev/null || true   # uncommitted work of the source tree
python -c "import __graft_entry__ as g; g.build()" | tail -1
OUT=build/san_py; mkdir -p $OUT
PYINC=$(python -c "import sysconfig; print(sysconfig.get_paths()['include'])")
PB=$(python -c "import pybind11; print(pybind11.get_include())")
objs=()
for src in uccl_b200/csrc/{fabric,coll,ep,p2p,common,ukernel,net,bind}/*.cc; do
  case "$src" in *nccl_net_plugin*|*nccl_shim*) continue;; esac
  o="$OUT/$(basename "$(dirname "$src")")_$(basename "${src%.cc}").o"; objs+=("$o")
  g++ -std=c++17 -O1 -g -fPIC -fvisibility=hidden -w -fsanitize=address -Iuccl_b200/csrc -I/usr/local/cuda/include \
      -I"$PYINC" -I"$PB" -c "$src" -o "$o" &
done
wait
cu_objs=$(ls build/obj/*.o | while read -r o; do b=$(basename "$o" .o); ls uccl_b200/csrc/*/"${b#*_}".cu >/dev/null 2>&1 && echo "$o"; done)
MOD=$(python -c "from uccl_b200 import _build; print(_build.module_path())")
g++ -shared -fsanitize=address -o "$MOD" "${objs[@]}" $cu_objs -L/usr/local/cuda/lib64 -lcudart -lrt -lpthread -ldl
LOG=$SCRATCH/asan_report
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)"
export ASAN_OPTIONS=detect_leaks=0:alloc_dealloc_mismatch=0:detect_odr_violation=0:verify_asan_link_order=0:halt_on_error=0:log_path=$LOG
# tests that compile and link their own C++ against the (uninstrumented) libraries, or shell out to tools, are left out
python -m pytest tests -q -m "not gpu" -p no:cacheprovider --deselect tests/test_host_p2p.py::test_uccl_engine_c_api_host_mode \
  --ignore tests/test_nccl_shim.py -k "not sanitizers and not tsan and not info_cli and not perf_gate" "$@" || true
unset LD_PRELOAD
if ls "$LOG".* >/dev/null 2>&1; then echo "sanitize_python: ASan REPORTS in $LOG.*"; head -30 "$LOG".* | cut -c1-200; exit 1; fi
echo "sanitize_python: clean"
