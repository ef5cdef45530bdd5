#!/usr/bin/env bash
# Host-side sanitizer build of the native core: every .cc of csrc/ compiled with -fsanitize=<san> by g++, the
# (uninstrumented, nvcc-built) kernel objects of build/obj reused, linked into build/san/libuccl_b200_nccl_<san>.so;
# then the C++ NCCL-API tests (host backend, 2 processes and 2 boxes x 2 ranks) run against it.  No GPU needed.
#   scripts/sanitize_host.sh [thread|address]
set -euo pipefail
cd "$(dirname "$0")/.."
SAN=${1:-thread}
FLAG=$SAN; [ "$SAN" = address ] && FLAG=address,undefined
python -c "import __graft_entry__ as g; g.build()" >/dev/null
OUT=build/san/$SAN; mkdir -p "$OUT"
PYINC=$(python -c "import sysconfig; print(sysconfig.get_paths()['include'])")
objs=()
for src in uccl_b200/csrc/{fabric,coll,ep,p2p,common,ukernel,net}/*.cc; do
  case "$src" in *bind_*|*nccl_net_plugin*) continue;; esac
  o="$OUT/$(basename "$(dirname "$src")")_$(basename "${src%.cc}").o"
  objs+=("$o")
  if [ ! -f "$o" ] || [ "$src" -nt "$o" ]; then
    g++ -std=c++17 -O1 -g -fPIC -fsanitize=$FLAG -Iuccl_b200/csrc -I/usr/local/cuda/include -I"$PYINC" -c "$src" -o "$o" &
  fi
done
wait
cu_objs=$(ls build/obj/*.o | while read -r o; do b=$(basename "$o" .o); ls uccl_b200/csrc/*/"${b#*_}".cu >/dev/null 2>&1 && echo "$o"; done)
LIB=$OUT/libuccl_b200_nccl_$SAN.so
g++ -shared -fsanitize=$FLAG -o "$LIB" "${objs[@]}" $cu_objs -L/usr/local/cuda/lib64 -lcudart -lrt -lpthread -ldl
export ASAN_OPTIONS=detect_leaks=0 TSAN_OPTIONS="halt_on_error=0 die_after_fork=0 second_deadlock_stack=1"
RC=0
run() {  # name, args...: prints the sanitizer findings (if any) and the test's verdict
  local name=$1; shift
  echo "== $name under $SAN"
  "$OUT/$name" "$@" > "$OUT/$name.log" 2>&1 || RC=1
  grep -E "^WARNING: ThreadSanitizer|ERROR: AddressSanitizer|runtime error:|SUMMARY|: OK|FAILED" "$OUT/$name.log" | sort | uniq -c | sort -rn | head -20 || true
  # (the API test passes the invalid enum value 99 on purpose: UBSan's "not a valid value for type ncclDataType_t" is that)
  if grep -E "^WARNING: ThreadSanitizer|ERROR: AddressSanitizer|runtime error:" "$OUT/$name.log" | grep -qv "ncclDataType_t"; then RC=1; fi
}
LINK=(-I/usr/include -Iuccl_b200/csrc -I/usr/local/cuda/include -L"$OUT" -luccl_b200_nccl_$SAN -Wl,-rpath,"$PWD/$OUT"
      -Wl,-rpath,/usr/local/cuda/lib64 -lpthread)
for t in nccl_api_test nccl_multibox_test host_world_stress uk_net_stress; do
  g++ -std=c++17 -O1 -g -fsanitize=$FLAG tests/cpp/$t.cc "${LINK[@]}" -o "$OUT/$t"
done
run nccl_api_test          # 2 processes, every NCCL entry point on the host backend
run nccl_multibox_test     # 2 boxes x 2 ranks: MultiComm over the datagram transport
run host_world_stress 40   # 4 ranks as threads: native collectives + ukernel worker threads
run uk_net_stress 30       # 3 ranks as threads: ukernel plans over the datagram transport (receiver threads)
[ $RC = 0 ] && echo "sanitize_host($SAN): clean" || { echo "sanitize_host($SAN): FINDINGS (logs in $OUT)"; exit 1; }
