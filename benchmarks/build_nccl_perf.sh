#!/bin/bash
# Builds the nccl-tests style harness twice: against this repo's NCCL drop-in and (if present)
# against the system / torch-bundled NCCL as the baseline.
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(dirname "$here")
out=${1:-$root/build}
mkdir -p "$out"
python -c "import sys; sys.path.insert(0, '$root'); from uccl_b200 import _build; _build.build()"
g++ -std=c++17 -O2 "$here/nccl_perf.cc" -I/usr/include -I/usr/local/cuda/include \
    -L"$root/uccl_b200/lib" -luccl_b200_nccl -Wl,-rpath,"$root/uccl_b200/lib" \
    -L/usr/local/cuda/lib64 -lcudart -lpthread -o "$out/nccl_perf_uccl_b200"
# one process per rank (ncclCommInitRank): also the way to exercise the multi-box path (UCCL_B200_LOCAL_SIZE)
g++ -std=c++17 -O2 "$here/nccl_perf_mp.cc" -I/usr/include -I/usr/local/cuda/include \
    -L"$root/uccl_b200/lib" -luccl_b200_nccl -Wl,-rpath,"$root/uccl_b200/lib" \
    -L/usr/local/cuda/lib64 -lcudart -lpthread -o "$out/nccl_perf_mp_uccl_b200"
nccl_lib=$(python - <<'PY'
import glob, os, sys
cands = glob.glob(os.path.join(sys.prefix, "lib/python*/site-packages/nvidia/nccl/lib/libnccl.so.2")) + glob.glob("/usr/lib/x86_64-linux-gnu/libnccl.so.2")
print(cands[0] if cands else "")
PY
)
if [ -n "$nccl_lib" ]; then
  g++ -std=c++17 -O2 "$here/nccl_perf.cc" -I/usr/include -I/usr/local/cuda/include \
      "$nccl_lib" -Wl,-rpath,"$(dirname "$nccl_lib")" -L/usr/local/cuda/lib64 -lcudart -lpthread -o "$out/nccl_perf_nccl"
fi
ls -la "$out"/nccl_perf_*
