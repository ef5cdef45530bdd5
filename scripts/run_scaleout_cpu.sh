#!/usr/bin/env bash
# Reproduces the CPU-only measurements of the scale-out stack into profiles/ (loopback UDP, no GPU, no NIC):
#   transport sweep (clean / 1 % loss / MTU-1500 datagrams), hierarchical vs flat all-reduce with 4 processes
#   standing in for 2 boxes x 2 ranks (Python path), and the same topology through the NCCL-API drop-in (C++ path).
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
cd "$ROOT"
OUT=profiles
python -c "from uccl_b200 import _build; _build.build()"
python benchmarks/net_bench.py --iters 50 | tail -1 > $OUT/net_loopback_cpu.json
python benchmarks/net_bench.py --iters 50 --drop 0.01 --max-bytes 16777216 | tail -1 > $OUT/net_loopback_cpu_loss1pct.json
python benchmarks/net_bench.py --iters 30 --payload 1400 --max-bytes 16777216 | tail -1 > $OUT/net_loopback_cpu_mtu1500.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29761 \
  benchmarks/multinode_allreduce.py --local-size 2 --bind 127.0.0.1 --max-bytes 67108864 --iters 5 --flat 2>/dev/null \
  | tail -1 > $OUT/multinode_allreduce_cpu_emulated.json
mkdir -p build
g++ -std=c++17 -O2 benchmarks/nccl_perf_mp.cc -I/usr/include -I/usr/local/cuda/include -Luccl_b200/lib -luccl_b200_nccl \
  -Wl,-rpath,"$ROOT/uccl_b200/lib" -L/usr/local/cuda/lib64 -lcudart -lpthread -o build/nccl_perf_mp_uccl_b200
{
  echo "# NCCL API (libuccl_b200_nccl.so, host backend), 4 processes; hierarchical = UCCL_B200_LOCAL_SIZE=2 (2 boxes x 2)"
  for op in allreduce allgather reducescatter alltoall; do
    echo "## $op, 2 boxes x 2 ranks (MultiComm)"
    UCCL_B200_LOCAL_SIZE=2 UCCL_B200_NET_BIND_IP=127.0.0.1 build/nccl_perf_mp_uccl_b200 -n 4 -o $op -b 4K -e 16M -f 4 -i 5 -w 2 | grep -v "^#  " 
  done
  echo "## allreduce, one box of 4 ranks (shared-memory host backend; no network)"
  build/nccl_perf_mp_uccl_b200 -n 4 -o allreduce -b 4K -e 16M -f 4 -i 5 -w 2 | grep -v "^#  "
} > $OUT/nccl_api_multibox_cpu_emulated.txt 2>&1
echo "wrote $OUT/net_loopback_cpu*.json $OUT/multinode_allreduce_cpu_emulated.json $OUT/nccl_api_multibox_cpu_emulated.txt"
