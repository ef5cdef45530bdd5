#!/usr/bin/env bash
# Every benchmark of the repo on N GPUs, JSON results into OUT (default gpurun_out/).
set -uo pipefail
N=${1:-8}
OUT=${2:-gpurun_out}
cd "$(dirname "$0")/.."
mkdir -p "$OUT"
R="scripts/run_node.sh $N"
$R bench.py --gpus "$N" --steps 20 --warmup 5 | tee "$OUT/bench$N.json"
$R benchmarks/allreduce_perf.py --out "$OUT/ar$N.json"
for coll in allgather reduce_scatter alltoall; do
  $R benchmarks/allreduce_perf.py --coll $coll --out "$OUT/${coll}$N.json" || true
done
$R benchmarks/ep_sweep.py --ll --out "$OUT/ep$N.json"
$R benchmarks/ep_baseline.py --out "$OUT/ep_baseline$N.json" || true
python benchmarks/p2p_bench.py --out "$OUT/p2p.json" || true
python benchmarks/d2h_fifo_bench.py --out "$OUT/d2h.json" || true
$R benchmarks/uk_bench.py --out "$OUT/uk$N.json" || true
python benchmarks/compress_bench.py --out "$OUT/compress.json" || true
$R benchmarks/p2p_traffic.py --pattern permutation --out "$OUT/perm$N.json" || true
$R benchmarks/p2p_traffic.py --pattern incast --out "$OUT/incast$N.json" || true
python benchmarks/hostlink_bench.py --out "$OUT/hostlink.json" || true
$R benchmarks/sm_partition_bench.py --out "$OUT/sm_partition$N.json" || true
