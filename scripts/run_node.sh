#!/usr/bin/env bash
# One-node launcher (role of the reference's per-cluster run_*.sh / scripts/*): runs a python entry
# point on N GPUs with one process per GPU.
#   scripts/run_node.sh 8 bench.py --gpus 8
#   scripts/run_node.sh 8 benchmarks/allreduce_perf.py --max-bytes 1073741824
set -euo pipefail
N=${1:?number of GPUs}
shift
PORT=${MASTER_PORT:-$((20000 + RANDOM % 20000))}
cd "$(dirname "$0")/.."
if [ "$N" -eq 1 ]; then
  exec python "$@"
fi
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$PORT" "$@"
