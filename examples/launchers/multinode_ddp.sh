#!/usr/bin/env bash
# DDP across boxes on the "uccl_b200" torch backend: NVLink kernels inside a box, datagram rails between.
#   NODE_RANK=0 NNODES=2 MASTER=10.0.0.1 ./examples/launchers/multinode_ddp.sh
# Single machine stand-in (2 "boxes" of 2 ranks, CPU backend works too):
#   UCCL_B200_LOCAL_SIZE=2 NNODES=1 NPROC=4 ./examples/launchers/multinode_ddp.sh --cpu
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
NNODES="${NNODES:-1}"; NODE_RANK="${NODE_RANK:-0}"; MASTER="${MASTER:-127.0.0.1}"; NPROC="${NPROC:-8}"
export UCCL_B200_NET_IFNAME="${UCCL_B200_NET_IFNAME:-}"      # e.g. "mlx5_,eth" to restrict the NICs
exec python -m torch.distributed.run --nnodes "$NNODES" --node-rank "$NODE_RANK" --nproc-per-node "$NPROC" \
  --master-addr "$MASTER" --master-port "${MASTER_PORT:-29500}" \
  "$ROOT/examples/ddp_train.py" --backend uccl_b200 "$@"
