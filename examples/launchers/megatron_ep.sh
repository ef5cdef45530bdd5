#!/usr/bin/env bash
# Megatron-LM MoE pre-training with the flex token dispatcher on DeepEP, DeepEP being this repository's `deep_ep`
# package (role of ep/bench/megatron/deepseekv3_uep_slurm_pretrain.sh; its NCCL twin differs only in the dispatcher
# flags).  One node, 8 GPUs, a DeepSeek-V3 style MoE cut down to NUM_LAYERS layers.  Not run in this repository's
# environment (no Megatron-LM checkout): a template that records the flags that matter.
#
#   MEGATRON=/path/to/Megatron-LM examples/launchers/megatron_ep.sh [extra pretrain_gpt.py arguments]
#   DISPATCHER=alltoall ...     # Megatron's NCCL all-to-all dispatcher instead, for an A/B run
set -euo pipefail
REPO=$(cd "$(dirname "$0")/../.." && pwd)
: "${MEGATRON:?set MEGATRON to a Megatron-LM checkout (core_r0.12 or newer: flex dispatcher)}"
export PYTHONPATH="$REPO:$MEGATRON:${PYTHONPATH:-}"   # our deep_ep first: it must shadow an upstream install
export CUDA_DEVICE_MAX_CONNECTIONS=${CUDA_DEVICE_MAX_CONNECTIONS:-32}
NPROC=${NPROC:-8}
NUM_LAYERS=${NUM_LAYERS:-12}
DISPATCHER=${DISPATCHER:-flex}
if [ "$DISPATCHER" = flex ]; then
  DISPATCH_ARGS=(--moe-token-dispatcher-type flex --moe-enable-deepep)
else
  DISPATCH_ARGS=(--moe-token-dispatcher-type alltoall)
fi
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NPROC" --master-addr 127.0.0.1 \
  --master-port "${MASTER_PORT:-29517}" "$MEGATRON/pretrain_gpt.py" \
  --tensor-model-parallel-size 1 --pipeline-model-parallel-size 1 --expert-model-parallel-size "$NPROC" \
  --num-layers "$NUM_LAYERS" --hidden-size 7168 --ffn-hidden-size 18432 --num-attention-heads 128 \
  --seq-length 4096 --max-position-embeddings 4096 --micro-batch-size 1 --global-batch-size "$((NPROC * 8))" \
  --num-experts 256 --moe-router-topk 8 --moe-ffn-hidden-size 2048 --moe-router-force-load-balancing \
  --moe-grouped-gemm --moe-permute-fusion "${DISPATCH_ARGS[@]}" \
  --bf16 --mock-data --tokenizer-type NullTokenizer --vocab-size 129280 \
  --train-iters "${TRAIN_ITERS:-30}" --lr 1e-4 --log-interval 1 --log-throughput --no-save-optim --no-save-rng \
  "$@"
