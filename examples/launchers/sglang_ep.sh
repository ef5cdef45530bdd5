#!/usr/bin/env bash
# SGLang with DeepEP MoE dispatch through uccl_b200 (role of ep/bench/sglang/common_launch.sh).
set -euo pipefail
REPO=$(cd "$(dirname "$0")/../.." && pwd)
export PYTHONPATH="$REPO:${PYTHONPATH:-}"
MODEL=${1:-deepseek-ai/DeepSeek-V3}
exec python -m sglang.launch_server --model-path "$MODEL" --tp 8 --ep-size 8 --enable-deepep-moe \
  --deepep-mode "${DEEPEP_MODE:-auto}" --trust-remote-code "${@:2}"
