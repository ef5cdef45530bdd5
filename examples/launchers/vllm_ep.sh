#!/usr/bin/env bash
# vLLM with expert parallelism over uccl_b200's DeepEP-compatible Buffer (role of the reference's
# ep/bench/vllm/launch_vllm.sh).  vLLM imports `deep_ep`; this repository provides that package.
set -euo pipefail
REPO=$(cd "$(dirname "$0")/../.." && pwd)
export PYTHONPATH="$REPO:${PYTHONPATH:-}"
export VLLM_ALL2ALL_BACKEND=${VLLM_ALL2ALL_BACKEND:-deepep_high_throughput}   # or deepep_low_latency
MODEL=${1:-deepseek-ai/DeepSeek-V3}
TP=${TP:-1}
DP=${DP:-8}
exec vllm serve "$MODEL" --tensor-parallel-size "$TP" --data-parallel-size "$DP" --enable-expert-parallel \
  --trust-remote-code "${@:2}"
