#!/bin/bash
# Round-2 GPU session D (2 GPUs): reworked send/recv + P2P engine: tests, nccl-tests (alltoall / sendrecv), p2p sweep.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_gpu_ep.py > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/d_summary.txt
tail -6 gpurun_out/d_tests.log
NCCL_TESTS_MAX=256M NCCL_TESTS_ITERS=10 timeout 400 bash scripts/run_nccl_tests.sh $N all_reduce alltoall sendrecv > gpurun_out/d_nccl_tests.log 2>&1; echo "nccl_tests rc=$?" | tee -a gpurun_out/d_summary.txt
cat gpurun_out/nccl_tests_$N/table.md | grep -E "^\| (1048576|2097152|16777216|268435456) |###"
timeout 400 python benchmarks/p2p_bench.py --sizes 131072,1048576,8388608,134217728,536870912 --out gpurun_out/d_p2p2.json > gpurun_out/d_p2p.log 2>&1; echo "p2p rc=$?" | tee -a gpurun_out/d_summary.txt
cat gpurun_out/d_p2p.log | tail -20
