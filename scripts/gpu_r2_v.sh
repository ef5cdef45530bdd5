#!/bin/bash
# quick 2-GPU validation of the send/recv + P2P changes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_send_recv_api.py tests/test_gpu_p2p.py tests/test_nccl_shim.py tests/test_gpu_collectives.py -m gpu -q --timeout 300 > gpurun_out/v_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/v_tests.log
NCCL_TESTS_MAX=256M NCCL_TESTS_ITERS=10 timeout 400 bash scripts/run_nccl_tests.sh 2 alltoall sendrecv > gpurun_out/v_nccl_tests.log 2>&1; echo "nccl_tests rc=$?"
grep -E "^\| (65536|1048576|2097152|16777216|268435456) |###" gpurun_out/nccl_tests_2/table.md
timeout 400 python benchmarks/p2p_bench.py --sizes 1048576,8388608,134217728,536870912 --out gpurun_out/v_p2p2.json > gpurun_out/v_p2p.log 2>&1; echo "p2p rc=$?"
cut -c1-260 gpurun_out/v_p2p.log | tail -14
