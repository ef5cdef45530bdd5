#!/bin/bash
# Round-2 GPU session I (8 GPUs): final validation -- send/recv + broadcast fixes under nccl-tests' data check, bench.py
# with its on-device verification, NVLS unroll.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
NCCL_TESTS_ITERS=8 NCCL_TESTS_WARMUP=3 timeout 600 bash scripts/run_nccl_tests.sh $N alltoall sendrecv broadcast all_reduce reduce_scatter > gpurun_out/i_nccl_tests.log 2>&1; echo "nccl_tests rc=$?" | tee gpurun_out/i_summary.txt
grep -E "^\| (65536|1048576|16777216|268435456|1073741824) |###" gpurun_out/nccl_tests_$N/table.md
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/i_bench$N.json 2> gpurun_out/i_bench$N.err; echo "bench rc=$?" | tee -a gpurun_out/i_summary.txt
cut -c1-300 gpurun_out/i_bench$N.json; tail -2 gpurun_out/i_bench$N.err
for u in 4 8; do
  UCCL_B200_NVLS_UNROLL=$u timeout 200 $TR benchmarks/allreduce_perf.py --coll allreduce --min-bytes 67108864 --max-bytes 1073741824 --factor 4 --iters 10 --out gpurun_out/i_ar_unroll$u.json > gpurun_out/i_ar_unroll$u.log 2>&1; echo "ar unroll $u rc=$?" | tee -a gpurun_out/i_summary.txt
  grep -E "^ *(268435456|1073741824) B" gpurun_out/i_ar_unroll$u.log | cut -c1-330
done
