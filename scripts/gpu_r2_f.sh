#!/bin/bash
# Round-2 GPU session F (8 GPUs): nccl-tests through the drop-in, DDP ResNet-50 (BASELINE configs #2 and #5),
# plain-buffer all-reduce variants, low-latency EP timeline, NVLS unroll.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
NCCL_TESTS_ITERS=10 NCCL_TESTS_WARMUP=3 timeout 900 bash scripts/run_nccl_tests.sh $N all_reduce all_gather reduce_scatter alltoall broadcast sendrecv > gpurun_out/f_nccl_tests.log 2>&1; echo "nccl_tests rc=$?" | tee gpurun_out/f_summary.txt
grep -E "^\| (65536|1048576|16777216|268435456|1073741824) |###" gpurun_out/nccl_tests_$N/table.md
rm -f gpurun_out/f_ddp$N.jsonl
for be in nccl uccl_b200 "uccl_b200 --sym-buckets" "hook --sym-buckets"; do
  timeout 300 $TR examples/ddp_train.py --backend $be --model resnet50 --batch 64 --steps 30 --warmup 8 --json gpurun_out/f_ddp$N.jsonl >> gpurun_out/f_ddp.log 2>&1; echo "ddp $be rc=$?" | tee -a gpurun_out/f_summary.txt
done
cat gpurun_out/f_ddp$N.jsonl
timeout 300 $TR benchmarks/ar_plain_bench.py --sizes 134217728,268435456,1073741824 --out gpurun_out/f_ar_plain$N.json > gpurun_out/f_ar_plain.log 2>&1; echo "ar_plain rc=$?" | tee -a gpurun_out/f_summary.txt
tail -4 gpurun_out/f_ar_plain.log | cut -c1-700
timeout 200 $TR benchmarks/ll_trace.py --out gpurun_out/f_ll_trace$N.json > gpurun_out/f_ll_trace.log 2>&1; echo "ll_trace rc=$?" | tee -a gpurun_out/f_summary.txt
timeout 300 $TR benchmarks/ep_sweep.py --impls tma --sms 24 --modes fp8_fused --iters 5 --ll --out gpurun_out/f_ll$N.json > gpurun_out/f_ll$N.log 2>&1; echo "ll rc=$?" | tee -a gpurun_out/f_summary.txt
grep '"ll"' gpurun_out/f_ll$N.log | cut -c1-300
for u in 4 8; do
  UCCL_B200_NVLS_UNROLL=$u timeout 200 $TR benchmarks/allreduce_perf.py --coll allreduce --min 16777216 --max 1073741824 --factor 4 --iters 10 --out gpurun_out/f_ar_unroll$u.json > gpurun_out/f_ar_unroll$u.log 2>&1; echo "ar unroll $u rc=$?" | tee -a gpurun_out/f_summary.txt
  grep -E "^ *(268435456|1073741824) B" gpurun_out/f_ar_unroll$u.log | cut -c1-400
done
