#!/bin/bash
# Round-2 GPU session F (8 GPUs): nccl-tests through the drop-in and DDP ResNet-50 (BASELINE configs #2 and #5).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
NCCL_TESTS_ITERS=10 NCCL_TESTS_WARMUP=3 timeout 900 bash scripts/run_nccl_tests.sh $N all_reduce all_gather reduce_scatter alltoall broadcast sendrecv > gpurun_out/f_nccl_tests.log 2>&1; echo "nccl_tests rc=$?" | tee gpurun_out/f_summary.txt
grep -E "^\| (1048576|16777216|268435456|1073741824) |###" gpurun_out/nccl_tests_$N/table.md
rm -f gpurun_out/f_ddp$N.jsonl
for be in nccl uccl_b200 "uccl_b200 --sym-buckets" hook "hook --sym-buckets"; do
  timeout 300 $TR examples/ddp_train.py --backend $be --model resnet50 --batch 64 --steps 30 --warmup 8 --json gpurun_out/f_ddp$N.jsonl >> gpurun_out/f_ddp.log 2>&1; echo "ddp $be rc=$?" | tee -a gpurun_out/f_summary.txt
done
cat gpurun_out/f_ddp$N.jsonl
