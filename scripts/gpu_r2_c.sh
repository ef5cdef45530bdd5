#!/bin/bash
# Round-2 GPU session C (2 GPUs): validate the reworked kernels + the new client paths before spending 8-GPU time.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 900 python -m pytest tests/test_gpu_ep.py tests/test_zz_gpu_send_recv_api.py -q --timeout 300 > gpurun_out/c_tests_ep.log 2>&1; echo "tests_ep rc=$?" | tee gpurun_out/c_summary.txt
tail -6 gpurun_out/c_tests_ep.log
timeout 400 $TR benchmarks/ep_sweep.py --impls reg,tma --sms 24,48,96 --iters 10 --ll --out gpurun_out/c_ep$N.json > gpurun_out/c_ep$N.log 2>&1; echo "sweep rc=$?" | tee -a gpurun_out/c_summary.txt
grep -E '"tma".*fp8|"ll"' gpurun_out/c_ep$N.log | cut -c1-200
timeout 400 $TR bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/c_ref$N.json 2> gpurun_out/c_ref$N.err; echo "ref rc=$?" | tee -a gpurun_out/c_summary.txt
cut -c1-900 gpurun_out/c_ref$N.json
NCCL_TESTS_MAX=64M NCCL_TESTS_ITERS=5 timeout 300 bash scripts/run_nccl_tests.sh $N all_reduce alltoall > gpurun_out/c_nccl_tests.log 2>&1; echo "nccl_tests rc=$?" | tee -a gpurun_out/c_summary.txt
tail -12 gpurun_out/c_nccl_tests.log
for be in nccl uccl_b200 "uccl_b200 --sym-buckets" hook; do
  timeout 200 $TR examples/ddp_train.py --backend $be --model resnet18 --steps 10 --warmup 3 --json gpurun_out/c_ddp$N.jsonl >> gpurun_out/c_ddp.log 2>&1; echo "ddp $be rc=$?" | tee -a gpurun_out/c_summary.txt
done
cat gpurun_out/c_ddp$N.jsonl
