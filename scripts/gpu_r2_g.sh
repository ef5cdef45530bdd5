#!/bin/bash
# Round-2 GPU session G (4 GPUs): the N = 4 hole -- EP sweep, bench + anchor, nccl-tests through the drop-in.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
timeout 300 $TR benchmarks/ep_sweep.py --impls reg,tma --sms 24,48,96 --iters 10 --ll --out gpurun_out/g_ep$N.json > gpurun_out/g_ep$N.log 2>&1; echo "sweep rc=$?" | tee gpurun_out/g_summary.txt
grep -E '"sms": 24|"ll"' gpurun_out/g_ep$N.log | cut -c1-260
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g_bench$N.json 2> gpurun_out/g_bench$N.err; echo "bench rc=$?" | tee -a gpurun_out/g_summary.txt
cut -c1-300 gpurun_out/g_bench$N.json; tail -3 gpurun_out/g_bench$N.err
timeout 400 $TR bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/g_ref$N.json 2> gpurun_out/g_ref$N.err; echo "ref rc=$?" | tee -a gpurun_out/g_summary.txt
NCCL_TESTS_ITERS=8 NCCL_TESTS_WARMUP=3 timeout 600 bash scripts/run_nccl_tests.sh $N all_reduce all_gather reduce_scatter alltoall broadcast sendrecv > gpurun_out/g_nccl_tests.log 2>&1; echo "nccl_tests rc=$?" | tee -a gpurun_out/g_summary.txt
grep -E "^\| (65536|1048576|16777216|268435456|1073741824) |###" gpurun_out/nccl_tests_$N/table.md
