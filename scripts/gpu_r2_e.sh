#!/bin/bash
# Round-2 GPU session E (8 GPUs): the headline at the reference's SM budget + sweep, same-box DeepEP anchor,
# nccl-tests through the drop-in, plain-buffer all-reduce variants, DDP ResNet-50.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/e_gpus.txt 2>&1
timeout 400 $TR benchmarks/ep_sweep.py --impls reg,tma --sms 16,24,32,48,64,96 --iters 10 --ll --out gpurun_out/e_ep$N.json > gpurun_out/e_ep$N.log 2>&1; echo "sweep rc=$?" | tee gpurun_out/e_summary.txt
grep -E '"sms": (24|48|96)|"ll"' gpurun_out/e_ep$N.log | cut -c1-260
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/e_bench$N.json 2> gpurun_out/e_bench$N.err; echo "bench rc=$?" | tee -a gpurun_out/e_summary.txt
cut -c1-400 gpurun_out/e_bench$N.json
timeout 400 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/e_ref$N.json 2> gpurun_out/e_ref$N.err; echo "ref rc=$?" | tee -a gpurun_out/e_summary.txt
cut -c1-1200 gpurun_out/e_ref$N.json
timeout 300 $TR benchmarks/ar_plain_bench.py --out gpurun_out/e_ar_plain$N.json > gpurun_out/e_ar_plain.log 2>&1; echo "ar_plain rc=$?" | tee -a gpurun_out/e_summary.txt
tail -6 gpurun_out/e_ar_plain.log | cut -c1-600
