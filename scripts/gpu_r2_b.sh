#!/bin/bash
# Round-2 GPU session B (2 GPUs): EP tests (new TMA / LL paths), EP sweep reg vs tma, LL timeline, bench + reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_gpu_ep.py tests/test_zz_gpu_send_recv_api.py -q --timeout 300 > gpurun_out/b_tests_ep.log 2>&1; echo "tests_ep rc=$?" | tee gpurun_out/b_summary.txt
tail -8 gpurun_out/b_tests_ep.log
timeout 600 $TR benchmarks/ep_sweep.py --impls reg,tma --sms 16,24,32,48,64,96 --iters 10 --ll --out gpurun_out/b_ep2.json > gpurun_out/b_ep2.log 2>&1; echo "sweep rc=$?" | tee -a gpurun_out/b_summary.txt
timeout 200 $TR benchmarks/ll_trace.py --out gpurun_out/b_ll_trace2.json > gpurun_out/b_ll_trace.log 2>&1; echo "ll_trace rc=$?" | tee -a gpurun_out/b_summary.txt
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b_bench2.json 2> gpurun_out/b_bench2.err; echo "bench rc=$?" | tee -a gpurun_out/b_summary.txt
timeout 400 $TR bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > gpurun_out/b_ref2.json 2> gpurun_out/b_ref2.err; echo "ref rc=$?" | tee -a gpurun_out/b_summary.txt
cut -c1-1500 gpurun_out/b_bench2.json; cut -c1-1500 gpurun_out/b_ref2.json
