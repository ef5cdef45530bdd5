#!/bin/bash
# Round-2 GPU session A (1 GPU): smoke, the GPU test-suite, EP N=1 sweep (register path vs TMA), bench, launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/a_summary.txt
timeout 900 python -m pytest tests/test_gpu_ep.py -q -x --timeout 300 > gpurun_out/a_tests_ep.log 2>&1; echo "tests_ep rc=$?" | tee -a gpurun_out/a_summary.txt
tail -5 gpurun_out/a_tests_ep.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_gpu_ep.py > gpurun_out/a_tests_all.log 2>&1; echo "tests_all rc=$?" | tee -a gpurun_out/a_summary.txt
tail -5 gpurun_out/a_tests_all.log
timeout 300 python benchmarks/ep_sweep.py --impls reg,tma --sms 24,64,148 --iters 10 --out gpurun_out/a_ep1.json > gpurun_out/a_ep1.log 2>&1; echo "sweep rc=$?" | tee -a gpurun_out/a_summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench1.json 2> gpurun_out/a_bench1.err; echo "bench rc=$?" | tee -a gpurun_out/a_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/a_smoke_launches.csv python __graft_entry__.py smoke > gpurun_out/a_smoke_ncu.log 2>&1; echo "smoke_ncu rc=$?" | tee -a gpurun_out/a_summary.txt
cat gpurun_out/a_bench1.json | cut -c1-600
