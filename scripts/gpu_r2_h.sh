#!/bin/bash
# Round-2 GPU session H (1 GPU): full GPU test-suite, smoke (+ under ncu), N=1 bench + sweep, ncu --set full captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/h_summary.txt
tail -4 gpurun_out/h_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/h_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/h_smoke_launches.csv python __graft_entry__.py smoke > gpurun_out/h_smoke_ncu.log 2>&1; echo "smoke_ncu rc=$?" | tee -a gpurun_out/h_summary.txt
timeout 300 python benchmarks/ep_sweep.py --impls reg,tma --sms 24,148 --iters 10 --out gpurun_out/h_ep1.json > gpurun_out/h_ep1.log 2>&1; echo "sweep rc=$?" | tee -a gpurun_out/h_summary.txt
cut -c1-220 gpurun_out/h_ep1.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/h_bench1.json 2> gpurun_out/h_bench1.err; echo "bench rc=$?" | tee -a gpurun_out/h_summary.txt
cut -c1-300 gpurun_out/h_bench1.json
bash scripts/ncu_capture.sh 2>&1 | tail -6 | tee -a gpurun_out/h_summary.txt
