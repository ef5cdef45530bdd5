#!/bin/bash
# First GPU call of the next round (1 GPU, ~25 min): everything that was written after round 2's GPU budget was spent,
# in the order ROADMAP.md lists it.  Logs go to gpurun_out/; copy what should be judged into profiles/ afterwards
# (python scripts/ncu_summarise.py does that for the ncu CSVs).
#
#   gpurun --timeout 1800 -- bash scripts/validate_first.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/vf_clocks.csv 2>&1
# 1. the whole GPU suite: the tests of the unvalidated pieces sort last (test_zzz_*, test_zzzz_*)
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/vf_gpu_tests.log
echo "gpu tests rc=$?"; tail -3 gpurun_out/vf_gpu_tests.log
# 2. smoke + headline bench (pick_impl / refill-rule changes are on this path)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/vf_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/vf_bench1.json 2> gpurun_out/vf_bench1.err; echo "bench rc=$?"
# 3. new benchmarks
timeout 300 python benchmarks/sm_partition_bench.py --out gpurun_out/vf_sm_partition1.json > gpurun_out/vf_sm_partition1.log 2>&1; echo "sm_partition rc=$?"
timeout 300 python benchmarks/hostlink_bench.py --out gpurun_out/vf_hostlink.json > gpurun_out/vf_hostlink.log 2>&1; echo "hostlink rc=$?"
timeout 300 python benchmarks/p2p_bench.py --out gpurun_out/vf_p2p.json > gpurun_out/vf_p2p.log 2>&1; echo "p2p rc=$?"
# 4. the ncu captures that round 2 lost to the 64 MiB return cap (CSV export happens on the box)
bash scripts/ncu_capture.sh
