#!/bin/bash
# `ncu --set full` captures of every hot kernel on ONE GPU (benchmarks/ncu_targets.py); reports land in gpurun_out/.
# The CSV summaries kept under profiles/ are produced afterwards, without a GPU, by scripts/ncu_summarise.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# expert-parallel kernels, with source correlation (compiled with -lineinfo)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ep_' --launch-count 40 \
  -o gpurun_out/ncu_ep -f python benchmarks/ncu_targets.py --reps 1 > gpurun_out/ncu_ep.log 2>&1
echo "ncu ep rc=$?"
# collectives, send/recv and the P2P copy engine
timeout 900 ncu --set full --clock-control none -k regex:'ar_|ag_|rs_|red_|a2a|xchg|bcast|sendrecv|p2p_copy' --launch-count 40 \
  -o gpurun_out/ncu_coll -f python benchmarks/ncu_targets.py --reps 1 > gpurun_out/ncu_coll.log 2>&1
echo "ncu coll rc=$?"
# every launch with its device time (cold caches, serialised: compare shares)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ub::' -c 400 --csv --log-file gpurun_out/ncu_launches.csv \
  python benchmarks/ncu_targets.py --reps 2 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep
