#!/bin/bash
# `ncu --set full` captures of every hot kernel on ONE GPU (benchmarks/ncu_targets.py).  gpurun brings back at most
# 64 MiB, so the raw metric pages are exported to CSV on the box and only the EP report (with source correlation) is kept.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# expert-parallel kernels, with source correlation (compiled with -lineinfo)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ep_' --launch-count 30 \
  -o gpurun_out/ncu_ep -f python benchmarks/ncu_targets.py --reps 1 > gpurun_out/ncu_ep.log 2>&1
echo "ncu ep rc=$?"
ncu -i gpurun_out/ncu_ep.ncu-rep --page raw --csv > gpurun_out/ncu_ep_raw.csv 2>/dev/null
# collectives, send/recv and the P2P copy engine (report too large to bring back: CSV only)
timeout 900 ncu --set full --clock-control none -k regex:'ar_|ag_|rs_|red_|a2a|xchg|bcast|sendrecv|p2p_copy' --launch-count 30 \
  -o /tmp/ncu_coll -f python benchmarks/ncu_targets.py --reps 1 > gpurun_out/ncu_coll.log 2>&1
echo "ncu coll rc=$?"
ncu -i /tmp/ncu_coll.ncu-rep --page raw --csv > gpurun_out/ncu_coll_raw.csv 2>/dev/null
# every launch with its device time (cold caches, serialised: compare shares)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ub::' -c 400 --csv --log-file gpurun_out/ncu_launches.csv \
  python benchmarks/ncu_targets.py --reps 2 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/ | head -20; du -sh gpurun_out
