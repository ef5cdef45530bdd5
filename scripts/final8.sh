#!/usr/bin/env bash
# The round's 8-GPU measurement pass (bounded: every step has its own timeout).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
t0=$(date +%s)
step() { echo "== $1 (t+$(( $(date +%s) - t0 ))s)"; }
step "pytest: LL exchange + RS push + NVLS on 8 real GPUs"
timeout 100 python -m pytest tests/test_gpu_collectives.py -m gpu -q -x -k "ll_exchange or push_staging or nvls" > $O/tests8.log 2>&1; tail -3 $O/tests8.log
step "bench.py --gpus 8"
timeout 110 $T --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench8.json 2> $O/bench8.err; tail -c 300 $O/bench8.json; tail -2 $O/bench8.err
step "allreduce sweep"
timeout 100 $T --master-port 29602 benchmarks/allreduce_perf.py --iters 10 --out $O/ar8.json > $O/ar8.log 2>&1; tail -3 $O/ar8.log
step "allgather sweep"
timeout 60 $T --master-port 29603 benchmarks/allreduce_perf.py --coll allgather --factor 16 --iters 10 --out $O/allgather8.json > $O/ag8.log 2>&1; tail -2 $O/ag8.log
step "reduce_scatter sweep (push staging, default)"
timeout 60 $T --master-port 29604 benchmarks/allreduce_perf.py --coll reduce_scatter --factor 16 --iters 10 --out $O/reduce_scatter8.json > $O/rs8.log 2>&1; tail -2 $O/rs8.log
step "reduce_scatter sweep (pull staging)"
UCCL_B200_RS_PUSH=0 timeout 60 $T --master-port 29605 benchmarks/allreduce_perf.py --coll reduce_scatter --min 4194304 --factor 16 --iters 10 --out $O/reduce_scatter8_pull.json > $O/rs8p.log 2>&1; tail -2 $O/rs8p.log
step "EP sweep"
timeout 80 $T --master-port 29606 benchmarks/ep_sweep.py --sms 64,96 --ll --iters 10 --out $O/ep8.json > $O/ep8.log 2>&1; tail -3 $O/ep8.log
step "alltoall sweep"
timeout 60 $T --master-port 29607 benchmarks/allreduce_perf.py --coll alltoall --factor 16 --iters 10 --out $O/alltoall8.json > $O/a2a8.log 2>&1; tail -2 $O/a2a8.log
step "done"
