#!/usr/bin/env bash
# First GPU call of the next round, in one go (run through gpurun; everything lands in gpurun_out/):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_first.sh'                 # 1 GPU
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_round2_first.sh 8'      # 8 GPUs (adds the multi-box emulations)
# Every step has its own timeout so that one hang cannot eat the call.
set -uo pipefail
N="${1:-1}"
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
# 1. the verified suite, then the hardware-pending tests (non-strict xfail: read XPASS / XFAIL in the summary)
timeout 1200 python -m pytest tests -m gpu -x -q --ignore=tests/test_zz_gpu_send_recv_api.py > $OUT/pytest_gpu.log 2>&1
timeout 900 python -m pytest tests/test_zz_gpu_send_recv_api.py -m gpu -q -rxXs > $OUT/pytest_pending.log 2>&1
tail -5 $OUT/pytest_gpu.log; tail -15 $OUT/pytest_pending.log
# 2. smoke + headline bench
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1
if [ "$N" -gt 1 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus "$N" --steps 50 --warmup 5 > $OUT/bench$N.json 2> $OUT/bench$N.err
else
  timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 > $OUT/bench1.json 2> $OUT/bench1.err
fi
# 3. scale-out stack on the box's CPUs / GPUs (loopback rails)
timeout 300 python benchmarks/net_bench.py --iters 30 > $OUT/net_bench.log 2>&1
timeout 300 env UCCL_B200_NET_SPIN_US=50 python benchmarks/net_bench.py --iters 30 > $OUT/net_bench_spin50.log 2>&1
if [ "$N" -ge 4 ]; then
  bash benchmarks/build_nccl_perf.sh build > $OUT/build_nccl_perf.log 2>&1
  L=$((N / 2))
  for pb in 8388608 1000000000; do
    timeout 600 env UCCL_B200_LOCAL_SIZE=$L UCCL_B200_NET_BIND_IP=127.0.0.1 UCCL_B200_MN_PIPELINE_BYTES=$pb \
      build/nccl_perf_mp_uccl_b200 -n "$N" -g "$N" -o allreduce -b 64K -e 1G -f 4 -i 10 -w 3 > $OUT/nccl_mp_multibox_pipe$pb.log 2>&1
  done
  timeout 600 build/nccl_perf_mp_uccl_b200 -n "$N" -g "$N" -o allreduce -b 64K -e 1G -f 4 -i 10 -w 3 > $OUT/nccl_mp_onebox.log 2>&1
  for e in 1 2 4; do
    timeout 600 env UCCL_B200_NET_ENGINES=$e python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port 29512 benchmarks/multinode_allreduce.py --local-size $L --bind 127.0.0.1 --max-bytes 268435456 --iters 5 \
      > $OUT/multinode_allreduce_engines$e.log 2>&1
  done
fi
ls -la $OUT | tail -30
