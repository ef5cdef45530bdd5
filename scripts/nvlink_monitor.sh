#!/usr/bin/env bash
# NVLink throughput monitor (role of the reference's scripts/bw_monitor.sh for NICs): prints per-GPU
# TX/RX GB/s computed from the nvidia-smi NVLink byte counters every INTERVAL seconds.
INTERVAL=${1:-1}
read_counters() {
  nvidia-smi nvlink -gt d 2>/dev/null | awk '
    /^GPU/ {gpu=$2; sub(":", "", gpu)}
    /Data Tx:/ {tx[gpu]+=$(NF-1)}
    /Data Rx:/ {rx[gpu]+=$(NF-1)}
    END {for (g in tx) printf "%s %d %d\n", g, tx[g], rx[g]}' | sort -n
}
prev=$(read_counters)
if [ -z "$prev" ]; then echo "nvidia-smi nvlink counters unavailable" >&2; exit 1; fi
while true; do
  sleep "$INTERVAL"
  cur=$(read_counters)
  paste <(echo "$prev") <(echo "$cur") | awk -v dt="$INTERVAL" '
    {printf "GPU%s  tx %7.1f GB/s  rx %7.1f GB/s\n", $1, ($5-$2)/1048576/dt, ($6-$3)/1048576/dt}'
  echo "--"
  prev=$cur
done
