#!/usr/bin/env bash
# 8-GPU session (charged 8x: keep it short): dp_check, weak scaling at 8, strong scaling at 8, and the 1-GPU line of the same box.  usage: gpu_multi8.sh tag [N]
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-multi8}; N=${2:-8}
mkdir -p "$OUT"
SUM="$OUT/SUMMARY.txt"
: > "$SUM"
run() { local name=$1 secs=$2; shift 2; local t0; t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; local rc=$?
        echo "$name: rc=$rc, $(( $(date +%s) - t0 )) s -- $(tail -n 1 "$OUT/$name.log" | cut -c1-400)" >> "$SUM"; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run bench_weak_$N 200 $TR --nproc-per-node $N --master-port $((29540 + N)) bench.py --gpus $N --steps 300 --warmup 5 --no-cpu-baseline
run bench_weak_1 120 python bench.py --gpus 1 --steps 300 --warmup 5 --no-cpu-baseline
run dp_check_$N 300 $TR --nproc-per-node $N --master-port 29533 tools/dp_check.py
cp gpurun_out/dp_check_${N}gpu.json "$OUT/" 2>/dev/null
if [ "${3:-}" = "strong" ]; then run bench_strong_$N 200 $TR --nproc-per-node $N --master-port $((29560 + N)) bench.py --gpus $N --steps 300 --warmup 5 --no-cpu-baseline --scaling strong; fi
cat "$SUM"
