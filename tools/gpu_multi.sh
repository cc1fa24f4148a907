#!/usr/bin/env bash
# Multi-GPU session on ONE box with N GPUs visible: dp_check at N, weak scaling at 1/2/4/8 (up to N), strong scaling at N.  usage: gpu_multi.sh tag N
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-multi}; N=${2:-2}
mkdir -p "$OUT"
SUM="$OUT/SUMMARY.txt"
: > "$SUM"
run() { local name=$1 secs=$2; shift 2; local t0; t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; local rc=$?
        echo "$name: rc=$rc, $(( $(date +%s) - t0 )) s -- $(tail -n 1 "$OUT/$name.log" | cut -c1-400)" >> "$SUM"; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1 || echo "build failed" >> "$SUM"
run dp_check_$N 900 $TR --nproc-per-node $N --master-port 29533 tools/dp_check.py
cp gpurun_out/dp_check_${N}gpu.json "$OUT/" 2>/dev/null
run bench_weak_1 300 python bench.py --gpus 1 --steps 300 --warmup 5 --no-cpu-baseline
for W in 2 4 8; do
  if [ "$W" -le "$N" ]; then
    run bench_weak_$W 400 $TR --nproc-per-node $W --master-port $((29540 + W)) bench.py --gpus $W --steps 300 --warmup 5 --no-cpu-baseline
    run bench_strong_$W 400 $TR --nproc-per-node $W --master-port $((29560 + W)) bench.py --gpus $W --steps 300 --warmup 5 --no-cpu-baseline --scaling strong
  fi
done
cat "$SUM"
