#!/usr/bin/env bash
# MMA cost per operand layout (tc_time5) + the full GPU suite (all new parity tests, not stopping at the first failure)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/call3
mkdir -p "$OUT"
SUM="$OUT/SUMMARY.txt"
: > "$SUM"
run() { local name=$1 secs=$2; shift 2; local t0; t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; local rc=$?
        echo "$name: rc=$rc, $(( $(date +%s) - t0 )) s -- $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" >> "$SUM"; }
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tc_time5 tests/cuda/tc_time5.cu > "$OUT/nvcc_probes.log" 2>&1
run probe_tc_time5 60 /tmp/tc_time5
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1 || echo "build failed" >> "$SUM"
run pytest_gpu 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -s
cat "$SUM"
