#!/usr/bin/env bash
# One GPU call: layout probe, the full GPU suite, bench (lego, fox, timing-only knobs), microbench.  usage: gpu_check.sh [tag]
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-check}
mkdir -p "$OUT"
SUM="$OUT/SUMMARY.txt"
: > "$SUM"
run() { local name=$1 secs=$2; shift 2; local t0; t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; local rc=$?
        echo "$name: rc=$rc, $(( $(date +%s) - t0 )) s -- $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" >> "$SUM"; }
bench() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
          run "bench_$name" 300 env "${envs[@]}" python bench.py --steps 300 --warmup 5 --no-cpu-baseline "$@"
          python tools/bench_line_summary.py "$OUT/bench_$name.log" "$name" >> "$SUM"; }
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tc_time6 tests/cuda/tc_time6.cu > "$OUT/nvcc_probes.log" 2>&1
run probe_tc_time6 60 /tmp/tc_time6
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1 || echo "build failed" >> "$SUM"
run pytest_gpu 1200 python -m pytest tests -m gpu -q -p no:cacheprovider
bench default --
bench bwd_no_wgrad_TIMING_ONLY NGP_BWD_DEBUG=4 --
bench bwd_no_scatter_TIMING_ONLY NGP_BWD_DEBUG=2 --
bench fox -- --workload fox
run microbench 300 python tools/microbench.py
run ref_gpu_compare 400 python tools/ref_gpu_compare.py
cat "$SUM"
