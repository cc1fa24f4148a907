#!/usr/bin/env bash
# Short GPU call: selected tests (-k "$2"), lego + fox bench lines, march / hash / composite against the reference kernels.  usage: gpu_quick.sh tag [pytest -k expr]
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-quick}
mkdir -p "$OUT"
SUM="$OUT/SUMMARY.txt"
: > "$SUM"
run() { local name=$1 secs=$2; shift 2; local t0; t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; local rc=$?
        echo "$name: rc=$rc, $(( $(date +%s) - t0 )) s -- $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" >> "$SUM"; }
bench() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
          run "bench_$name" 300 env "${envs[@]}" python bench.py --steps 300 --warmup 5 --no-cpu-baseline "$@"
          python tools/bench_line_summary.py "$OUT/bench_$name.log" "$name" >> "$SUM"; }
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1 || echo "build failed" >> "$SUM"
if [ -n "${2:-}" ]; then run pytest_gpu 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "$2"; else run pytest_gpu 1200 python -m pytest tests -m gpu -q -p no:cacheprovider; fi
bench default --
bench no_graphs NGP_GRAPHS=0 --
bench fox -- --workload fox
run ref_gpu_compare 400 python tools/ref_gpu_compare.py
cat "$SUM"
