#!/usr/bin/env bash
# A/B of the TMA path + full tests + sweep tail.  usage: gpu_ab.sh tag
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-ab}
mkdir -p "$OUT"
SUM="$OUT/SUMMARY.txt"
: > "$SUM"
run() { local name=$1 secs=$2; shift 2; local t0; t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; local rc=$?
        echo "$name: rc=$rc, $(( $(date +%s) - t0 )) s -- $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" >> "$SUM"; }
bench() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
          run "bench_$name" 300 env "${envs[@]}" python bench.py --steps 300 --warmup 5 --no-cpu-baseline "$@"
          python tools/bench_line_summary.py "$OUT/bench_$name.log" "$name" >> "$SUM"; }
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1 || echo "build failed" >> "$SUM"
run pytest_gpu 1200 python -m pytest tests -m gpu -q -p no:cacheprovider
bench tma NGP_GRAPHS=0 --
bench no_tma NGP_GRAPHS=0 NGP_NO_TMA=1 --
bench tma_again NGP_GRAPHS=0 --
bench default --
bench fox -- --workload fox
run ref_gpu_compare 400 python tools/ref_gpu_compare.py
for L in 20 21 22; do run sweep_$L 600 python tools/sweep.py --gpus 1 --tag r02tail --steps 100 --log2 $L; done
cp profiles/r02tail_sweep_lego.json "$OUT/" 2>/dev/null
cat "$SUM"
