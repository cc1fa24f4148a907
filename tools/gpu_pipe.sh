#!/usr/bin/env bash
# A/B of the software pipeline over steps (front of step i+1 under step i).  usage: gpu_pipe.sh tag
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-pipe}
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
bench seq NGP_PIPELINE=0 --
bench pipe_fwd NGP_PIPE_AT=fwd --
bench pipe_front NGP_PIPE_AT=front --
bench fox_fwd NGP_PIPE_AT=fwd -- --workload fox
bench fox_front NGP_PIPE_AT=front -- --workload fox
run psnr_lego 300 python tools/train_psnr.py --steps 3000 --evals 1000,3000 --out $OUT/psnr_lego.json
cat "$SUM"
