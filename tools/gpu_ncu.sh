#!/usr/bin/env bash
# ncu: launch list of 3 steady-state steps + one full capture of a kernel.  usage: gpu_ncu.sh tag kernel_regex [extra bench args]
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-ncu}; K=${2:-network_bwd}; shift 2 || true
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1
NGP_PROFILE=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > "$OUT/bench_launches.log" 2>&1
NGP_PROFILE=1 timeout 800 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$K" -c 1 -o "$OUT/prof_$K" -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > "$OUT/bench_full.log" 2>&1
ls -la "$OUT"
