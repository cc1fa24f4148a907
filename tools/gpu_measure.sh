#!/usr/bin/env bash
# Measurement session: ncu full captures of the two fused kernels, launch list, config #5 sweep, fox lines, PSNR-vs-seconds runs.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-measure}
mkdir -p "$OUT"
SUM="$OUT/SUMMARY.txt"
: > "$SUM"
run() { local name=$1 secs=$2; shift 2; local t0; t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; local rc=$?
        echo "$name: rc=$rc, $(( $(date +%s) - t0 )) s -- $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" >> "$SUM"; }
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1 || echo "build failed" >> "$SUM"
export NGP_PROFILE=1
run ncu_launches 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/launches.csv" python bench.py --steps 3 --warmup 3 --no-cpu-baseline
for K in network_bwd256 network_fwd march_count composite_loss_bwd adam_ema; do
  run ncu_full_$K 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$K" -c 1 -o "$OUT/prof_$K" -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline
done
unset NGP_PROFILE
run bench_lego 300 python bench.py --steps 1000 --warmup 5
run bench_fox 300 python bench.py --workload fox --steps 500 --warmup 5 --no-cpu-baseline
python -c "import sys; sys.path.insert(0, 'tests/golden'); from make_fox_small import materialise; materialise('/tmp/fox_small')" > "$OUT/fox_small.log" 2>&1
run bench_fox_real 300 python bench.py --workload fox --data-dir /tmp/fox_small --steps 500 --warmup 5 --no-cpu-baseline
run psnr_fox_real 600 python tools/train_psnr.py --workload fox --data-dir /tmp/fox_small --steps 5000 --evals 250,500,1000,2000,5000 --val-images 4 --budget-seconds 20 --out "$OUT/psnr_fox_small.json"
run psnr_lego_synth 900 python tools/train_psnr.py --steps 20000 --evals 500,1000,2000,5000,10000,20000 --images 100 --res 400 --out "$OUT/psnr_lego_synthetic.json"
run sweep_1gpu 1200 python tools/sweep.py --gpus 1 --tag r02 --steps 100
cp profiles/r02_sweep_lego.json "$OUT/" 2>/dev/null
cat "$SUM"
