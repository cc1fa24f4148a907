#!/usr/bin/env bash
# Everything that was written after round 1's GPU budget ran out, measured in ONE gpurun call:
#     /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/first_gpu_call.sh'
# Every step runs under its own `timeout` (a hung kernel must not hold the box), writes to gpurun_out/first_call/, and a failing
# step does not stop the others.  Read gpurun_out/first_call/SUMMARY.txt first.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/first_call
mkdir -p "$OUT"
SUM="$OUT/SUMMARY.txt"
: > "$SUM"

run() {   # run <name> <seconds> <command...>
    local name=$1 secs=$2; shift 2
    local t0
    t0=$(date +%s)
    timeout "$secs" "$@" > "$OUT/$name.log" 2>&1
    local rc=$?
    echo "$name: rc=$rc, $(( $(date +%s) - t0 )) s -- $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" >> "$SUM"
}

bench() {  # bench <name> [ENV=VAL ...] -- [bench args...]
    local name=$1; shift
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    run "bench_$name" 300 env "${envs[@]}" python bench.py --steps 300 --warmup 5 --no-cpu-baseline "$@"
    python tools/bench_line_summary.py "$OUT/bench_$name.log" "$name" >> "$SUM"
}

python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1 || echo "build failed" >> "$SUM"

# 1. the default suite (includes tests/test_zz_link_compat.py: link symbols, .pkl checkpoints, fox end to end, no-sync renderer)
run pytest_gpu 900 python -m pytest tests -m gpu -q -p no:cacheprovider
# 2. the opt-in kernel variants
run pytest_experimental 900 env NGP_EXPERIMENTAL=1 python -m pytest tests/test_zz_experimental.py -q -p no:cacheprovider
run pytest_experimental_sw8 600 env NGP_EXPERIMENTAL=1 NGP_BWD_SCATTER_WARPS=8 python -m pytest tests/test_zz_experimental.py -q -p no:cacheprovider -k saved

# 3. probes: MMA time per operand-major combination; one chain stage with parts switched off, 1/2/4 chains per CTA
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tc_time4 tests/cuda/tc_time4.cu > "$OUT/nvcc_probes.log" 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include -o /tmp/tc_stage tests/cuda/tc_stage.cu >> "$OUT/nvcc_probes.log" 2>&1
run probe_tc_time4 60 /tmp/tc_time4
run probe_tc_stage 120 /tmp/tc_stage

# 3b. standalone hash kernels on ray-ordered samples of a real training state, default vs run-length variants (and the reference's own kernels)
run ref_gpu_compare_default 400 python tools/ref_gpu_compare.py
run ref_gpu_compare_hash_runlen 400 env NGP_HASH_RUNLEN=1 python tools/ref_gpu_compare.py
run microbench 300 python tools/microbench.py

# 4. A/B of the opt-in variants on the headline workload (roofline.stage_ms shows the stage each one touches)
bench default --
bench save_act NGP_SAVE_ACT=1 --
bench save_act_sw8 NGP_SAVE_ACT=1 NGP_BWD_SCATTER_WARPS=8 --
bench march_pipe NGP_MARCH_PIPE=1 --
bench composite_pipe NGP_COMPOSITE_PIPE=1 --
bench all_optin NGP_SAVE_ACT=1 NGP_MARCH_PIPE=1 NGP_COMPOSITE_PIPE=1 --
bench bwd_no_wgrad_TIMING_ONLY NGP_BWD_DEBUG=4 --
bench bwd_no_scatter_TIMING_ONLY NGP_BWD_DEBUG=2 --
bench default_again --

# 5. BASELINE config #3 (fox stand-in) and config #5 (batch-size sweep, one GPU)
bench fox -- --workload fox
python -c "import sys; sys.path.insert(0, 'tests/golden'); from make_fox_small import materialise; materialise('/tmp/fox_small')" > "$OUT/fox_small.log" 2>&1
bench fox_real_capture_180x320 -- --workload fox --data-dir /tmp/fox_small
run sweep_1gpu 900 python tools/sweep.py --gpus 1 --tag r02 --steps 100
cp profiles/r02_sweep_lego.json "$OUT/" 2>/dev/null

# 6. clock64 timeline of one backward tile (stage B2 step by step); restores the default build afterwards
if NGP_NVCC_FLAGS=-DNGP_TIMELINE python jnerf_b200/build.py --force > "$OUT/build_timeline.log" 2>&1; then
    run timeline_bwd 120 python tools/dbg_timeline_bwd.py
fi
python jnerf_b200/build.py --force > "$OUT/build_restore.log" 2>&1 || echo "RESTORING THE DEFAULT BUILD FAILED" >> "$SUM"

cat "$SUM"
