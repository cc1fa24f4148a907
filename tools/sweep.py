#!/usr/bin/env python
"""BASELINE config #5: ray-batch sweep, target_batch_size 2^16 .. 2^22 samples/iter, on 1/2/4/8 GPUs -- one bench.py line per
point, collected into profiles/<tag>_sweep.json with the roofline-curve columns (it/s, rays/s, samples/s, fused forward /
backward GB/s and TFLOP/s from the per-stage CUDA events bench.py already takes).

    python tools/sweep.py --gpus 1 --tag r02                      # on the GPU box (gpurun -- 'python tools/sweep.py ...')
    python tools/sweep.py --gpus 1 --workload fox --log2 18       # config #3 at its one batch size

Every point is a separate bench.py process (torchrun for N > 1) so that buffers sized by target_batch_size are rebuilt."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_point(gpus, log2, workload, steps, pretrain, images, res, port):
    cmd = ["bench.py", "--gpus", str(gpus), "--steps", str(steps), "--warmup", "5", "--pretrain", str(pretrain), "--workload", workload,
           "--target-batch", str(1 << log2), "--no-cpu-baseline"]
    if images:
        cmd += ["--images", str(images)]
    if res:
        cmd += ["--res", str(res)]
    if gpus > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port)] + cmd
    else:
        cmd = [sys.executable] + cmd
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1800)
    line = next((l for l in reversed(p.stdout.splitlines()) if l.startswith("{")), None)
    if p.returncode != 0 or line is None:
        return {"error": (p.stderr or p.stdout)[-2000:], "returncode": p.returncode}
    return json.loads(line)


def summarise(log2, gpus, r):
    if "error" in r:
        return {"log2_target": log2, "gpus": gpus, "error": r["error"][-300:]}
    roof = r["roofline"]
    st = roof["stage_ms"]
    n = roof["samples_per_launch"]
    return {"log2_target": log2, "gpus": gpus, "iters_per_s": r["iters_per_s"], "rays_per_s": r["value"], "samples_per_s": r["samples_per_s"],
            "ms_per_step": r["ms_per_step"], "e2e_rays_per_s": r["e2e"]["value"], "samples_per_launch": n,
            "fwd_ms": st["network_fwd"], "bwd_ms": st["network_bwd"], "march_ms": st["march"], "composite_ms": st["composite_loss_bwd"],
            "optimizer_ms": st["adam_ema"],
            "fwd_gbs": n * 624 / (st["network_fwd"] * 1e-3) / 1e9, "bwd_gbs": n * 1124 / (st["network_bwd"] * 1e-3) / 1e9,
            "mlp_tflops_fwd": n * 20480 / (st["network_fwd"] * 1e-3) / 1e12, "mlp_tflops_bwd": n * 61440 / (st["network_bwd"] * 1e-3) / 1e12,
            "sm_mhz": (r.get("clocks") or {}).get("sm_mhz"), "clock_reasons": (r.get("clocks") or {}).get("reasons")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, nargs="+", default=[1])
    ap.add_argument("--log2", type=int, nargs="+", default=[16, 17, 18, 19, 20, 21, 22])
    ap.add_argument("--workload", default="lego", choices=["lego", "fox"])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--pretrain", type=int, default=256)
    ap.add_argument("--images", type=int, default=0)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--tag", default="r02")
    args = ap.parse_args()
    rows, raw = [], []
    for g in args.gpus:
        for k, l2 in enumerate(args.log2):
            r = run_point(g, l2, args.workload, args.steps, args.pretrain, args.images, args.res, 29511 + k)
            raw.append(r)
            rows.append(summarise(l2, g, r))
            print(json.dumps(rows[-1]), flush=True)
    out = os.path.join(ROOT, "profiles", f"{args.tag}_sweep_{args.workload}.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump({"workload": args.workload, "points": rows, "bench_lines": raw}, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
