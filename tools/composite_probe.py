#!/usr/bin/env python
"""Where the fused training tail (ngp_composite_loss_bwd) spends its time: synthetic ray batches of known lengths, CUDA-event timing.
usage (GPU box): python tools/composite_probe.py > gpurun_out/composite_probe.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from jnerf_b200 import ops  # noqa: E402


def run(name, lengths, reps=50):
    dev = "cuda"
    R = len(lengths)
    n = torch.tensor(lengths, dtype=torch.int32, device=dev)
    base = torch.cumsum(n, 0, dtype=torch.int32) - n
    ns = torch.stack([n, base], 1).contiguous()
    S = int(n.sum())
    g = torch.Generator(device=dev).manual_seed(1)
    net = (torch.randn((S, 4), device=dev, generator=g) * 0.5).half()
    net[:, 3] -= 3.0                                             # thin medium: transmittance survives along the ray
    coords = torch.rand((S, 7), device=dev, generator=g)
    coords[:, 3] = 0.02
    bg = torch.rand((R, 3), device=dev, generator=g)
    target = torch.rand((R, 3), device=dev, generator=g)
    mean = torch.full((1,), 0.5, device=dev)
    dnet = torch.empty_like(net)
    for _ in range(5):
        ops.composite_loss_bwd(net, coords, ns, ns, bg, target, mean, dnet=dnet)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.composite_loss_bwd(net, coords, ns, ns, bg, target, mean, dnet=dnet)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:44s} rays {R:5d} samples {S:8d}  {e0.elapsed_time(e1) / reps * 1e3:8.2f} us", flush=True)


def main():
    run("1 ray x 32", [32])
    run("1 ray x 128", [128])
    run("1 ray x 512", [512])
    run("1 ray x 1024", [1024])
    run("4096 rays x 64", [64] * 4096)
    run("4096 rays x 32", [32] * 4096)
    run("2048 rays x 128", [128] * 2048)
    run("3000 rays x 87", [87] * 3000)
    run("4095 x 56 + 1 x 1024", [56] * 4095 + [1024])
    run("4032 x 48 + 64 x 1024", [48] * 4032 + [1024] * 64)
    run("2900 x 64 + 100 x 600", [64] * 2900 + [600] * 100)
    # a trained-scene-like mix: 40 % empty rays, the rest geometric up to 600
    torch.manual_seed(0)
    mix = (torch.rand(3000) < 0.4).int() * 0
    L = torch.clamp((torch.rand(3000) ** 2 * 600).int(), 1, 600)
    L[torch.rand(3000) < 0.4] = 0
    run("mix: 40 % empty, rest ~ u^2 * 600", L.tolist())


if __name__ == "__main__":
    main()
