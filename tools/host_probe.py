#!/usr/bin/env python
"""Host-side cost of one training step: wall-clock of ENQUEUEING steps (no synchronisation inside the window, short enough that the
launch queue never fills) against the device time of the same steps, plus a cProfile breakdown of where the enqueue time goes.
usage (GPU box): python tools/host_probe.py [--profile] > gpurun_out/host_probe.txt"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    from jnerf_b200 import plugin  # noqa: F401
    from jnerf_b200.runner import Runner, lego_cfg
    from jnerf_b200.utils.config import get_cfg, update_cfg
    get_cfg().clear()
    update_cfg(**lego_cfg(fp16=True, synthetic=True, seed=1))
    r = Runner()
    for _ in range(100):
        r.train_step()
    torch.cuda.synchronize()
    # windows that do not contain a 16-step edge (host sync at 15, grid update at 0): steps 1..14 of a window
    res = []
    for rep in range(8):
        while r.cfg.m_training_step % 16 != 1:
            r.train_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(13):
            r.train_step()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        res.append(((t1 - t0) / 13 * 1e6, e0.elapsed_time(e1) / 13 * 1e3))
    for h, d in res:
        print(f"host enqueue {h:7.1f} us/step   device {d:7.1f} us/step")
    # per-step device time across two 16-step windows (events between consecutive steps on the main stream)
    while r.cfg.m_training_step % 16 != 12:
        r.train_step()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
    t_host = []
    evs[0].record()
    for k in range(40):
        t0 = time.perf_counter()
        r.train_step()
        t_host.append((time.perf_counter() - t0) * 1e6)
        evs[k + 1].record()
    torch.cuda.synchronize()
    first = r.cfg.m_training_step - 40
    print("step%16: device us (host us)")
    print("  ".join(f"{(first + k) % 16}: {evs[k].elapsed_time(evs[k + 1]) * 1e3:.0f} ({t_host[k]:.0f})" for k in range(40)))
    # the occupancy-grid update of every 16th step, piece by piece
    from jnerf_b200 import ops
    s_, m_ = r.sampler, r.model

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    print(f"update_density_grid (whole)        {timed(s_.update_density_grid):8.1f} us")
    G3 = 128 ** 3
    n = G3 * (s_.max_cascade + 1) // 4
    pos_u, idx_u = ops.grid_generate_samples(n, s_.rng, s_.density_grid_ema_step, s_.aabb_range, s_.density_grid, s_.max_cascade + 1, -0.01)
    pos_n, idx_n = ops.grid_generate_samples(n, s_.rng, s_.density_grid_ema_step, s_.aabb_range, s_.density_grid, s_.max_cascade + 1, 0.01)
    pos = torch.cat([pos_u, pos_n])
    idx = torch.cat([idx_u, idx_n])
    print(f"generate_samples x2 + cat          {timed(lambda: (ops.grid_generate_samples(n, s_.rng, s_.density_grid_ema_step, s_.aabb_range, s_.density_grid, s_.max_cascade + 1, -0.01), ops.grid_generate_samples(n, s_.rng, s_.density_grid_ema_step, s_.aabb_range, s_.density_grid, s_.max_cascade + 1, 0.01), torch.cat([pos_u, pos_n]), torch.cat([idx_u, idx_n]))):8.1f} us")
    print(f"density of {pos.shape[0]} points, as generated   {timed(lambda: m_.density(pos)):8.1f} us")
    order = torch.argsort(idx.long())
    pos_sorted = pos[order].contiguous()
    print(f"density of the same points, cell-sorted     {timed(lambda: m_.density(pos_sorted)):8.1f} us")
    key = (idx.long() >> 12)                                        # 512 buckets of 16^3 cells (morton order)
    order_b = torch.argsort(key, stable=True)
    pos_b = pos[order_b].contiguous()
    print(f"density of the same points, 512 buckets     {timed(lambda: m_.density(pos_b)):8.1f} us")
    out = m_.density(pos).reshape(-1).contiguous()
    print(f"tmp.zero_ + splat + ema + bitfield + mean   {timed(lambda: (s_.density_grid_tmp.zero_(), ops.grid_splat(idx, out, s_.density_grid_tmp), ops.grid_ema(s_.density_grid, s_.density_grid_tmp, 0.95), ops.grid_update_bitfield(s_.density_grid, s_.density_grid_mean, s_.density_grid_bitfield, s_.NERF_CASCADES))):8.1f} us")
    print(f"argsort of {idx.numel()} keys                    {timed(lambda: torch.argsort(idx.long())):8.1f} us")
    if a.profile:
        pr = cProfile.Profile()
        torch.cuda.synchronize()
        pr.enable()
        for _ in range(64):
            r.train_step()
        pr.disable()
        torch.cuda.synchronize()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
        print(s.getvalue())


if __name__ == "__main__":
    main()
