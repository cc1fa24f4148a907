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
