#!/usr/bin/env python
"""Data-parallel contract on real GPUs: W ranks x R rays == 1 rank x W*R rays.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py

Trains 300 steps data-parallel in both exchange modes (p2p: ngp_dp_exchange_step over NVLink peer memory; nccl: reduce-scatter /
all-gather) and checks
  (1) the first forward pass equals the single-GPU forward of the global batch (same rays, jitter stream, background),
  (2) every rank holds a bit-identical hash table and MLP weights after every exchange,
  (3) the two exchange modes follow the same loss trajectory, and training converges (loss, PSNR of a training view),
  (4) a checkpoint written from the sharded optimizer state restores the same state on every rank.
Parameter-level equality with a single-GPU run of the global batch is NOT expected: fp16 gradient accumulation underflows
differently under the per-rank loss scale 128/R_local (calc_rgb.h:100-101) and Adam's first steps move every touched entry by
+-lr; the gradient-level contract is tested in tests/test_gpu_runner.py::test_dp_shards_reproduce_single_rank_gradients.
Prints one JSON line on rank 0 and exits non-zero on failure."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def make_runner(rank, world, pg, rays, target_batch=1 << 18):
    from jnerf_b200.runner import Runner, lego_cfg
    from jnerf_b200.utils.config import get_cfg, update_cfg
    get_cfg().clear()
    update_cfg(**lego_cfg(fp16=True, synthetic=True, seed=1, target_batch_size=target_batch))
    cfg = get_cfg()
    cfg.dataset.train.n_images = 12
    cfg.dataset.train.H = cfg.dataset.train.W = 200
    cfg.dataset.val = None
    cfg.dataset.train.pop("root_dir", None)
    r = Runner(rank=rank, world_size=world, process_group=pg)
    r.sampler.n_rays_per_batch = rays              # first 16 steps; cfg.n_rays_per_batch (4096) keeps sizing the raw sample capacity
    return r


def psnr_views(r, views=(0, 3, 6, 9)):
    """Mean PSNR over four training views (one 200 x 200 image alone moves by several hundredths of a dB between two identical runs)."""
    tot = 0.0
    for v in views:
        img, tar = r.render_img("train", v)
        tot += float(-10.0 * torch.log10(((img - tar) ** 2).mean()).item())
    return tot / len(views)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pg = dist.group.WORLD
    from jnerf_b200 import lib, plugin  # noqa: F401
    lib.load()
    K, R = 300, 96                                 # R small enough that the 2^18-sample capacity never truncates in the first steps
    ok = {}

    # single-GPU forward of the global batch: the data-parallel shards must see the same rays, jitter and background
    r1 = make_runner(0, 1, None, R * world, target_batch=world << 18)   # the global batch must not be truncated to ONE rank's capacity
    loss_single = float(r1.train_step().float().mean().item())
    del r1
    # equal global batch, equal iterations on ONE GPU (SURVEY 8e's criterion), twice: the second run measures how far two runs of the
    # very same single-GPU configuration drift apart (gradient atomics are summed in a different order every time)
    psnr_single, loss_single_traj, rays_single = [], [], []
    for _ in range(2):
        r1 = make_runner(0, 1, None, R * world, target_batch=world << 18)     # the ray batch adapts to the GLOBAL sample budget, as the W ranks' does
        tr = []
        for _ in range(K):
            rays_single.append(r1.sampler.n_rays_per_batch)
            tr.append(r1.train_step().float().mean().reshape(1))
        loss_single_traj.append([float(x) for x in torch.cat(tr).tolist()])
        psnr_single.append(psnr_views(r1))
        del r1

    tables, losses = {}, {}
    for mode in ("p2p", "nccl"):
        os.environ["NGP_DP_EXCHANGE"] = mode
        rw = make_runner(rank, world, pg, R)
        assert rw.dp_mode == mode
        ls, rays_dp = [], []
        for _ in range(K):
            rays_dp.append(rw.sampler.n_rays_per_batch * world)
            l = rw.train_step().float().mean().reshape(1)
            dist.all_reduce(l)
            ls.append(float(l.item()) / world)
        rw._table_ready()
        torch.cuda.synchronize()
        table = rw.model.pos_encoder.m_grid.data
        ref = table.clone()
        dist.broadcast(ref, src=0)
        ok[f"{mode}_tables_identical_across_ranks"] = bool(torch.equal(ref, table))
        wref = rw.model.rgb_mlp.con_weights.data.clone()
        dist.broadcast(wref, src=0)
        ok[f"{mode}_weights_identical_across_ranks"] = bool(torch.equal(wref, rw.model.rgb_mlp.con_weights.data))
        ok[f"{mode}_train_view_psnr_db"] = psnr_views(rw)
        # checkpoint round trip from the sharded optimizer state
        path = os.path.join(tempfile.gettempdir(), f"dp_check_{mode}.ckpt")
        rw.save_ckpt(path)
        dist.barrier()
        st = rw._st[id(rw.model.pos_encoder.m_grid)]
        m0, v0, ms0 = st.m.clone(), st.v.clone(), st.master.clone()
        st.m.zero_(); st.v.zero_(); st.master.zero_()
        rw.load_ckpt(path)
        ok[f"{mode}_ckpt_roundtrip"] = bool(torch.equal(st.m, m0) and torch.equal(st.v, v0) and torch.equal(st.master, ms0))
        tables[mode], losses[mode] = table.float().clone(), ls
        ok[f"{mode}_global_rays_every_16"] = rays_dp[::16]
        del rw
    ok["world_size"] = world
    # every rank has trained the same single-GPU configuration twice: 2 W samples of the single-GPU result (the runs differ only in the
    # order the gradient atomics were summed in), i.e. its mean and its spread
    every = [None] * world
    dist.all_gather_object(every, psnr_single)
    every = [x for pair in every for x in pair]
    mean_single = sum(every) / len(every)
    ok["single_gpu_global_batch_train_view_psnr_db"] = every
    ok["single_gpu_psnr_mean_db"] = mean_single
    ok["single_gpu_psnr_std_db"] = (sum((x - mean_single) ** 2 for x in every) / (len(every) - 1)) ** 0.5
    ok["dp_minus_single_psnr_db"] = {m: ok[f"{m}_train_view_psnr_db"] - mean_single for m in ("p2p", "nccl")}
    ok["single_gpu_run_to_run_psnr_db"] = max(every) - min(every)
    ok["loss_step0_single_gpu_global_batch"] = loss_single
    ok["loss_step0_dp"] = {m: losses[m][0] for m in losses}
    ok["single_gpu_rays_every_16"] = rays_single[:K:16]
    # where the trajectories part: relative loss difference against the first single-GPU run, per step (the second single-GPU run
    # gives the floor: two runs of the same configuration)
    rel = lambda a, b: [abs(x - y) / abs(y) for x, y in zip(a, b)]
    for name, tr in (("single_rerun", loss_single_traj[1]), ("p2p", losses["p2p"]), ("nccl", losses["nccl"])):
        d = rel(tr, loss_single_traj[0])
        ok[f"loss_rel_diff_vs_single_{name}"] = {"steps_0_15": max(d[:16]), "steps_16_63": max(d[16:64]), "steps_64_299": max(d[64:]),
                                                  "first_step_over_1e-3": next((i for i, x in enumerate(d) if x > 1e-3), None)}
    ok["loss_first8"] = {m: [round(x, 5) for x in losses[m][:8]] for m in losses}
    ok["loss_last"] = {m: sum(losses[m][-20:]) / 20 for m in losses}
    ok["p2p_vs_nccl_first8_max_rel_diff"] = max(abs(a - b) / abs(a) for a, b in zip(losses["p2p"][:8], losses["nccl"][:8]))
    good = all(v for k, v in ok.items() if k.endswith(("_ranks", "_roundtrip")))
    good = good and all(abs(v - loss_single) <= 1e-3 * loss_single for v in ok["loss_step0_dp"].values())
    good = good and ok["p2p_vs_nccl_first8_max_rel_diff"] < 0.02
    good = good and all(ok[f"{m}_train_view_psnr_db"] > 20.0 for m in ("p2p", "nccl")) and all(v < 0.5 * loss_single for v in ok["loss_last"].values())
    # PSNR at equal global batch and iterations: within 0.15 dB of the mean of the 2 W single-GPU runs, or within three
    # of their standard deviations (measured at W = 2 and 8: p2p -0.12 .. -0.00, NCCL -0.04 .. +0.09 dB -- the two exchange modes, which
    # differ only in the order / precision of the cross-rank sum, end up to 0.19 dB apart from each other).  Training is chaotic at the level of the gradient atomics' summation order: two identical single-GPU runs end
    # 0.004-0.05 dB apart after 300 steps, the p2p and the NCCL exchange (same arithmetic, different order) up to 0.1 dB.
    good = good and all(abs(v) <= max(0.15, 3 * ok["single_gpu_psnr_std_db"]) for v in ok["dp_minus_single_psnr_db"].values())
    flag = torch.tensor([int(good)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok["pass"] = bool(flag.item())
    if rank == 0:
        print(json.dumps(ok), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(ok, open(os.path.join(ROOT, "gpurun_out", f"dp_check_{world}gpu.json"), "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok["pass"] else 1)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(2)
