"""N>1 host logic on CPU: world_size-2 gloo processes exercise jnerf_b200/dp.py (sharding, RNG stream offset, gradient
all-reduce + 1/W scaling, global batch-size adaptation) with the oracle standing in for the CUDA kernels.
Contract: W ranks with global batch B reproduce one rank with batch B."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from jnerf_b200 import dp
    torch.set_num_threads(1)
    # global batch (identical on every rank), contiguous shard per rank
    n = 96
    o, d = ol.random_rays(n * world, seed=3)
    bits, _ = ol.sphere_bitfield(0.3, shell=0.03)
    lo, hi = dp.shard_range(n, rank)
    rng = ol.pcg32_seed()
    rng_local = ol.pcg32_advance(rng.copy(), dp.ray_stream_offset(n, rank))
    coords, _, numsteps, cnt = ol.march(o[lo:hi], d[lo:hi], bits, max_samples=n * 1024, rng=rng_local)
    S = int(cnt[1])
    # "gradient" of this shard with the local normalisation 128/R_local folded in
    cfg = ol.HashCfg(1, log2_hashmap_size=14)
    dy = (np.random.default_rng(7).standard_normal((n * world * 1024, 32)) * 1e-3).astype(np.float32)
    base = sum_prev = 0
    pos = np.ascontiguousarray(coords[:S, :3])
    g_local = torch.from_numpy(ol.hash_bwd(cfg, pos, dy[:S] * (128.0 / n)))
    count = torch.tensor([S], dtype=torch.int32)
    scale = dp.allreduce_grads((g_local,), None, world)
    dp.global_mean_count(count, None, world)
    new_rays = dp.adapt_rays_per_batch(n, int(count.item()), 1 << 12)
    ret[rank] = dict(numsteps=numsteps.copy(), coords=coords[:S].copy(), grad=(g_local * scale).numpy(), count=int(count.item()),
                     new_rays=new_rays, S=S)
    dist.destroy_process_group()


def test_two_rank_data_parallel_matches_single_rank():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29511 + os.getpid() % 1000, ret), nprocs=world, join=True)
    n = 96
    o, d = ol.random_rays(n * world, seed=3)
    bits, _ = ol.sphere_bitfield(0.3, shell=0.03)
    coords, _, numsteps, cnt = ol.march(o, d, bits, max_samples=n * world * 1024, rng=ol.pcg32_seed())
    # (1) sample indices: the two shards, concatenated, are bit-identical to the single-rank march of the global batch
    assert np.array_equal(np.concatenate([ret[0]["numsteps"][:, 0], ret[1]["numsteps"][:, 0]]), numsteps[:, 0])
    S0, S1 = ret[0]["S"], ret[1]["S"]
    assert S0 + S1 == int(cnt[1])
    both = np.concatenate([ret[0]["coords"], ret[1]["coords"]])
    assert np.array_equal(both.view(np.uint32), coords[:S0 + S1].view(np.uint32))
    # (2) all-reduced, 1/W-scaled gradient == gradient of the global batch normalised by 128/R_global
    cfg = ol.HashCfg(1, log2_hashmap_size=14)
    dy = (np.random.default_rng(7).standard_normal((n * world * 1024, 32)) * 1e-3).astype(np.float32)
    dy_glob = np.concatenate([dy[:S0], dy[:S1]]) * (128.0 / (n * world))
    g_ref = ol.hash_bwd(cfg, np.ascontiguousarray(coords[:S0 + S1, :3]), dy_glob)
    assert np.allclose(ret[0]["grad"], ret[1]["grad"], rtol=0, atol=0)
    assert np.abs(ret[0]["grad"] - g_ref).max() <= 1e-5 * np.abs(g_ref).max()
    # (3) every rank derives the same adaptive batch size from the all-reduced counter
    assert ret[0]["count"] == ret[1]["count"] == (S0 + S1) // 2 and ret[0]["new_rays"] == ret[1]["new_rays"]


def test_adapt_rays_per_batch_matches_reference_formula():
    sys.path.insert(0, ROOT)
    from jnerf_b200 import dp
    # density_grid_sampler.py:266-271: int(min(div_round_up(int(R*target/measured),128)*128, target))
    for R, measured, target in ((4096, 500000.0, 1 << 18), (4096, 10.0, 1 << 18), (1920, 262144.0, 1 << 18), (128, 0.0, 1 << 18)):
        m = max(measured, 1)
        ref = int(min(((int(R * target / m)) + 127) // 128 * 128, target))
        assert dp.adapt_rays_per_batch(R, measured, target) == ref


def _sharded_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from jnerf_b200 import dp
    torch.set_num_threads(1)
    n = 6_098_120 // 64 * 2 + 6                                    # not a multiple of world*256: exercises the padding
    P = dp.padded_len(n, world)
    lo, hi = dp.slice_bounds(P, world, rank)
    g = torch.Generator().manual_seed(100 + rank)
    grad = torch.zeros(P)
    grad[:n] = torch.randn(n, generator=g)
    table = torch.zeros(P)
    table[:n] = torch.arange(n, dtype=torch.float32) * 1e-3       # identical on every rank
    # reduce-scatter -> "optimizer" on the slice (plain SGD stands in for ngp_adam_ema) -> all-gather
    my_grad = torch.empty(hi - lo)
    dp.reduce_scatter_sum(my_grad, grad, None, world, rank)
    my_slice = table[lo:hi] - 0.1 * my_grad / world
    work = dp.all_gather_slices(table, my_slice.contiguous(), None, world, async_op=True)
    work.wait()
    ret[rank] = dict(table=table[:n].numpy().copy(), P=P, lo=lo, hi=hi, tail=float(table[n:].abs().sum()))
    dist.destroy_process_group()


def test_sharded_table_optimizer_equals_allreduce_update():
    """dp.py sharded optimizer plumbing: reduce-scatter + per-slice update + all-gather == all-reduce + full update, on every rank."""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_sharded_worker, args=(world, 30511 + os.getpid() % 1000, ret), nprocs=world, join=True)
    n = 6_098_120 // 64 * 2 + 6
    assert ret[0]["P"] % (world * 256) == 0 and ret[0]["P"] >= n and ret[0]["hi"] == ret[1]["lo"] and ret[1]["hi"] == ret[0]["P"]
    gsum = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
    want = torch.arange(n, dtype=torch.float32) * 1e-3 - 0.1 * gsum / world
    for r in range(world):
        assert np.array_equal(ret[r]["table"], want.numpy()) and ret[r]["tail"] == 0.0
