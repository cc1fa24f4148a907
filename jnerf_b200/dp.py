"""Data-parallel host logic: one process per GPU, torch.distributed for the plumbing.

The reference has no multi-GPU path for NGP; the contract is "W ranks with global batch B == one GPU with batch B":
  * every rank draws the same global pixel batch and takes the contiguous shard [rank*n, (rank+1)*n);
  * per-ray jitter is indexed by the GLOBAL ray id (ray_sampler.h:30), i.e. the pcg32 state is advanced by shard_start*8;
  * one all-reduce (sum) per gradient buffer per step; because each rank normalises by its LOCAL ray count
    (loss_scale = 128/R_local, calc_rgb.h:100-101) the sum is W x the global-batch gradient -> scale by 1/W in the optimizer;
  * the occupancy grid is replicated: its update consumes only parameters and the shared RNG stream, both identical on all ranks;
  * the adaptive ray-batch size is derived from the all-reduced sample counter so that every rank picks the same value.
Works with NCCL (GPU) and gloo (CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_per_rank, rank):
    """[start, end) of this rank's rays inside the global batch of world_size * n_per_rank rays."""
    return rank * n_per_rank, (rank + 1) * n_per_rank


def ray_stream_offset(n_per_rank, rank, samples_per_ray=8):
    """pcg32 advance that makes local ray i consume the stream of global ray rank*n_per_rank + i (N_MAX_RANDOM_SAMPLES_PER_RAY = 8)."""
    return rank * n_per_rank * samples_per_ray


def allreduce_grads(buffers, group, world_size):
    """Sum the gradient buffers over ranks; returns the factor the optimizer must apply (1/world_size)."""
    if world_size > 1:
        for b in buffers:
            dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world_size


def global_mean_count(counter, group, world_size):
    """All-reduce an integer sample counter and return the per-rank mean (identical on every rank)."""
    if world_size > 1:
        dist.all_reduce(counter, op=dist.ReduceOp.SUM, group=group)
        counter //= world_size
    return counter


def adapt_rays_per_batch(n_rays_per_batch, measured_per_step, target_batch_size):
    """DensityGridSampler.update_batch_rays (density_grid_sampler.py:266-271) as a pure function."""
    measured = max(measured_per_step, 1)
    rays = int(n_rays_per_batch * target_batch_size / measured)
    return int(min(((rays + 127) // 128) * 128, target_batch_size))
