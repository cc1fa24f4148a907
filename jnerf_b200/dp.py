"""Data-parallel host logic: one process per GPU, torch.distributed for the plumbing.

The reference has no multi-GPU path for NGP; the contract is "W ranks with global batch B == one GPU with batch B":
  * every rank draws the same global pixel batch and takes the contiguous shard [rank*n, (rank+1)*n);
  * per-ray jitter is indexed by the GLOBAL ray id (ray_sampler.h:30), i.e. the pcg32 state is advanced by shard_start*8;
  * one all-reduce (sum) per gradient buffer per step; because each rank normalises by its LOCAL ray count
    (loss_scale = 128/R_local, calc_rgb.h:100-101) the sum is W x the global-batch gradient -> scale by 1/W in the optimizer;
  * the occupancy grid is replicated: its update consumes only parameters and the shared RNG stream, both identical on all ranks;
  * the adaptive ray-batch size is derived from the all-reduced sample counter so that every rank picks the same value.
  * the 12.2 M-entry hash table -- 99.9 % of the parameters -- uses a SHARDED optimizer: reduce-scatter of its gradient, fused
    Adam+EMA on this rank's 1/W slice only (optimizer state exists only for that slice), all-gather of the updated fp16
    slice.  Same bytes on NVLink as the all-reduce, 1/W of the 171 MB optimizer-state traffic per GPU, and the all-gather
    runs under the next step's ray generation + march (neither reads the table).  The two small MLP weight tensors keep
    the plain all-reduce (one flat fp32 buffer that also carries the sample counter).
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


def padded_len(n, world_size, align=256):
    """Smallest length >= n that splits into world_size slices of a multiple of `align` elements (vector-load alignment)."""
    q = world_size * align
    return (n + q - 1) // q * q


def slice_bounds(padded_total, world_size, rank):
    """[lo, hi) of this rank's slice of a padded_len()-sized buffer."""
    per = padded_total // world_size
    assert per * world_size == padded_total
    return rank * per, (rank + 1) * per


def reduce_scatter_sum(out_slice, full, group, world_size, rank):
    """out_slice <- sum over ranks of full[slice_bounds(rank)]."""
    lo, hi = slice_bounds(full.numel(), world_size, rank)
    if world_size == 1:
        out_slice.copy_(full[lo:hi])
    elif dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(out_slice, full, op=dist.ReduceOp.SUM, group=group)
    else:                                        # gloo has no reduce-scatter: all-reduce, keep the slice
        tmp = full.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        out_slice.copy_(tmp[lo:hi])


def all_gather_slices(full, my_slice, group, world_size, async_op=False):
    """full <- concatenation over ranks of my_slice.  Returns the work handle when async_op (call .wait() before reading `full`)."""
    if world_size == 1:
        full.copy_(my_slice)
        return None
    return dist.all_gather_into_tensor(full, my_slice, group=group, async_op=async_op)


def global_mean_count(counter, group, world_size):
    """All-reduce an integer sample counter and return the per-rank mean (identical on every rank)."""
    if world_size > 1:
        dist.all_reduce(counter, op=dist.ReduceOp.SUM, group=group)
        counter //= world_size
    return counter


def global_sum_count(counter, group, world_size):
    """All-reduce an integer sample counter in place (the sum over ranks, identical on every rank)."""
    if world_size > 1:
        dist.all_reduce(counter, op=dist.ReduceOp.SUM, group=group)
    return counter


def adapt_rays_per_batch(n_rays_per_batch, measured_per_step, target_batch_size):
    """DensityGridSampler.update_batch_rays (density_grid_sampler.py:266-271) as a pure function."""
    measured = max(measured_per_step, 1)
    rays = int(n_rays_per_batch * target_batch_size / measured)
    return int(min(((rays + 127) // 128) * 128, target_batch_size))


# ------------------------------------------------------------------------------------------------------------------
# NVLink peer-memory exchange (ngp_dp_exchange_step): every rank maps every rank's arena through CUDA IPC.
class PeerArena:
    """One device allocation per rank holding everything the peers touch:
         [ table fp16 P | table_grad fp16 P | w_grad fp32 n_w_pad | flags u32 64 ]
    so that a single IPC handle per rank is exchanged (torch.distributed all_gather_object) and mapped (ngp_ipc_open).
    `peers(name)` returns a ctypes array of `world` device pointers (own rank: the local pointer)."""

    ALIGN = 256

    def __init__(self, n_table, n_w, world_size, rank, group, device="cuda", ipc=True):
        from . import lib
        self.world, self.rank, self.group = world_size, rank, group
        self.P = padded_len(n_table, world_size)
        self.n_w = n_w
        a = self.ALIGN
        up = lambda x: (x + a - 1) // a * a                                   # noqa: E731
        self.off = {"table": 0, "table_grad": up(self.P * 2)}
        self.off["w_grad"] = self.off["table_grad"] + up(self.P * 2)
        self.off["flags"] = self.off["w_grad"] + up((n_w + 8) * 4)
        self.nbytes = self.off["flags"] + 64 * 4
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.table = self.buf[self.off["table"]:self.off["table"] + self.P * 2].view(torch.float16)
        self.table_grad = self.buf[self.off["table_grad"]:self.off["table_grad"] + self.P * 2].view(torch.float16)
        self.w_grad = self.buf[self.off["w_grad"]:self.off["w_grad"] + (n_w + 8) * 4].view(torch.float32)
        self.flags = self.buf[self.off["flags"]:self.off["flags"] + 256].view(torch.int32)
        self.grads = self.buf[self.off["table_grad"]:self.off["flags"]]         # both gradient buffers: one memset clears them
        # load the wait kernel now: with lazy module loading its first launch may have to wait for running kernels to drain,
        # which must not happen while an exchange kernel is spinning on a peer (epoch 0 returns immediately)
        lib.call("ngp_dp_exchange_wait", torch.cuda.current_stream().cuda_stream, 1, self.flags.data_ptr(), 0)
        self.base = [None] * world_size
        self._opened = []
        self.base[rank] = self.buf.data_ptr()
        if world_size > 1 and ipc:                                               # ipc=False: the caller fills self.base (single-process tests)
            import ctypes
            import numpy as np
            err = None
            h = np.zeros(64, np.uint8)
            off = ctypes.c_uint64(0)
            try:
                lib.call("ngp_ipc_export", self.buf.data_ptr(), h.ctypes.data, ctypes.addressof(off))
            except lib.NgpError as e:
                err = e
            mine = (h.tobytes(), int(off.value), err is None)
            everyone = [None] * world_size
            dist.all_gather_object(everyone, mine, group=group)
            for r, (hb, o, good) in enumerate(everyone):
                if r == rank or err is not None or not good:
                    continue
                hr = np.frombuffer(hb, np.uint8).copy()
                out = ctypes.c_void_p(0)
                try:
                    lib.call("ngp_ipc_open", hr.ctypes.data, o, ctypes.addressof(out))
                except lib.NgpError as e:
                    err = e
                    continue
                self.base[r] = int(out.value)
                self._opened.append((int(out.value), o))
            # every rank learns whether EVERY mapping succeeded (so that all ranks fall back together); also: every arena is
            # zeroed and mapped before the first epoch
            good = torch.tensor([int(err is None and all(g for _, _, g in everyone))], device=device)
            dist.all_reduce(good, op=dist.ReduceOp.MIN, group=group)
            if not bool(good.item()):
                self.close()
                raise RuntimeError(f"peer mapping failed on at least one rank ({err})")

    def peers(self, name):
        import ctypes
        return (ctypes.c_void_p * self.world)(*[b + self.off[name] for b in self.base])

    def close(self):
        from . import lib
        for p, o in self._opened:
            lib.call("ngp_ipc_close", p, o)
        self._opened = []


def peer_exchange_available(world_size):
    """True when every visible GPU pair can map each other's memory (NVLink / NVSwitch box)."""
    if world_size <= 1 or not torch.cuda.is_available() or torch.cuda.device_count() < world_size:
        return False
    me = torch.cuda.current_device()
    return all(torch.cuda.can_device_access_peer(me, d) for d in range(torch.cuda.device_count()) if d != me)
