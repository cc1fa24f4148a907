"""Functional wrappers: torch tensors in, C-ABI calls (include/ngp_b200.h) on the current CUDA stream.
torch provides device memory and streams only; every op below runs in libngp_b200.so."""
import numpy as np
import torch

from . import lib

F32, F16 = 0, 1


def _dt(t):
    if t.dtype == torch.float16:
        return F16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"unsupported dtype {t.dtype}")


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor required"
    return t.data_ptr()


# The raw handle of torch's current stream.  torch.cuda.current_stream() builds a Stream object and re-checks the device on every call
# (~3 us; 16 calls a training step were a sixth of the step's host time, tools/host_probe.py): the two C getters are what it wraps.
_raw_stream, _raw_device = getattr(torch._C, "_cuda_getCurrentRawStream", None), getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


class HashLevels:
    """R1: offsets (host, HE/grid_encode.py:17-39) + the device level table the kernels stage in shared memory."""

    def __init__(self, aabb_scale=1, n_levels=16, base_resolution=16, log2_hashmap_size=19, device="cuda", primes=(1, 19349663, 83492791)):
        import ctypes as C
        self.primes = tuple(int(p) & 0xFFFFFFFF for p in primes)
        self.n_levels, self.base_resolution = n_levels, base_resolution
        self.offsets = np.zeros(n_levels + 1, np.uint32)
        pls = C.c_double()
        lib.call("ngp_hash_offsets", float(aabb_scale), n_levels, base_resolution, log2_hashmap_size, self.offsets.ctypes.data, C.addressof(pls))
        self.per_level_scale = pls.value
        self.log2_per_level_scale = float(np.float32(np.log2(self.per_level_scale)))
        self.n_entries = int(self.offsets[-1])
        self.n_params = 2 * self.n_entries
        self.table = torch.empty(n_levels * 32, dtype=torch.uint8, device=device)
        lib.call("ngp_hash_level_table_primes", _stream(), self.offsets.ctypes.data, n_levels, base_resolution, self.log2_per_level_scale,
                 _p(self.table), *self.primes)


def hash_fwd(x, grid, levels):
    out = torch.empty((x.shape[0], 32), dtype=grid.dtype, device=x.device)
    lib.call("ngp_hash_fwd", _stream(), x.shape[0], _p(x), _p(grid), _dt(grid), _p(levels.table), _p(out))
    return out


def hash_bwd(x, dy, levels, grid_grad=None):
    if grid_grad is None:
        # the kernel's memset is skipped for an empty batch (as HE/grid_encode.py:142-144 does): hand out zeros in that case
        alloc = torch.zeros if x.shape[0] == 0 else torch.empty
        grid_grad = alloc(levels.n_params, dtype=dy.dtype, device=x.device)
    lib.call("ngp_hash_bwd", _stream(), x.shape[0], _p(x), _p(dy), _dt(dy), _p(levels.table), _p(grid_grad), levels.n_params)
    return grid_grad


def sh_fwd(dirs, dtype=torch.float16):
    out = torch.empty((dirs.shape[0], 16), dtype=dtype, device=dirs.device)
    lib.call("ngp_sh_fwd", _stream(), dirs.shape[0], _p(dirs), F16 if dtype == torch.float16 else F32, _p(out))
    return out


def mlp_fwd(W, X, n_hidden_matmuls, save_inter=True):
    n = X.shape[0]
    inter = torch.empty(((n_hidden_matmuls + 1) * n, 64), dtype=torch.float16, device=X.device) if save_inter else None
    Y = torch.empty((n, 16), dtype=torch.float16, device=X.device)
    lib.call("ngp_mlp_fwd", _stream(), _p(W), _p(X), _p(inter), _p(Y), n_hidden_matmuls, n)
    return Y, inter


def mlp_bwd(W, X, inter, dY, n_hidden_matmuls, n_out_valid, need_dx=True, need_temps=False):
    n = X.shape[0]
    dX = torch.empty((n, 32), dtype=torch.float16, device=X.device) if need_dx else None
    temps = torch.empty(((n_hidden_matmuls + 1) * n, 64), dtype=torch.float16, device=X.device) if need_temps else None
    dW = torch.empty(W.numel(), dtype=torch.float32, device=X.device)
    lib.call("ngp_mlp_bwd", _stream(), _p(W), _p(X), _p(inter), _p(dY), _p(dX), _p(temps), _p(dW), n_hidden_matmuls, n_out_valid, n)
    return dX, temps, dW


def mlp_bwd_dgrad(W, inter, dY_feature_major, n_hidden_matmuls, need_dx=False):
    """Data-gradient chain only (the contract of the reference's link-level mlp_fused_backward_func): dY is (16, n)."""
    n = dY_feature_major.shape[1]
    dX = torch.empty((n, 32), dtype=torch.float16, device=W.device) if need_dx else None
    temps = torch.empty(((n_hidden_matmuls + 1) * n, 64), dtype=torch.float16, device=W.device)
    lib.call("ngp_mlp_bwd_dgrad", _stream(), _p(W), _p(inter), _p(dY_feature_major), _p(dX), _p(temps), n_hidden_matmuls, n)
    return dX, temps


def network_fwd(coords, grid, levels, wd, wr, n_dev=None, save_enc=True, out=None, enc=None):
    n = coords.shape[0]
    if out is None:
        out = torch.empty((n, 4), dtype=torch.float16, device=coords.device)
    if enc is None and save_enc:
        enc = torch.empty((n, 32), dtype=torch.float16, device=coords.device)
    lib.call("ngp_network_fwd", _stream(), n, _p(n_dev), _p(coords), _p(grid), _p(levels.table), _p(wd), _p(wr), _p(out), _p(enc))
    return out, enc


def network_bwd(coords, enc, levels, wd, wr, dout, grid_grad, dwd, dwr, n_dev=None):
    lib.call("ngp_network_bwd", _stream(), coords.shape[0], _p(n_dev), _p(coords), _p(enc), _p(levels.table), _p(wd), _p(wr), _p(dout),
             _p(grid_grad), _p(dwd), _p(dwr))


def density_fwd(pos, grid, levels, wd):
    out = torch.empty(pos.shape[0], dtype=torch.float16, device=pos.device)
    lib.call("ngp_density_fwd", _stream(), pos.shape[0], _p(pos), _p(grid), _p(levels.table), _p(wd), _p(out))
    return out


def march(rays_o, rays_d, bitfield, aabb, max_samples, cone_angle, near, cascades, const_dt, rng, coords=None, workspace=None):
    R = rays_o.shape[0]
    dev = rays_o.device
    counters = torch.empty(2, dtype=torch.int32, device=dev)
    ray_idx = torch.zeros(R, dtype=torch.int32, device=dev)
    numsteps = torch.empty((R, 2), dtype=torch.int32, device=dev)
    if coords is None:
        coords = torch.empty((max_samples, 7), dtype=torch.float32, device=dev)
    if workspace is None:
        workspace = torch.empty(int(lib.load().ngp_march_workspace_bytes(R)), dtype=torch.uint8, device=dev)
    lib.call("ngp_march", _stream(), R, float(aabb[0]), float(aabb[1]), max_samples, _p(rays_o), _p(rays_d), _p(bitfield), float(cone_angle),
             float(near), cascades, int(const_dt), int(rng[0]), int(rng[1]), _p(counters), _p(ray_idx), _p(numsteps), _p(coords), _p(workspace))
    return coords, ray_idx, numsteps, counters


def compact(coords, numsteps, max_compacted, alias=False, zero_fill=True):
    R = numsteps.shape[0]
    dev = coords.device
    out = coords if alias else torch.empty((max_compacted, 7), dtype=torch.float32, device=dev)
    ns = torch.empty((R, 2), dtype=torch.int32, device=dev)
    counters = torch.empty(2, dtype=torch.int32, device=dev)
    lib.call("ngp_compact", _stream(), R, max_compacted, _p(coords), _p(numsteps), _p(out), _p(ns), _p(counters), int(zero_fill))
    return out, ns, counters


def composite_fwd(net, coords, numsteps_in, numsteps_c, bg, cascades=5):
    R = numsteps_c.shape[0]
    rgb = torch.empty((R, 3), dtype=torch.float32, device=net.device)
    lib.call("ngp_composite_fwd", _stream(), R, _p(net), _dt(net), _p(coords), _p(numsteps_in), _p(numsteps_c), _p(bg), cascades, _p(rgb))
    return rgb


def composite_bwd(net, coords, numsteps_c, loss_grad, rgb_ray, mean, cascades=5):
    R = numsteps_c.shape[0]
    dnet = torch.empty_like(net)
    lib.call("ngp_composite_bwd", _stream(), R, net.shape[0], _p(net), _dt(net), _p(coords), _p(numsteps_c), _p(loss_grad), _p(rgb_ray), _p(mean),
             cascades, _p(dnet))
    return dnet


def composite_infer(net, coords, numsteps, cascades=5):
    R = numsteps.shape[0]
    rgb = torch.empty((R, 3), dtype=torch.float32, device=net.device)
    alpha = torch.empty((R, 1), dtype=torch.float32, device=net.device)
    lib.call("ngp_composite_infer", _stream(), R, _p(net), _dt(net), _p(coords), _p(numsteps), cascades, _p(rgb), _p(alpha))
    return rgb, alpha


def composite_loss_bwd(net, coords, numsteps_in, numsteps_c, bg, target, mean, delta=0.1, cascades=5, dnet=None, rgb=None, loss=None, reg_scale=1.0):
    R = numsteps_c.shape[0]
    dev = net.device
    if dnet is None:
        dnet = torch.zeros_like(net)
    if rgb is None:
        rgb = torch.empty((R, 3), dtype=torch.float32, device=dev)
    if loss is None:
        loss = torch.empty(R, dtype=torch.float32, device=dev)
    lib.call("ngp_composite_loss_bwd", _stream(), R, net.shape[0], _p(net), _p(coords), _p(numsteps_in), _p(numsteps_c), _p(bg), _p(target),
             float(delta), _p(mean), cascades, _p(rgb), _p(loss), _p(dnet), float(reg_scale))
    return rgb, loss, dnet


def grid_mark_untrained(grid, focal, xforms, res):
    lib.call("ngp_grid_mark_untrained", _stream(), grid.numel(), _p(grid), xforms.shape[0], _p(focal), _p(xforms), int(res[0]), int(res[1]))


def grid_generate_samples(n, rng, step_dev, aabb, grid, n_cascades, thresh):
    pos = torch.empty((n, 3), dtype=torch.float32, device=grid.device)
    idx = torch.empty(n, dtype=torch.int32, device=grid.device)
    lib.call("ngp_grid_generate_samples", _stream(), n, int(rng[0]), int(rng[1]), _p(step_dev), float(aabb[0]), float(aabb[1]), _p(grid), _p(pos),
             _p(idx), n_cascades, float(thresh))
    return pos, idx


def grid_splat(indices, mlp_out, grid_tmp):
    lib.call("ngp_grid_splat", _stream(), indices.numel(), _p(indices), _p(mlp_out), _dt(mlp_out), _p(grid_tmp))


def grid_ema(grid, grid_tmp, decay=0.95):
    lib.call("ngp_grid_ema", _stream(), grid.numel(), float(decay), _p(grid), _p(grid_tmp))


def grid_update_bitfield(grid, mean, bitfield, cascades=5):
    lib.call("ngp_grid_update_bitfield", _stream(), _p(grid), _p(mean), _p(bitfield), cascades)


def adam_ema(param, grad, m, v, master, lr, step, beta1=0.9, beta2=0.99, eps=1e-15, ema_decay=0.95, grad_scale=1.0, zero_grad=True):
    lib.call("ngp_adam_ema", _stream(), param.numel(), _p(param), _dt(param), _p(grad), _dt(grad), float(grad_scale), _p(m), _p(v), _p(master),
             float(lr), float(beta1), float(beta2), float(eps), int(step), float(ema_decay), int(zero_grad))


def raygen(pix, W, H, xforms, focal, principal):
    n = pix.numel()
    dev = pix.device
    img = torch.empty(n, dtype=torch.int32, device=dev)
    o = torch.empty((n, 3), dtype=torch.float32, device=dev)
    d = torch.empty((n, 3), dtype=torch.float32, device=dev)
    lib.call("ngp_raygen", _stream(), n, _p(pix), W, H, _p(xforms), _p(focal), _p(principal), _p(img), _p(o), _p(d))
    return img, o, d


def prepare_batch(pix, W, H, xforms, focal, principal, images, bg):
    n = pix.numel()
    dev = pix.device
    img = torch.empty(n, dtype=torch.int32, device=dev)
    o = torch.empty((n, 3), dtype=torch.float32, device=dev)
    d = torch.empty((n, 3), dtype=torch.float32, device=dev)
    target = torch.empty((n, 3), dtype=torch.float32, device=dev)
    lib.call("ngp_prepare_batch", _stream(), n, _p(pix), W, H, _p(xforms), _p(focal), _p(principal), _p(images), int(images.dtype == torch.uint8),
             _p(bg), _p(img), _p(o), _p(d), _p(target))
    return img, o, d, target


def blend_target(rgba, bg, target=None):
    """target = rgb*a + bg*(1-a) (runner.py:68) for an (n,4) f32 RGBA batch."""
    n = rgba.shape[0]
    if target is None:
        target = torch.empty((n, 3), dtype=torch.float32, device=rgba.device)
    lib.call("ngp_blend_target", _stream(), n, _p(rgba), _p(bg), _p(target))
    return target


# ---- device-resident step state (include/ngp_b200.h: ngp_step_state_*) -- what makes a training step replayable as a CUDA graph ----
def step_state_new(device="cuda"):
    return torch.zeros(int(lib.load().ngp_step_state_bytes()), dtype=torch.uint8, device=device)


def step_state_set(state, rng, pix_cursor, adam_steps_done, lr, beta1=0.9, beta2=0.99, eps=1e-15, ema_decay=0.95, grad_scale=1.0):
    lib.call("ngp_step_state_set", _stream(), _p(state), int(rng[0]), int(rng[1]), int(pix_cursor), int(adam_steps_done), float(lr), float(beta1),
             float(beta2), float(eps), float(ema_decay), float(grad_scale))


def step_state_tick(state, pix_advance, lr, beta1=0.9, beta2=0.99, eps=1e-15, ema_decay=0.95, grad_scale=1.0):
    lib.call("ngp_step_state_tick", _stream(), _p(state), int(pix_advance), float(lr), float(beta1), float(beta2), float(eps), float(ema_decay),
             float(grad_scale))


def prepare_batch_dev(n, pix_list, state, pix_offset, W, H, xforms, focal, principal, images, bg):
    dev = pix_list.device
    img = torch.empty(n, dtype=torch.int32, device=dev)
    o = torch.empty((n, 3), dtype=torch.float32, device=dev)
    d = torch.empty((n, 3), dtype=torch.float32, device=dev)
    target = torch.empty((n, 3), dtype=torch.float32, device=dev)
    lib.call("ngp_prepare_batch_dev", _stream(), n, _p(pix_list), _p(state), int(pix_offset), W, H, _p(xforms), _p(focal), _p(principal), _p(images),
             int(images.dtype == torch.uint8), _p(bg), _p(img), _p(o), _p(d), _p(target))
    return img, o, d, target


def march_dev(rays_o, rays_d, bitfield, aabb, max_samples, cone_angle, near, cascades, const_dt, state, ray_offset=0, coords=None, workspace=None):
    R = rays_o.shape[0]
    dev = rays_o.device
    counters = torch.empty(2, dtype=torch.int32, device=dev)
    ray_idx = torch.zeros(R, dtype=torch.int32, device=dev)
    numsteps = torch.empty((R, 2), dtype=torch.int32, device=dev)
    if coords is None:
        coords = torch.empty((max_samples, 7), dtype=torch.float32, device=dev)
    if workspace is None:
        workspace = torch.empty(int(lib.load().ngp_march_workspace_bytes(R)), dtype=torch.uint8, device=dev)
    lib.call("ngp_march_dev", _stream(), R, float(aabb[0]), float(aabb[1]), max_samples, _p(rays_o), _p(rays_d), _p(bitfield), float(cone_angle),
             float(near), cascades, int(const_dt), _p(state), int(ray_offset), _p(counters), _p(ray_idx), _p(numsteps), _p(coords), _p(workspace))
    return coords, ray_idx, numsteps, counters


def adam_ema_dev(param, grad, m, v, master, state, zero_grad=True):
    lib.call("ngp_adam_ema_dev", _stream(), param.numel(), _p(param), _dt(param), _p(grad), _dt(grad), _p(m), _p(v), _p(master), _p(state), int(zero_grad))


def pcg32_seed(seed=1337, seq=1):
    si = np.zeros(2, np.uint64)
    lib.load().ngp_pcg32_seed(seed, seq, si.ctypes.data)
    return si


def pcg32_advance(si, delta=1 << 32):
    lib.load().ngp_pcg32_advance(si.ctypes.data, delta)
    return si


def dp_exchange_step(world, rank, slice_len, n_w, peer_table, peer_table_grad, peer_w_grad, peer_flags, epoch, m, v, master, w_param, w_m, w_v,
                     w_master, lr, step, beta1=0.9, beta2=0.99, eps=1e-15, ema_decay=0.95, grad_scale=1.0):
    """8e: gradient reduce-scatter + Adam/EMA on this rank's table slice + all-gather + MLP-weight all-reduce/update in ONE
    launch over NVLink peer memory.  peer_* are ctypes arrays of `world` device pointers (dp.PeerArena.peers)."""
    import ctypes
    assert m.numel() == slice_len and v.numel() == slice_len and master.numel() == slice_len and w_param.numel() == n_w
    lib.call("ngp_dp_exchange_step", _stream(), int(world), int(rank), int(slice_len), int(n_w), ctypes.addressof(peer_table),
             ctypes.addressof(peer_table_grad), ctypes.addressof(peer_w_grad), ctypes.addressof(peer_flags), int(epoch), _p(m), _p(v), _p(master),
             _p(w_param), _p(w_m), _p(w_v), _p(w_master), float(grad_scale), float(lr), float(beta1), float(beta2), float(eps), int(step),
             float(ema_decay))


def dp_exchange_wait(world, my_flags, epoch):
    lib.call("ngp_dp_exchange_wait", _stream(), int(world), _p(my_flags), int(epoch))
