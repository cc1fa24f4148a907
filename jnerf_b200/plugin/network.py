"""FMLP / FullyFusedMlp_weight / NGPNetworks: mirror of ops/code_ops/fully_fused_mlp.py and
models/networks/ngp_network.py.  `NGPNetworks.execute` runs the whole encoder + MLP stack as ONE kernel
(ngp_network_fwd / ngp_network_bwd); `fused=False` composes the per-operator plugins exactly as the reference does."""
import math

import torch

from .. import ops
from ..utils.config import get_cfg
from ..utils.registry import ENCODERS, NETWORKS, build_from_cfg
from .module import Module


def invariant_uniform(shape, gen):
    """jittor.init.invariant_uniform (mode fan_in+fan_out average): U(+-sqrt(3 / ((fan_in + fan_out) / 2)))."""
    bound = math.sqrt(3.0 / ((shape[0] + shape[1]) / 2.0))
    return (torch.rand(shape, device="cuda", generator=gen) * 2 - 1) * bound


class _FMLPFn(torch.autograd.Function):
    """FullyFusedMlp_weight.execute / .grad (OPS/fully_fused_mlp.py:42-145)."""

    @staticmethod
    def forward(ctx, x, con_weights, n_hidden_matmuls, n_out_valid):
        x = x.contiguous()
        y, inter = ops.mlp_fwd(con_weights, x, n_hidden_matmuls, save_inter=True)
        ctx.save_for_backward(x, con_weights, inter)
        ctx.cfg = (n_hidden_matmuls, n_out_valid)
        return y

    @staticmethod
    def backward(ctx, grads):
        x, w, inter = ctx.saved_tensors
        nhm, n_valid = ctx.cfg
        dx, _, dw = ops.mlp_bwd(w, x, inter, grads.contiguous(), nhm, n_valid, need_dx=True)
        return dx, dw.to(w.dtype), None, None


class FMLP(Module):
    """models/networks/ngp_network.py:8-37.  weight_shapes e.g. [32, 64, 16] or [32, 64, 64, 3]."""

    def __init__(self, weight_shapes, weights=None, gen=None):
        super().__init__()
        assert len(weight_shapes) > 2 and weight_shapes[0] == 32 and all(s == 64 for s in weight_shapes[1:-1]) and weight_shapes[-1] <= 16, \
            "not supported WIDTH: the fused kernels are built for 32 -> 64 (x k) -> <=16"
        self.output_shape1 = weight_shapes[-1]
        self.n_hidden_matmuls = len(weight_shapes) - 3
        con = []
        for i in range(len(weight_shapes) - 1):
            w = invariant_uniform((weight_shapes[i], weight_shapes[i + 1]), gen).half() if weights is None else weights[i]
            if i == len(weight_shapes) - 2 and w.shape[1] < 16:            # pad the last layer to 16 outputs (:26-27)
                w = torch.cat([w, torch.zeros((w.shape[0], 16 - w.shape[1]), dtype=w.dtype, device=w.device)], -1)
            con.append(w.t().contiguous().reshape(-1))                      # (out, in) row-major (:28)
        self.con_weights = torch.nn.Parameter(torch.cat(con).half())

    def execute(self, x):
        if x.shape[0] == 0:
            return torch.empty((0, self.output_shape1), dtype=torch.float16, device=x.device)
        ret = _FMLPFn.apply(x, self.con_weights, self.n_hidden_matmuls, self.output_shape1)
        return ret[:, :self.output_shape1] if self.output_shape1 != ret.shape[1] else ret


class _NetworkFn(torch.autograd.Function):
    """Whole NGPNetworks.execute_ (ngp_network.py:77-84) as one forward kernel and one backward kernel."""

    @staticmethod
    def forward(ctx, coords, m_grid, wd, wr, levels):
        out, enc = ops.network_fwd(coords, m_grid, levels, wd, wr, save_enc=True)
        ctx.save_for_backward(coords, enc, wd, wr)
        ctx.levels = levels
        ctx.n_params = m_grid.numel()
        return out

    @staticmethod
    def backward(ctx, dout):
        coords, enc, wd, wr = ctx.saved_tensors
        gg = torch.zeros(ctx.n_params, dtype=torch.float16, device=coords.device)
        dwd = torch.zeros(wd.numel(), dtype=torch.float32, device=coords.device)
        dwr = torch.zeros(wr.numel(), dtype=torch.float32, device=coords.device)
        ops.network_bwd(coords, enc, ctx.levels, wd, wr, dout.contiguous(), gg, dwd, dwr)
        return None, gg, dwd.half(), dwr.half(), None


@NETWORKS.register_module()
class NGPNetworks(Module):
    """models/networks/ngp_network.py:39-95."""

    def __init__(self, use_fully=True, density_hidden_layer=1, density_n_neurons=64, rgb_hidden_layer=2, rgb_n_neurons=64, fused=True):
        super().__init__()
        self.use_fully = use_fully
        self.cfg = get_cfg()
        self.using_fp16 = bool(self.cfg.fp16)
        self.pos_encoder = build_from_cfg(self.cfg.encoder.pos_encoder, ENCODERS)
        self.dir_encoder = build_from_cfg(self.cfg.encoder.dir_encoder, ENCODERS)
        gen = torch.Generator(device="cuda").manual_seed(int(self.cfg.seed or 1) + 1)
        self.fused = bool(fused and use_fully and self.using_fp16)
        if self.use_fully and self.using_fp16:
            assert self.pos_encoder.out_dim % 16 == 0 and self.dir_encoder.out_dim % 16 == 0
            self.density_mlp = FMLP([self.pos_encoder.out_dim, density_n_neurons, 16], gen=gen)
            self.rgb_mlp = FMLP([self.dir_encoder.out_dim + 16, rgb_n_neurons, rgb_n_neurons, 3], gen=gen)
        else:
            # the reference's own fallback when fp16 / FFMLP is off (ngp_network.py:54-67): library GEMMs
            if self.use_fully and not self.using_fp16:
                print("Warning: FFMLPs only support float16. Automatically use original MLPs instead.")
            L = torch.nn.Linear
            self.density_mlp = torch.nn.Sequential(L(self.pos_encoder.out_dim, density_n_neurons, bias=False), torch.nn.ReLU(),
                                                   L(density_n_neurons, 16, bias=False)).cuda()
            self.rgb_mlp = torch.nn.Sequential(L(self.dir_encoder.out_dim + 16, rgb_n_neurons, bias=False), torch.nn.ReLU(),
                                               L(rgb_n_neurons, rgb_n_neurons, bias=False), torch.nn.ReLU(),
                                               L(rgb_n_neurons, 3, bias=False)).cuda()
            if self.using_fp16:
                self.density_mlp.half()
                self.rgb_mlp.half()

    def execute(self, pos_input, dir_input):
        if self.fused:
            coords = torch.zeros((pos_input.shape[0], 7), dtype=torch.float32, device=pos_input.device)
            coords[:, :3] = pos_input
            coords[:, 4:] = dir_input
            return self.execute_coords(coords)
        return self.execute_(pos_input, dir_input)

    def execute_coords(self, coords):
        """Fused path on NerfCoordinate rows (N,7) -- what the sampler hands over without slicing copies."""
        return _NetworkFn.apply(coords, self.pos_encoder.m_grid, self.density_mlp.con_weights, self.rgb_mlp.con_weights, self.pos_encoder.levels)

    def execute_(self, pos_input, dir_input):
        dir_input = self.dir_encoder(dir_input)
        pos_input = self.pos_encoder(pos_input)
        density = self.density_mlp(pos_input)
        rgb = torch.cat([density, dir_input], -1)
        rgb = self.rgb_mlp(rgb)
        return torch.cat([rgb, density[..., :1]], -1)

    def density(self, pos_input):
        if self.fused:
            with torch.no_grad():
                return ops.density_fwd(pos_input.contiguous(), self.pos_encoder.m_grid, self.pos_encoder.levels,
                                       self.density_mlp.con_weights).unsqueeze(-1)
        density = self.pos_encoder(pos_input)
        return self.density_mlp(density)[:, :1]

    def set_fp16(self):
        pass   # parameters are created in their final dtype
