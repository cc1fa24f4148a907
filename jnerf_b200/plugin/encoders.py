"""HashEncoder / GridEncode / SHEncoder: the host-side mirror of JNeRF's encoder plugins
(models/position_encoders/hash_encoder/{hash_encoder,grid_encode}.py, sh_encoder/sh_encoder.py).
Same registry names, constructor arguments, attributes (`m_grid`, `out_dim`) and `execute` signatures; the
bodies call libngp_b200's C ABI instead of `jt.code` kernel strings."""
import torch

from .. import ops
from ..utils.config import get_cfg
from ..utils.registry import ENCODERS
from .module import Module

DEFAULT_HASH = "p0 ^ p1 * 19349663 ^ p2 * 83492791"


def parse_hash_func(expr):
    """cfg.hash_func -> (prime0, prime1, prime2).  The reference pastes the string into its kernel source as
    `#define get_index(p0,p1,p2) <hash_func>` (HE/hash_encoder.py:13-16); the kernels here take the XOR-of-products family the NGP
    configs use -- `p0 [* a] ^ p1 [* b] ^ p2 [* c]` in any order, unsigned 32-bit arithmetic -- as three multipliers in the level table."""
    import re
    primes = {}
    for term in "".join(str(expr).split()).split("^"):
        m = re.fullmatch(r"p([012])(?:\*(\d+)[uU]?)?|(\d+)[uU]?\*p([012])", term)
        if not m:
            raise NotImplementedError(f"hash_func {expr!r}: term {term!r} is not of the form pK or pK * N")
        k, n = (int(m.group(1)), m.group(2)) if m.group(1) is not None else (int(m.group(4)), m.group(3))
        if k in primes:
            raise NotImplementedError(f"hash_func {expr!r}: p{k} appears twice")
        primes[k] = int(n) & 0xFFFFFFFF if n is not None else 1
    if sorted(primes) != [0, 1, 2]:
        raise NotImplementedError(f"hash_func {expr!r}: every one of p0, p1, p2 must appear exactly once")
    return primes[0], primes[1], primes[2]


class _GridEncodeFn(torch.autograd.Function):
    """GridEncode.execute / .grad (HE/grid_encode.py:66-190): returns (None, grid_gradient) -- no dL/dx."""

    @staticmethod
    def forward(ctx, x, m_grid, levels):
        ctx.save_for_backward(x)
        ctx.levels = levels
        return ops.hash_fwd(x.contiguous(), m_grid, levels)

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        g = ops.hash_bwd(x.contiguous(), grad_out.contiguous(), ctx.levels)
        return None, g, None


class GridEncode:
    """HE/grid_encode.py:11-64.  The level table replaces `m_hashmap_offsets_table`; no scratch buffers are kept
    (the reference pre-allocates 100 MB + 537 MB of SoA scratch, :55-61)."""

    def __init__(self, hash_func_header, aabb_scale=1, n_pos_dims=3, n_features_per_level=2, n_levels=16, base_resolution=16,
                 log2_hashmap_size=19, n_rays_per_batch=4096, MAX_STEP=1024, using_fp16=False):
        assert n_pos_dims == 3 and n_features_per_level == 2 and n_levels == 16, "kernels are specialised for 3D, F=2, L=16"
        # hash_func_header: the reference's "#define get_index(p0,p1,p2) <expr>" (or just <expr>, or empty for the default)
        expr = hash_func_header.split(")", 1)[1] if "get_index" in hash_func_header else hash_func_header
        self.levels = ops.HashLevels(aabb_scale, n_levels, base_resolution, log2_hashmap_size, primes=parse_hash_func(expr.strip() or DEFAULT_HASH))
        self.m_n_params = self.levels.n_params
        self.m_per_level_scale = self.levels.per_level_scale
        self.m_n_levels = n_levels
        self.m_base_resolution = base_resolution
        self.grad_type = torch.float16 if using_fp16 else torch.float32

    def __call__(self, x, m_grid):
        assert m_grid.dtype == self.grad_type
        return _GridEncodeFn.apply(x, m_grid, self.levels)


@ENCODERS.register_module()
class HashEncoder(Module):
    """HE/hash_encoder.py:7-30 (ignores its own arguments and hard-codes L=16, F=2, base 16, T=2^19, like the reference)."""

    def __init__(self, n_pos_dims=3, n_features_per_level=2, n_levels=16, base_resolution=16, log2_hashmap_size=19):
        super().__init__()
        self.cfg = get_cfg()
        using_fp16 = bool(self.cfg.fp16)
        aabb_scale = self.cfg.dataset_obj.aabb_scale if self.cfg.dataset_obj is not None else 1
        self.hash_func = self.cfg.hash_func or DEFAULT_HASH
        self.hash_func_header = f"#define get_index(p0,p1,p2) {self.hash_func}"
        self.encoder = GridEncode(self.hash_func_header, aabb_scale=aabb_scale, n_pos_dims=3, n_features_per_level=2, n_levels=16, base_resolution=16,
                                  log2_hashmap_size=19, using_fp16=using_fp16)
        self.grad_type = torch.float16 if using_fp16 else torch.float32
        g = torch.Generator(device="cuda").manual_seed(int(self.cfg.seed or 1))
        init = torch.rand(self.encoder.m_n_params, device="cuda", generator=g) * 2e-4 - 1e-4        # jt.init.uniform(-1e-4, 1e-4)
        self.m_grid = torch.nn.Parameter(init.to(self.grad_type))
        self.out_dim = n_features_per_level * n_levels

    @property
    def levels(self):
        return self.encoder.levels

    def execute(self, x):
        assert self.m_grid.dtype == self.grad_type
        return self.encoder(x, self.m_grid)


@ENCODERS.register_module()
class SHEncoder(Module):
    """SH/sh_encoder.py:9-56: degree-4 spherical harmonics of a direction in [0,1]^3, 16 outputs, no gradient."""

    def __init__(self):
        super().__init__()
        self.cfg = get_cfg()
        self.grad_type = torch.float16 if self.cfg.fp16 else torch.float32
        self.m_sh_degree = 4
        self.out_dim = 16

    def execute(self, x):
        with torch.no_grad():
            return ops.sh_fwd(x.contiguous(), self.grad_type)
