"""Per-stage clock64 timeline of one tile of the fused backward kernel (build with NGP_NVCC_FLAGS=-DNGP_TIMELINE)."""
import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jnerf_b200 import ops, lib
N = 262144
dev = "cuda"
lv = ops.HashLevels(1)
grid = (torch.rand(lv.n_params, device=dev) * 2e-4 - 1e-4).half()
coords = torch.rand(N, 7, device=dev)
wd = (torch.rand(3072, device=dev) - 0.5).half(); wr = (torch.rand(7168, device=dev) - 0.5).half()
out, enc = ops.network_fwd(coords, grid, lv, wd, wr)
dout = (torch.randn(N, 4, device=dev) * 1e-3).half()
gg = torch.zeros(lv.n_params, dtype=torch.float16, device=dev); dwd = torch.zeros(3072, device=dev); dwr = torch.zeros(7168, device=dev)
for _ in range(3): ops.network_bwd(coords, enc, lv, wd, wr, dout, gg, dwd, dwr)
torch.cuda.synchronize()
h = np.zeros(64, np.int64); l = lib.load(); l.ngp_debug_read_timeline_bwd.argtypes = [C.c_void_p]; l.ngp_debug_read_timeline_bwd(h.ctypes.data)
names = ['start', 'loads issued', 'sync', 'fwd chain (4 stages)', 'B1 issued', 'B1 waited', 'B1 epi', 'B1 sync', 'B2 issued', 'B2 waited', 'B2 epi+sync', 'B3..B5', 'scatter']
for i in range(1, 13): print(f'  {names[i]:22s} +{h[i]-h[i-1]:7d}   (t={h[i]-h[0]})')
print('B2 stage in detail (thread 0 of CTA 0; commit -> flip includes the MMA execution):')
fine = [('B1 sync done -> 4 dgrad MMAs issued', 7, 8), ('tcgen05.commit issued', 8, 21), ('commit -> mbarrier flip observed (pipe.wait)', 21, 9),
        ('4 x tcgen05.ld issued', 9, 16), ('tcgen05.wait::ld', 16, 17), ('convert + ReLU mask (8 LDS) + 8 STS', 17, 18),
        ('tcgen05.fence + fence.proxy.async', 18, 19), ('named barrier (128 threads)', 19, 20)]
for name, a, b in fine: print(f'  {name:48s} {h[b]-h[a]:7d}')
