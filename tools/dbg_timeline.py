import sys, ctypes as C, numpy as np, torch
sys.path.insert(0,__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from jnerf_b200 import ops, lib
N=262144
wr=(torch.rand(7168,device='cuda')-0.5).half(); X=torch.randn(N,32,device='cuda').half()
for _ in range(3): ops.mlp_fwd(wr,X,1)
torch.cuda.synchronize()
h=np.zeros(64,np.int64); l=lib.load(); l.ngp_debug_read_timeline.argtypes=[C.c_void_p]; l.ngp_debug_read_timeline(h.ctypes.data)
names=['start','x_loaded','sync1','issued','waited','epi_done','sync2','pre_out_issue','out_waited','tile_end']
print('mlp_fwd tile timeline (cycles, delta):')
for i in range(1,10): print(f'  {names[i]:14s} +{h[i]-h[i-1]:6d}   (t={h[i]-h[0]})')
