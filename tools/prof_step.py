"""Profile target: brings a small lego run to steady state, then executes a few training steps between
cudaProfilerStart/Stop (use with `ncu --profile-from-start off ...`)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jnerf_b200 import plugin  # noqa: F401,E402
from jnerf_b200.runner import Runner, lego_cfg  # noqa: E402
from jnerf_b200.utils.config import get_cfg, update_cfg  # noqa: E402

pre = int(sys.argv[1]) if len(sys.argv) > 1 else 96
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
update_cfg(**lego_cfg(fp16=True, synthetic=True, seed=1))
cfg = get_cfg()
cfg.dataset.train.n_images = 20
cfg.dataset.train.H = cfg.dataset.train.W = 400
cfg.dataset.val = None
r = Runner()
for _ in range(pre):
    r.train_step()
while cfg.m_training_step % 16 != 1:        # keep the grid update out of the captured steps
    r.train_step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(steps):
    r.train_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("rays/iter", r.sampler.n_rays_per_batch, "loss", float(r.last_loss.mean()))
