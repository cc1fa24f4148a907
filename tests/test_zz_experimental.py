"""Opt-in kernel variants that have NOT been run on a GPU yet (written after the round's GPU budget was spent).  Skipped unless
NGP_EXPERIMENTAL=1; the default product path does not use them."""
import numpy as np
import pytest
import torch

from conftest import experimental

pytestmark = [pytest.mark.gpu, experimental("saved-activation fused network pair")]


def test_saved_activation_pair_equals_recompute_pair():
    """ngp_network_fwd_saved / ngp_network_bwd_saved == ngp_network_fwd / ngp_network_bwd: same outputs bit for bit, same gradients up
    to the order of the atomic accumulations (the activation image holds the fp16 values the recomputation produces)."""
    from jnerf_b200 import ops
    rng = np.random.default_rng(3)
    for n in (1000, 128 * 300 + 17):
        lv = ops.HashLevels(1)
        coords = torch.from_numpy(rng.random((n, 7), dtype=np.float32)).cuda()
        grid = torch.from_numpy(rng.uniform(-1, 1, lv.n_params).astype(np.float16)).cuda()
        wd = torch.from_numpy(rng.uniform(-0.25, 0.25, 3072).astype(np.float16)).cuda()
        wr = torch.from_numpy(rng.uniform(-0.25, 0.25, 7168).astype(np.float16)).cuda()
        dout = torch.from_numpy((rng.standard_normal((n, 4)) * 0.05).astype(np.float16)).cuda()
        n_dev = torch.tensor([n - 5], dtype=torch.int32, device="cuda")          # live rows < buffer rows, last tile ragged
        out0, enc0 = ops.network_fwd(coords, grid, lv, wd, wr, n_dev=n_dev)
        act = ops.network_act_buffer(n)
        out1, enc1 = ops.network_fwd_saved(coords, grid, lv, wd, wr, act, n_dev=n_dev)
        torch.cuda.synchronize()
        assert ops.lib.load().ngp_debug_timeout_flag() == 0
        assert torch.equal(out0[:n - 5], out1[:n - 5]) and torch.equal(enc0[:n - 5], enc1[:n - 5])
        res = []
        for saved in (False, True):
            gg = torch.zeros(lv.n_params, dtype=torch.float16, device="cuda")
            dwd, dwr = torch.zeros(3072, device="cuda"), torch.zeros(7168, device="cuda")
            if saved:
                ops.network_bwd_saved(coords, enc1, act, lv, wd, wr, dout, gg, dwd, dwr, n_dev=n_dev)
            else:
                ops.network_bwd(coords, enc0, lv, wd, wr, dout, gg, dwd, dwr, n_dev=n_dev)
            torch.cuda.synchronize()
            assert ops.lib.load().ngp_debug_timeout_flag() == 0
            res.append((gg.float(), dwd, dwr))
        for a, b, tol in zip(res[0], res[1], (2e-2, 1e-4, 1e-4)):               # fp16 atomics / fp32 atomics: order only
            scale = float(a.abs().max())
            assert scale > 0 and float((a - b).abs().max()) <= tol * scale, (float((a - b).abs().max()), scale)


def test_training_with_saved_activations_matches_default(monkeypatch):
    from test_gpu_runner import make_runner
    ra = make_runner(seed=3)
    la = [float(ra.train_step().mean()) for _ in range(20)]
    monkeypatch.setenv("NGP_SAVE_ACT", "1")
    rb = make_runner(seed=3)
    assert rb.save_act
    lb = [float(rb.train_step().mean()) for _ in range(20)]
    assert abs(la[0] - lb[0]) <= 1e-6 * max(1.0, abs(la[0]))                    # identical forward
    assert abs(la[-1] - lb[-1]) <= 5e-2 * max(abs(la[-1]), 1e-3)                # same trajectory up to atomic-order noise


def test_pipelined_march_count_is_still_bit_exact():
    """NGP_MARCH_PIPE=1 (count pass computes the next chunk's t values under the latency of the occupancy lookup): the same bit-exact
    march comparisons as the default kernel.  The switch is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, NGP_MARCH_PIPE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_ops.py"), "-q", "-x", "-m", "gpu", "-k", "march", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]


def test_pipelined_composite_matches_oracle():
    """NGP_COMPOSITE_PIPE=1 (fused composite + Huber + backward with the next chunk's loads in flight): the same oracle comparisons
    as the default kernel (test_composite's fp16 branch covers ngp_composite_loss_bwd) plus the fused-step-vs-autograd check."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, NGP_COMPOSITE_PIPE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_ops.py"), os.path.join(here, "test_gpu_runner.py"), "-q", "-x", "-m", "gpu",
                        "-k", "test_composite or fused_step", "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "3 passed" in r.stdout, r.stdout[-2000:]


def test_run_length_standalone_hash_kernels_match_oracle():
    """NGP_HASH_RUNLEN=1 (standalone ngp_hash_fwd / ngp_hash_bwd walking 16 consecutive points per thread with the corner values /
    fp32 corner accumulators kept while the grid cell does not change): the oracle comparisons of the default kernels, plus the
    per-operator autograd step, which drives them on ray-ordered samples."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, NGP_HASH_RUNLEN="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_ops.py"), os.path.join(here, "test_gpu_runner.py"), "-q", "-x", "-m", "gpu",
                        "-k", "test_hash or fused_step or api_surface", "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]
