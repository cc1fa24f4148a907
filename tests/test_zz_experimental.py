"""Opt-in kernel variants still awaiting a promote-or-delete decision.  Skipped unless NGP_EXPERIMENTAL=1; the default product path
does not use them."""
import numpy as np
import pytest
import torch

from conftest import experimental

pytestmark = [pytest.mark.gpu, experimental("opt-in kernel variants")]


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
