"""Oracle (C restatement) against the committed golden vectors, which were produced by executing the reference's
own kernel sources (tests/golden/make_golden.py).  CPU only; runs here and on the GPU box."""
import os
import sys
import numpy as np
import pytest
import oracle_lib as ol

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import digest, inputs, table  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ngp_golden.npz"))
INP = inputs()


@pytest.fixture(autouse=True)
def host_arith():
    ol.oracle().orc_set_fma_mode(0)      # goldens come from the host build of the reference (no FMA contraction)
    yield
    ol.oracle().orc_set_fma_mode(1)


def same(a, key):
    assert np.array_equal(digest(a), G[key]), key


def test_config1_hash_forward_and_backward():
    cfg = ol.HashCfg(1, log2_hashmap_size=14)
    assert np.array_equal(cfg.offsets, G["cfg1_offsets"]) and cfg.n_entries == 245640
    for name, dt in (("f32", np.float32), ("f16", np.float16)):
        grid = table(cfg, dt)
        out = ol.hash_fwd(cfg, INP["x"], grid)
        assert np.array_equal(out[:256], G[f"hash_fwd_{name}_head"])
        same(out, f"hash_fwd_{name}_sha")
        gg = ol.hash_bwd(cfg, INP["x"], INP["dy"].astype(dt))
        assert np.array_equal(gg[:8192], G[f"hash_bwd_{name}_lvl0"])
        same(gg, f"hash_bwd_{name}_sha")


def test_sh():
    out = ol.sh(INP["dirs"], np.float32)
    assert np.array_equal(out[:128], G["sh_f32"])
    same(out, "sh_f32_sha")


def test_march_compact_composite():
    bits, _ = ol.sphere_bitfield(0.3)
    same(bits, "bitfield_sha")
    for const_dt, aabb in ((True, (0.0, 1.0)), (False, (-1.5, 2.5))):
        tag = "constdt" if const_dt else "cone"
        coords, _, numsteps, cnt = ol.march(INP["rays_o"], INP["rays_d"], bits, aabb=aabb, const_dt=const_dt, max_samples=300 * 1024)
        assert np.array_equal(numsteps, G[f"march_{tag}_numsteps"]) and np.array_equal(cnt, G[f"march_{tag}_counters"])
        S = int(cnt[1])
        same(coords[:S], f"march_{tag}_coords_sha")
        if not const_dt:
            continue
        cc, ns_c, ccnt = ol.compact(coords, numsteps, S - 777)
        assert np.array_equal(ns_c, G["compact_numsteps"]) and np.array_equal(ccnt, G["compact_counters"])
        same(cc, "compact_coords_sha")
        rng = np.random.default_rng(8)
        net = rng.standard_normal((S - 777, 4)).astype(np.float32)
        bg = rng.random((300, 3), dtype=np.float32)
        lg = rng.standard_normal((300, 3)).astype(np.float32)
        for name, dt in (("f32", np.float32), ("f16", np.float16)):
            rgb = ol.composite_fwd(net.astype(dt), cc, numsteps, ns_c, bg)
            assert np.array_equal(rgb, G[f"comp_{name}_rgb"])
            dnet = ol.composite_bwd(net.astype(dt), cc, ns_c, lg, rgb, 0.001)
            same(dnet, f"comp_{name}_dnet_sha")
            rgbi, alpha = ol.composite_infer(net.astype(dt), cc, ns_c)
            assert np.array_equal(rgbi, G[f"comp_{name}_rgbi"]) and np.array_equal(alpha, G[f"comp_{name}_alpha"])


def test_grid_maintenance():
    si = ol.pcg32_seed()
    assert np.array_equal(si, G["pcg32_seed1337"])
    rng = np.random.default_rng(9)
    n_el = ol.G3 * 5
    g_in = np.where(rng.random(n_el) < 0.3, rng.random(n_el) * 0.05, -1.0).astype(np.float32)
    n = 20000
    pos, idx = ol.generate_grid_samples(n, si, 3, (-1.5, 2.5), g_in, 3, 0.01)
    same(idx, "gridgen_idx_sha")
    same(pos, "gridgen_pos_sha")
    mlp = rng.standard_normal(n).astype(np.float32)
    tmp = np.zeros(n_el, np.float32)
    ol.splat(idx, mlp, tmp)
    ol.ema(g_in, tmp)
    same(g_in, "grid_after_ema_sha")
    mean = ol.grid_mean(g_in)
    assert np.float32(mean) == G["grid_mean"]
    same(ol.update_bitfield(g_in, mean), "grid_bitfield_sha")


def test_mlp_restatement_self_consistency():
    """The MLP is parity-unpinned (binary-only in the reference): check the restatement against float64 numpy."""
    rng = np.random.default_rng(4)
    n = 300
    for nhm, n_valid in ((0, 16), (1, 3)):
        shapes = [(64, 32)] + [(64, 64)] * nhm + [(16, 64)]
        Ws = [rng.uniform(-0.3, 0.3, s).astype(np.float16) for s in shapes]
        W = np.concatenate([w.ravel() for w in Ws])
        X = rng.standard_normal((n, 32)).astype(np.float16)
        Y, inter = ol.mlp_fwd(W, X, nhm)
        a = X.astype(np.float64)
        for li, w in enumerate(Ws[:-1]):
            a = np.maximum(a @ w.astype(np.float64).T, 0).astype(np.float16).astype(np.float64)
            assert np.abs(a - inter[li * n:(li + 1) * n]).max() < 2e-2
        yref = a @ Ws[-1].astype(np.float64).T
        assert np.abs(yref - Y).max() < 2e-2
        dY = rng.standard_normal((n, 16)).astype(np.float16)
        dY[:, n_valid:] = 0
        dX, temps, dW = ol.mlp_bwd(W, X, inter, dY, nhm, n_valid)
        # finite-difference-free check: dW of the output layer = dY^T h_last
        h_last = inter[nhm * n:(nhm + 1) * n].astype(np.float64)
        dWo = dY.astype(np.float64).T @ h_last
        off = 64 * 32 + nhm * 64 * 64
        assert np.abs(dW[off:].reshape(16, 64) - dWo).max() < 1e-3 * max(1, np.abs(dWo).max())
        g_last = (dY.astype(np.float64) @ Ws[-1].astype(np.float64)) * (h_last > 0)
        assert np.abs(temps[:n] - g_last).max() < 2e-2


def test_adam_ema_restatement():
    rng = np.random.default_rng(5)
    n = 1000
    p = rng.uniform(-1e-4, 1e-4, n).astype(np.float32)
    master = p.copy()
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    pr, mr, vr, er = p.astype(np.float64), np.zeros(n), np.zeros(n), p.astype(np.float64)
    for step in range(1, 6):
        g = rng.standard_normal(n).astype(np.float32) * 1e-3
        ol.adam_ema(p, g, m, v, master, 0.1, step)
        mr = 0.9 * mr + 0.1 * g
        vr = 0.99 * vr + 0.01 * g.astype(np.float64) ** 2
        pa = pr - mr * (0.1 * np.sqrt(1 - 0.99 ** step) / (1 - 0.9 ** step)) / (np.sqrt(vr) + 1e-15)
        pr = ((1 - 0.95) * pa + 0.95 * er * (1 - 0.95 ** (step - 1))) / (1 - 0.95 ** step)
        er = pr
        assert np.abs(p - pr).max() < 1e-5
