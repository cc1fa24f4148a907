"""Generates tests/golden/ngp_golden.npz by EXECUTING THE REFERENCE'S OWN KERNEL SOURCES on the host
(oracle/_ref, built from /root/reference by `make -C oracle ref`).  Run in the build container only:
    python tests/golden/make_golden.py
Inputs are regenerated from numpy default_rng seeds stored in the file; outputs are stored as small slices plus
sha256 digests of the full arrays (bit-exact pins)."""
import hashlib
import os
import sys
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import oracle_lib as ol


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def inputs():
    """Deterministic inputs shared by the generator and the tests."""
    d = {}
    rng = np.random.default_rng(0)
    d["x"] = rng.random((4096, 3), dtype=np.float32)            # BASELINE config #1: 4096 random 3D points
    d["dy"] = (np.random.default_rng(2).standard_normal((4096, 32)) * 1e-2).astype(np.float32)
    d["dirs"] = np.random.default_rng(3).random((1024, 3), dtype=np.float32)
    d["rays_o"], d["rays_d"] = ol.random_rays(300, seed=5)
    return d


def table(cfg, dtype):
    return np.random.default_rng(1).uniform(-1e-4, 1e-4, cfg.n_params).astype(dtype)


def main():
    assert ol.ref("ref_hash_cpu") is not None, "build oracle/_ref first (make -C oracle ref)"
    g = {}
    inp = inputs()
    cfg = ol.HashCfg(1, log2_hashmap_size=14)                   # L=16, T=2^14 -> 245 640 entries
    g["cfg1_offsets"] = cfg.offsets
    for name, dt in (("f32", np.float32), ("f16", np.float16)):
        grid = table(cfg, dt)
        out, pos_soa = ol.ref_hash_fwd(cfg, inp["x"], grid)
        g[f"hash_fwd_{name}_head"] = out[:256]
        g[f"hash_fwd_{name}_sha"] = digest(out)
        gg = ol.ref_hash_bwd(cfg, pos_soa, inp["dy"].astype(dt))
        g[f"hash_bwd_{name}_sha"] = digest(gg)
        g[f"hash_bwd_{name}_lvl0"] = gg[:8192]
    r = ol.ref("ref_sampler_cpu_constdt")
    sh = np.empty((1024, 16), np.float32)
    r.ref_sh_f32_constdt(ol._u32(1024), ol._ptr(inp["dirs"]), ol._ptr(sh))
    g["sh_f32"] = sh[:128]
    g["sh_f32_sha"] = digest(sh)
    bits, _ = ol.sphere_bitfield(0.3)
    g["bitfield_sha"] = digest(bits)
    for const_dt, aabb in ((True, (0.0, 1.0)), (False, (-1.5, 2.5))):
        tag = "constdt" if const_dt else "cone"
        coords, ray_idx, numsteps, cnt = ol.ref_march(inp["rays_o"], inp["rays_d"], bits, aabb=aabb, const_dt=const_dt, max_samples=300 * 1024)
        S = int(cnt[1])
        g[f"march_{tag}_numsteps"] = numsteps
        g[f"march_{tag}_counters"] = cnt
        g[f"march_{tag}_coords_sha"] = digest(coords[:S])
        g[f"march_{tag}_coords_head"] = coords[:64]
        if const_dt:
            cc, ns_c, ccnt = ol.ref_compact(coords, numsteps, S - 777)
            g["compact_numsteps"] = ns_c
            g["compact_counters"] = ccnt
            g["compact_coords_sha"] = digest(cc)
            rng = np.random.default_rng(8)
            net = rng.standard_normal((S - 777, 4)).astype(np.float32)
            bg = rng.random((300, 3), dtype=np.float32)
            lg = rng.standard_normal((300, 3)).astype(np.float32)
            for name, dt in (("f32", np.float32), ("f16", np.float16)):
                rgb, dnet, rgbi, alpha = ol.ref_composite(net.astype(dt), cc, numsteps, ns_c, bg, lg, mean=0.001)
                g[f"comp_{name}_rgb"] = rgb
                g[f"comp_{name}_dnet_sha"] = digest(dnet)
                g[f"comp_{name}_dnet_head"] = dnet[:64]
                g[f"comp_{name}_rgbi"] = rgbi
                g[f"comp_{name}_alpha"] = alpha
    # occupancy-grid maintenance
    si = ol.pcg32_seed()
    g["pcg32_seed1337"] = si
    rng = np.random.default_rng(9)
    n_el = ol.G3 * 5
    g_in = np.where(rng.random(n_el) < 0.3, rng.random(n_el) * 0.05, -1.0).astype(np.float32)
    n = 20000
    pb = np.empty((n, 3), np.float32)
    ib = np.empty(n, np.uint32)
    r.ref_generate_grid_samples_constdt(ol._u32(n), ol._u64(int(si[0])), ol._u64(int(si[1])), ol._u32(3), ol._f32(-1.5), ol._f32(2.5), ol._ptr(g_in),
                                        ol._ptr(pb), ol._ptr(ib), ol._u32(3), ol._f32(0.01))
    g["gridgen_idx_sha"] = digest(ib)
    g["gridgen_pos_sha"] = digest(pb)
    g["gridgen_idx_head"] = ib[:64]
    mlp = rng.standard_normal(n).astype(np.float32)
    tb = np.zeros(n_el, np.float32)
    r.ref_splat_f32_constdt(ol._u32(n), ol._ptr(ib), ol._ptr(mlp), ol._ptr(tb))
    gb = g_in.copy()
    r.ref_ema_constdt(ol._u32(n_el), ol._f32(0.95), ol._ptr(gb), ol._ptr(tb))
    g["grid_after_ema_sha"] = digest(gb)
    mean = ol.grid_mean(gb)
    g["grid_mean"] = np.float32(mean)
    bb = np.zeros(ol.G3 * 5 // 8, np.uint8)
    r.ref_update_bitfield_constdt(ol._ptr(gb), ol._ptr(np.array([mean], np.float32)), ol._ptr(bb))
    g["grid_bitfield_sha"] = digest(bb)
    out = os.path.join(os.path.dirname(__file__), "ngp_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
