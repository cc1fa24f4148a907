"""Edge cases of the sampler / renderer rows, oracle restatement vs the reference's own sources run on the host (oracle/_ref):
rays that miss the box, rays that start inside it, axis-parallel rays, an empty and a full occupancy grid (the 1024-step cap),
zero-sample rays in compaction and compositing, capacity 0.  Everything here is bit-exact."""
import numpy as np
import pytest

import oracle_lib as ol


@pytest.fixture(autouse=True)
def host_arithmetic():
    ol.oracle().orc_set_fma_mode(0)      # the reference sources run on the host here: no FMA contraction
    yield
    ol.oracle().orc_set_fma_mode(1)


def _same_march(o, d, bits, **kw):
    a = ol.march(o, d, bits, **kw)
    b = ol.ref_march(o, d, bits, **kw)
    assert (a[3] == b[3]).all()
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1])
    S = min(int(a[3][1]), kw.get("max_samples", 4096 * 1024))
    assert np.array_equal(a[0][:S].view(np.uint32), b[0][:S].view(np.uint32))
    return a


def _special_rays():
    o = np.array([[2.0, 2.0, 2.0],      # points away from the box: misses
                  [0.5, 0.5, 0.5],      # starts inside the box (and inside the sphere)
                  [0.5, 0.5, -1.0],     # axis-parallel, two zero direction components (division by zero in the slab test)
                  [-1.0, 0.5, 0.5],     # axis-parallel along x
                  [0.5, 0.5, 1.0],      # starts exactly on a face, pointing inwards
                  [0.2, 0.2, -0.5],     # grazes: inside the box, outside the sphere
                  [0.5, -1.0, 0.5]], np.float32)
    d = np.array([[1.0, 1.0, 1.0], [0.0, 0.6, 0.8], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 0.0, 1.0], [0.0, 1.0, 0.0]], np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


@pytest.mark.parametrize("const_dt", [True, False])
def test_special_rays(const_dt):
    bits, _ = ol.sphere_bitfield(0.3)
    o, d = _special_rays()
    aabb = (0.0, 1.0) if const_dt else (-1.5, 2.5)
    a = _same_march(o, d, bits, aabb=aabb, const_dt=const_dt, max_samples=len(o) * 1024)
    n = a[2][:, 0]
    assert n[0] == 0 and n[5] == 0 and n[1] > 0 and n[2] > 0 and n[3] > 0       # miss / graze give no samples, hits do


def test_empty_and_full_grid():
    o, d = ol.random_rays(48, seed=11)
    empty = np.zeros_like(ol.sphere_bitfield(0.3)[0])
    a = _same_march(o, d, empty, max_samples=48 * 1024)
    assert int(a[3][1]) == 0 and not a[2][:, 0].any()
    full = np.full_like(empty, 0xFF)
    a = _same_march(o, d, full, max_samples=48 * 1024)
    assert a[2][:, 0].max() <= 1024 and a[2][:, 0].max() > 500                    # NERF_STEPS cap (ray_sampler.h)
    # with every cell occupied the sample count follows from the box chord alone: never more than 1024, base = prefix sum
    assert np.array_equal(a[2][:, 1], np.concatenate([[0], np.cumsum(a[2][:-1, 0])]).astype(a[2].dtype))


def test_zero_sample_rays_through_compaction_and_compositing():
    bits, _ = ol.sphere_bitfield(0.3)
    o, d = ol.random_rays(120, seed=12)
    o2, d2 = _special_rays()
    o, d = np.concatenate([o2, o]), np.concatenate([d2, d])
    R = len(o)
    coords, _, numsteps, cnt = ol.march(o, d, bits, max_samples=R * 1024)
    S = int(cnt[1])
    assert (numsteps[:, 0] == 0).sum() >= 2
    for cap in (S, S // 2, 1, 0):                                                   # no truncation, truncation, degenerate capacities
        ca = ol.compact(coords, numsteps, cap)
        cb = ol.ref_compact(coords, numsteps, cap)
        assert np.array_equal(ca[0], cb[0]) and np.array_equal(ca[1], cb[1]) and (ca[2] == cb[2]).all()
    cc, ns_c, _ = ol.compact(coords, numsteps, S)
    rng = np.random.default_rng(13)
    net = rng.standard_normal((S, 4)).astype(np.float16)
    bg = rng.random((R, 3), dtype=np.float32)
    lg = rng.standard_normal((R, 3)).astype(np.float32)
    rgb_ref, dnet_ref, rgbi_ref, alpha_ref = ol.ref_composite(net, cc, numsteps, ns_c, bg, lg, mean=0.5)   # mean > 0.01: no L1 term
    rgb = ol.composite_fwd(net, cc, numsteps, ns_c, bg)
    np.testing.assert_array_equal(rgb, rgb_ref)
    empty = numsteps[:, 0] == 0
    np.testing.assert_array_equal(rgb[empty], bg[empty])                            # empty ray -> background (calc_rgb.h:35-39)
    np.testing.assert_array_equal(ol.composite_bwd(net, cc, ns_c, lg, rgb, 0.5), dnet_ref)
    rgbi, alpha = ol.composite_infer(net, cc, ns_c)
    np.testing.assert_array_equal(rgbi, rgbi_ref)
    np.testing.assert_array_equal(alpha, alpha_ref)
    assert not alpha[empty].any()


def test_hash_extreme_positions():
    """Positions on the faces / corners of [0,1]^3 and denormal-small ones: indices and interpolation stay bit-identical."""
    cfg = ol.HashCfg(1, log2_hashmap_size=14)
    rng = np.random.default_rng(14)
    x = rng.random((256, 3), dtype=np.float32)
    x[:8] = np.array([[0, 0, 0], [1, 1, 1], [0, 1, 0], [1, 0, 1], [0.5, 0, 1], [np.nextafter(np.float32(1), np.float32(0))] * 3,
                      [1e-38, 1e-38, 1e-38], [0.999999, 1e-7, 0.5]], np.float32)
    grid = rng.uniform(-1e-1, 1e-1, cfg.n_params).astype(np.float32)
    y = ol.hash_fwd(cfg, x, grid)
    y_ref, pos_soa = ol.ref_hash_fwd(cfg, x, grid)
    np.testing.assert_array_equal(y, y_ref)
    dy = rng.standard_normal((256, 32)).astype(np.float32)
    np.testing.assert_array_equal(ol.hash_bwd(cfg, x, dy), ol.ref_hash_bwd(cfg, pos_soa, dy))
