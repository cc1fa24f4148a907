"""Pins the C restatement (oracle/ngp_oracle.c) against the reference's own kernel sources executed on the host
(oracle/_ref, built by `make -C oracle ref` from /root/reference through ref_shim/shim.h).  Runs only where the
_ref build exists (this container); the committed goldens (tests/golden) carry the same pins to the GPU box.
Integer / index outputs must be bit-exact; float outputs are bit-exact too in host-shim arithmetic (fma_mode 0)
except where libm expf replaces the device intrinsic identically on both sides."""
import numpy as np
import pytest
import oracle_lib as ol

pytestmark = pytest.mark.skipif(ol.ref("ref_hash_cpu") is None or ol.ref("ref_sampler_cpu_constdt") is None,
                                reason="oracle/_ref not built (reference tree absent)")


@pytest.fixture(autouse=True)
def host_arith():
    ol.oracle().orc_set_fma_mode(0)
    yield
    ol.oracle().orc_set_fma_mode(1)


def test_pcg32_kat():
    r = ol.ref("ref_sampler_cpu_constdt")
    si = ol.pcg32_seed(1337)
    si_ref = np.zeros(2, np.uint64)
    r.ref_pcg32_seed_constdt(ol._u64(1337), ol._ptr(si_ref))
    assert (si == si_ref).all()
    assert int(si[0]) == 0x4cfa1d1cde85af8f and int(si[1]) == 3          # SURVEY.md section 8c KAT
    f1 = ol.oracle().orc_pcg32_next_float(ol._ptr(si))
    f2 = ol.oracle().orc_pcg32_next_float(ol._ptr(si))
    # values as produced by the reference pcg32 itself (SURVEY.md 8c lists the same two numbers in swapped order)
    assert abs(f1 - 0.147699356) < 1e-8 and abs(f2 - 0.471029401) < 1e-8
    a, b = ol.pcg32_seed(1337), ol.pcg32_seed(1337)
    ol.pcg32_advance(a, 1 << 32)
    r.ref_pcg32_advance_constdt(ol._ptr(b), ol._i64(1 << 32))
    assert (a == b).all()


@pytest.mark.parametrize("aabb,log2T", [(1, 14), (1, 19), (4, 19)])
def test_level_table(aabb, log2T):
    cfg = ol.HashCfg(aabb, log2_hashmap_size=log2T)
    expect = {(1, 14): 245640, (1, 19): 6098120, (4, 19): 6537456}[(aabb, log2T)]
    assert cfg.n_entries == expect
    assert cfg.offsets[0] == 0 and cfg.offsets[1] == 4096


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("log2T", [14, 19])
def test_hash_fwd_bwd(dtype, log2T):
    cfg = ol.HashCfg(1, log2_hashmap_size=log2T)
    rng = np.random.default_rng(0)
    n = 512
    x = rng.random((n, 3), dtype=np.float32)
    x[0] = 0.0
    x[1] = 1.0                                     # the corner cases that hit index == resolution
    grid = np.random.default_rng(1).uniform(-1e-4, 1e-4, cfg.n_params).astype(dtype)
    out_ref, pos_soa = ol.ref_hash_fwd(cfg, x, grid)
    out = ol.hash_fwd(cfg, x, grid)
    assert np.array_equal(out.view(np.uint16 if dtype == np.float16 else np.uint32),
                          out_ref.view(np.uint16 if dtype == np.float16 else np.uint32))
    dy = (rng.standard_normal((n, 32)) * 1e-2).astype(dtype)
    g_ref = ol.ref_hash_bwd(cfg, pos_soa, dy)
    g = ol.hash_bwd(cfg, x, dy)
    assert np.array_equal(g, g_ref)


def test_hash_kat():
    # hash(1,2,3)=212041242 -> %2^19=228890 ; %2^14=15898
    h = (1 ^ (2 * 19349663) ^ (3 * 83492791)) & 0xFFFFFFFF
    assert h == 212041242 and h % (1 << 19) == 228890 and h % (1 << 14) == 15898
    assert ol.oracle().orc_morton3D(1, 2, 3) == 53 and ol.oracle().orc_morton3D(127, 127, 127) == 2097151


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_sh(dtype):
    r = ol.ref("ref_sampler_cpu_constdt")
    d = np.random.default_rng(3).random((257, 3), dtype=np.float32)
    out_ref = np.empty((257, 16), dtype)
    (r.ref_sh_f32_constdt if dtype == np.float32 else r.ref_sh_f16_constdt)(ol._u32(257), ol._ptr(d), ol._ptr(out_ref))
    out = ol.sh(d, dtype)
    assert np.array_equal(out, out_ref)


@pytest.mark.parametrize("const_dt", [True, False])
def test_march_compact(const_dt):
    bits, _ = ol.sphere_bitfield(0.3)
    o, d = ol.random_rays(300, seed=5)
    aabb = (0.0, 1.0) if const_dt else (-1.5, 2.5)
    a = ol.march(o, d, bits, aabb=aabb, const_dt=const_dt, max_samples=300 * 1024)
    b = ol.ref_march(o, d, bits, aabb=aabb, const_dt=const_dt, max_samples=300 * 1024)
    assert (a[3] == b[3]).all() and a[3][1] > 1000
    assert np.array_equal(a[2], b[2])                                  # numsteps + base, ray order
    S = int(a[3][1])
    assert np.array_equal(a[0][:S].view(np.uint32), b[0][:S].view(np.uint32))   # sample coords bit-exact
    assert np.array_equal(a[1], b[1])
    if const_dt:
        cap = S - 777                                                  # force truncation
        ca = ol.compact(a[0], a[2], cap)
        cb = ol.ref_compact(b[0], b[2], cap)
        assert np.array_equal(ca[0], cb[0]) and np.array_equal(ca[1], cb[1]) and (ca[2] == cb[2]).all()


def test_march_overflow():
    bits, _ = ol.sphere_bitfield(0.3)
    o, d = ol.random_rays(64, seed=6)
    a = ol.march(o, d, bits, max_samples=2000)
    b = ol.ref_march(o, d, bits, max_samples=2000)
    assert np.array_equal(a[2], b[2]) and (a[3] == b[3]).all()
    assert (a[2][:, 0] == 0).any()


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_composite(dtype):
    bits, _ = ol.sphere_bitfield(0.3)
    o, d = ol.random_rays(200, seed=7)
    coords, _, numsteps, cnt = ol.march(o, d, bits, max_samples=200 * 1024)
    S = int(cnt[1])
    cc, ns_c, _ = ol.compact(coords, numsteps, S - 100)
    rng = np.random.default_rng(8)
    net = rng.standard_normal((S - 100, 4)).astype(dtype)
    bg = rng.random((200, 3), dtype=np.float32)
    lg = (rng.standard_normal((200, 3))).astype(np.float32)
    rgb_ref, dnet_ref, rgbi_ref, alpha_ref = ol.ref_composite(net, cc, numsteps, ns_c, bg, lg, mean=0.001)
    rgb = ol.composite_fwd(net, cc, numsteps, ns_c, bg)
    np.testing.assert_array_equal(rgb, rgb_ref)
    dnet = ol.composite_bwd(net, cc, ns_c, lg, rgb, 0.001)
    np.testing.assert_array_equal(dnet, dnet_ref)
    rgbi, alpha = ol.composite_infer(net, cc, ns_c)
    np.testing.assert_array_equal(rgbi, rgbi_ref)
    np.testing.assert_array_equal(alpha, alpha_ref)


def test_grid_update():
    r = ol.ref("ref_sampler_cpu_constdt")
    n_img = 7
    rng = np.random.default_rng(9)
    # cameras on a sphere looking at the centre, column-major 3x4
    xf = np.zeros((n_img, 12), np.float32)
    for j in range(n_img):
        v = rng.normal(size=3); v /= np.linalg.norm(v)
        pos = 0.5 + 1.2 * v
        zc = -v
        up = np.array([0, 0, 1.0]); xc = np.cross(up, zc); xc /= np.linalg.norm(xc); yc = np.cross(zc, xc)
        xf[j] = np.concatenate([xc, yc, zc, pos]).astype(np.float32)
    focal = np.full((n_img, 2), 1100.0, np.float32)
    n_el = ol.G3 * 5
    g_a = np.zeros(n_el, np.float32); g_b = np.zeros(n_el, np.float32)
    ol.mark_untrained(g_a, focal, xf, (800, 800))
    r.ref_mark_untrained_constdt(ol._u32(n_el), ol._ptr(g_b), ol._u32(n_img), ol._ptr(focal), ol._ptr(xf), ol._i32(800), ol._i32(800))
    assert np.array_equal(g_a, g_b) and (g_a < 0).any() and (g_a == 0).any()
    si = ol.pcg32_seed()
    for thresh, step, casc in [(-0.01, 0, 1), (0.01, 3, 3)]:
        n = 20000
        g_in = np.where(rng.random(n_el) < 0.3, rng.random(n_el) * 0.05, -1.0).astype(np.float32)
        pa, ia = ol.generate_grid_samples(n, si, step, (-1.5, 2.5), g_in, casc, thresh)
        pb = np.empty((n, 3), np.float32); ib = np.empty(n, np.uint32)
        r.ref_generate_grid_samples_constdt(ol._u32(n), ol._u64(int(si[0])), ol._u64(int(si[1])), ol._u32(step), ol._f32(-1.5), ol._f32(2.5),
                                            ol._ptr(g_in), ol._ptr(pb), ol._ptr(ib), ol._u32(casc), ol._f32(thresh))
        assert np.array_equal(ia, ib) and np.array_equal(pa, pb)
    for dt in (np.float32, np.float16):
        mlp = rng.standard_normal(n).astype(dt)
        ta = np.zeros(n_el, np.float32); tb = np.zeros(n_el, np.float32)
        ol.splat(ia, mlp, ta)
        (r.ref_splat_f32_constdt if dt == np.float32 else r.ref_splat_f16_constdt)(ol._u32(n), ol._ptr(ia), ol._ptr(mlp), ol._ptr(tb))
        assert np.array_equal(ta, tb)
    ga, gb = g_in.copy(), g_in.copy()
    ol.ema(ga, ta)
    r.ref_ema_constdt(ol._u32(n_el), ol._f32(0.95), ol._ptr(gb), ol._ptr(tb))
    assert np.array_equal(ga, gb)
    mean = ol.grid_mean(ga)
    ba = ol.update_bitfield(ga, mean)
    bb = np.zeros_like(ba)
    r.ref_update_bitfield_constdt(ol._ptr(gb), ol._ptr(np.array([mean], np.float32)), ol._ptr(bb))
    assert np.array_equal(ba, bb) and ba.any()
