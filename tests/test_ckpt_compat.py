"""N4: the reference's params.pkl wire format (runner/runner.py:123-151) -- container, structure, and the mapping
to and from this repo's checkpoint dict.  CPU only: the mapping works on numpy arrays."""
import hashlib
import pickle

import numpy as np
import pytest
import torch

from jnerf_b200.utils import ckpt_compat as cc

N_TABLE = 245640 * 2          # config #1 table (T = 2^14) keeps the test small; the mapping is size-agnostic


def reference_style_ckpt(seed=0, extra_vars=True, linear=False):
    """A dict with the structure the reference's save_ckpt produces (as indexed by its load_ckpt, runner.py:133-151)."""
    rng = np.random.default_rng(seed)
    h = lambda n, s=1.0: (rng.standard_normal(n) * s).astype(np.float16)          # noqa: E731
    model = {"pos_encoder.m_grid": h(N_TABLE, 1e-2)}
    if linear:
        f = lambda *s: rng.standard_normal(s).astype(np.float32)                  # noqa: E731
        model.update({"density_mlp.0.weight": f(64, 32), "density_mlp.2.weight": f(16, 64), "rgb_mlp.0.weight": f(64, 32),
                      "rgb_mlp.2.weight": f(64, 64), "rgb_mlp.4.weight": f(3, 64)})
        sizes = [N_TABLE, 2048, 1024, 2048, 4096, 192]
    else:
        model.update({"density_mlp.con_weights": h(3072), "rgb_mlp.con_weights": h(7168)})
        sizes = [N_TABLE, 3072, 7168]
    if extra_vars:                                                                 # Vars of jt.Function members may be swept in too
        model["pos_encoder.encoder.m_hashmap_offsets_table"] = np.arange(17, dtype=np.int32)
    G = 128 ** 3
    mean = np.zeros(16384, np.float32)
    mean[0] = 0.0123
    sampler = {"density_grid": rng.random(5 * G, dtype=np.float32), "density_grid_tmp": np.zeros(8, np.float32),
               "density_grid_bitfield": rng.integers(0, 256, 5 * G // 8, dtype=np.uint8), "density_grid_mean": mean,
               "density_grid_ema_step": np.array([77], np.int32), "measured_batch_size": np.array([0], np.int32)}
    return {
        "global_step": 1234, "model": model, "sampler": sampler,
        "optimizer": {"defaults": {"base_lr": 0.1, "decay_start": 20000, "decay_interval": 10000, "decay_base": 0.33, "decay_end": 10000000,
                                   "steps": 1234, "m_learning_rate_factor": 1}},
        "nested_optimizer": {"defaults": {"lr": 0.1, "eps": 1e-15, "betas": (0.9, 0.99), "weight_decay": 0,
                                          "param_groups": [{"values": [np.abs(h(n, 1e-3)) for n in sizes], "m": [h(n, 1e-2) for n in sizes]}]}},
        "ema_optimizer": {"defaults": {"lr": 0, "decay": 0.95, "steps": 1234, "param_groups": [{"values": [model[k] if not linear else h(n) for k, n in
                                                                                                            zip(list(model)[:len(sizes)], sizes)]}]}},
    }


def test_container_tagged_and_bare(tmp_path):
    ref = reference_style_ckpt()
    p = str(tmp_path / "params.pkl")
    cc.write_reference_ckpt(ref, p)
    raw = open(p, "rb").read()
    assert raw.endswith(b"HCAJSLHD") and hashlib.sha1(raw[:-28]).digest() == raw[-28:-8]
    back = cc.read_reference_ckpt(p)
    assert back["global_step"] == 1234 and np.array_equal(back["model"]["pos_encoder.m_grid"], ref["model"]["pos_encoder.m_grid"])
    bare = str(tmp_path / "bare.pkl")
    open(bare, "wb").write(pickle.dumps(ref, 4))                                   # older writers: no trailer
    assert cc.read_reference_ckpt(bare)["optimizer"]["defaults"]["steps"] == 1234
    bad = str(tmp_path / "bad.pkl")
    open(bad, "wb").write(raw[:1000] + bytes([raw[1000] ^ 1]) + raw[1001:])
    with pytest.raises(ValueError, match="SHA-1"):
        cc.read_reference_ckpt(bad)


def test_unpickler_refuses_code(tmp_path):
    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("true",))
    p = str(tmp_path / "evil.pkl")
    open(p, "wb").write(pickle.dumps({"model": Evil()}, 4))
    with pytest.raises(pickle.UnpicklingError, match="numpy arrays"):
        cc.read_reference_ckpt(p)


def test_reference_to_native_and_back():
    ref = reference_style_ckpt(seed=3)
    nat = cc.reference_to_native(ref, N_TABLE)
    assert nat["global_step"] == 1234 and nat["optimizer"] == {"steps": 1234, "m_learning_rate_factor": 1.0}
    assert list(nat["model"]) == ["pos_encoder.m_grid", "density_mlp.con_weights", "rgb_mlp.con_weights"]
    assert [a.size for a in nat["nested_optimizer"]["m"]] == [N_TABLE, 3072, 7168]
    pg = ref["nested_optimizer"]["defaults"]["param_groups"][0]
    for k in range(3):
        assert nat["nested_optimizer"]["v"][k].dtype == np.float32
        v16, m16 = pg["values"][k].astype(np.float32), pg["m"][k].astype(np.float32)
        flushed = (v16 == 0) & (m16 != 0)                          # fp16 second moment flushed to 0 under a live first moment: floored at m^2
        assert np.array_equal(nat["nested_optimizer"]["v"][k], np.where(flushed, m16 * m16, v16))         # jt.nn.Adam: values = 2nd moment
        assert np.array_equal(nat["nested_optimizer"]["m"][k], m16)
    assert np.array_equal(nat["nested_optimizer"]["master"][0], ref["model"]["pos_encoder.m_grid"].astype(np.float32))
    assert nat["nested_optimizer"]["n_step"] == nat["ema_optimizer"]["steps"] == 1234
    assert nat["sampler"]["density_grid_mean"].shape == (1,) and abs(nat["sampler"]["density_grid_mean"][0] - 0.0123) < 1e-7
    assert "n_rays_per_batch" not in nat["sampler"]                                                       # not a Var in the reference
    # and back: every field the reference's load_ckpt indexes (runner.py:133-151) is present with the reference's values
    nat["sampler"]["n_rays_per_batch"] = 2816
    nat["sampler"]["rng"] = np.array([123456789, 3], np.int64)
    ref2 = cc.native_to_reference(nat)
    assert ref2["global_step"] == 1234
    for k in ("pos_encoder.m_grid", "density_mlp.con_weights", "rgb_mlp.con_weights"):
        assert np.array_equal(ref2["model"][k], ref["model"][k]) and ref2["model"][k].dtype == np.float16
    nested = ref2["nested_optimizer"]["defaults"]["param_groups"][0]
    ema = ref2["ema_optimizer"]["defaults"]
    for i in range(3):
        assert np.array_equal(nested["m"][i], pg["m"][i])
        keep = ~((pg["values"][i] == 0) & (pg["m"][i] != 0))
        assert np.array_equal(nested["values"][i][keep], pg["values"][i][keep])
        assert np.array_equal(ema["param_groups"][0]["values"][i], ref["ema_optimizer"]["defaults"]["param_groups"][0]["values"][i])
    assert ema["steps"] == 1234 and ref2["optimizer"]["defaults"]["steps"] == 1234
    assert ref2["sampler"]["density_grid_mean"].shape == (16384,) and np.array_equal(ref2["sampler"]["density_grid"], ref["sampler"]["density_grid"])
    assert np.array_equal(ref2["sampler"]["density_grid_bitfield"], ref["sampler"]["density_grid_bitfield"])
    nat2 = cc.reference_to_native(ref2, N_TABLE)
    assert nat2["sampler"]["n_rays_per_batch"] == 2816 and np.array_equal(nat2["sampler"]["rng"], [123456789, 3])


def test_linear_checkpoint_folds_into_flat_layout():
    """fp32 nn.Linear fallback (ngp_network.py:59-67) -> con_weights layout of fully_fused_mlp.py:26-40: (out,in) row-major per
    layer, last layer padded to 16 rows."""
    ref = reference_style_ckpt(seed=5, linear=True)
    nat = cc.reference_to_native(ref, N_TABLE)
    wd, wr = nat["model"]["density_mlp.con_weights"], nat["model"]["rgb_mlp.con_weights"]
    assert wd.size == 3072 and wr.size == 7168
    m = ref["model"]
    assert np.array_equal(wd[:2048].reshape(64, 32), m["density_mlp.0.weight"]) and np.array_equal(wd[2048:].reshape(16, 64), m["density_mlp.2.weight"])
    last = wr[2048 + 4096:].reshape(16, 64)
    assert np.array_equal(last[:3], m["rgb_mlp.4.weight"]) and not last[3:].any()
    x = np.random.default_rng(0).standard_normal((5, 32)).astype(np.float32)
    h = np.maximum(x @ m["rgb_mlp.0.weight"].T, 0)
    h = np.maximum(h @ m["rgb_mlp.2.weight"].T, 0)
    y = h @ m["rgb_mlp.4.weight"].T
    W0, W1, W2 = cc._flat_to_linear(wr, 1, 3)
    y2 = np.maximum(np.maximum(x @ W0.T, 0) @ W1.T, 0) @ W2.T
    assert np.allclose(y, y2)
    assert not nat["nested_optimizer"]["m"][0].any()                   # per-matrix moments cannot be mapped: restart


def test_file_to_torch_dict(tmp_path):
    """The exact call Runner.load_ckpt makes for a .pkl path, on the CPU device."""
    ref = reference_style_ckpt(seed=7)
    p = str(tmp_path / "params.pkl")
    cc.write_reference_ckpt(ref, p)
    dt = {"pos_encoder.m_grid": torch.float16, "density_mlp.con_weights": torch.float16, "rgb_mlp.con_weights": torch.float16}
    ck = cc.load_native_from_reference_file(p, N_TABLE, "cpu", dt)
    assert ck["model"]["pos_encoder.m_grid"].dtype == torch.float16 and ck["nested_optimizer"]["m"][0].dtype == torch.float32
    assert isinstance(ck["nested_optimizer"]["master"], list) and ck["sampler"]["density_grid_bitfield"].dtype == torch.uint8
    assert ck["global_step"] == 1234 and ck["ema_optimizer"]["steps"] == 1234
    # what Runner.save_ckpt hands over: torch tensors
    back = cc.native_to_reference(ck)
    assert np.array_equal(back["model"]["rgb_mlp.con_weights"], ref["model"]["rgb_mlp.con_weights"])
