"""Checkpoint wire format of the reference: `params.pkl` written by Runner.save_ckpt
(runner/runner.py:123-131) with jt.save and read back by Runner.load_ckpt (:133-151).

What the file is.  jt.save converts every jt.Var to a numpy array and pickles the nested dict; Jittor's pickle writer is outside
the reference tree (third-party, version unpinned -- SURVEY 8c), so the container is restated from its published behaviour:
pickle protocol 4, followed by the 20-byte SHA-1 of the pickle bytes and the 8-byte tag b"HCAJSLHD"; older Jittor versions
wrote the bare pickle.  `read_reference_ckpt` accepts both (and verifies the digest when the tag is present);
`write_reference_ckpt` writes the tagged form.  The STRUCTURE of the dict, on the other hand, is pinned by the reference's own
load_ckpt, which indexes it explicitly:

    {'global_step': int,
     'model':            {'pos_encoder.m_grid': (n_params,), 'density_mlp.con_weights': (3072,), 'rgb_mlp.con_weights': (7168,), ...},
     'sampler':          {'density_grid': f32[C*128^3], 'density_grid_bitfield': u8[C*128^3/8], 'density_grid_mean': f32[16384],
                          'density_grid_ema_step': i32[1], ...},
     'optimizer':        {'defaults': {'steps', 'm_learning_rate_factor', 'base_lr', 'decay_*'}},            # ExpDecay, expdecay.py:8-19
     'nested_optimizer': {'defaults': {'lr', 'eps', 'betas', 'param_groups': [{'values': [v_i], 'm': [m_i], ...}]}},   # jt.nn.Adam: values = 2nd moment
     'ema_optimizer':    {'defaults': {'decay', 'steps', 'param_groups': [{'values': [ema_i]}]}}}             # ema.py:9-24

Parameters are matched by element count, not by key or position: whether Jittor also sweeps the Vars held by jt.Function members
(GridEncode's scratch buffers) into state_dict()/parameters() cannot be checked without Jittor, and the three
trainable tensors of NGP have distinct sizes.  `nn.Linear` checkpoints (the reference's fp32 fallback, ngp_network.py:59-67) are
folded into the flat `con_weights` layout of OPS/fully_fused_mlp.py:26-40.

Two reference behaviours that are NOT reproduced (documented deviations):
  * the reference excludes Adam's `n_step` from its state (optims/adam.py:13-16), so its bias correction restarts at 1 after a
    resume; here n_step is restored from the EMA step counter (the two advance together, runner.py:75-76);
  * `n_rays_per_batch` and the pcg32 state are not Vars and are therefore not in the reference file; they are carried under
    extra keys ('sampler' -> 'n_rays_per_batch', 'rng') which the reference's load_state_dict ignores.
  * Adam moments and EMA values are written in the PARAMETER dtype (fp16 for the NGP configs with fp16=True), as jt.nn.Adam / EMA
    allocate them (`jt.zeros(p.shape, p.dtype)`, `p.copy()`): a .pkl round trip rounds this repo's fp32 optimizer state to fp16.
    Use the native `.pt` format for a lossless resume; `.pkl` is the interchange format with the reference.
All functions work on numpy arrays / python scalars; nothing here touches the GPU."""
import hashlib
import io
import pickle

import numpy as np

_TAG = b"HCAJSLHD"
N_DENSITY, N_RGB = 3072, 7168                 # 32*64 + 64*16 ; 32*64 + 64*64 + 64*16  (ngp_network.py:52-53)


class _NumpyOnlyUnpickler(pickle.Unpickler):
    """A checkpoint is data: only numpy array reconstruction and builtin containers are allowed to be unpickled."""
    _OK = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"), ("numpy", "dtype"),
           ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("collections", "OrderedDict"),
           ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer")}

    def find_class(self, module, name):
        if (module, name) in self._OK:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"params.pkl may only contain numpy arrays and builtin containers, found {module}.{name}")


def read_reference_ckpt(path):
    raw = open(path, "rb").read()
    if raw.endswith(_TAG):
        body, digest = raw[:-28], raw[-28:-8]
        if hashlib.sha1(body).digest() != digest:
            raise ValueError(f"{path}: SHA-1 trailer does not match the pickle body (truncated or corrupted file)")
        raw = body
    return _NumpyOnlyUnpickler(io.BytesIO(raw)).load()


def write_reference_ckpt(ckpt, path):
    body = pickle.dumps(_to_numpy(ckpt), 4)
    with open(path, "wb") as f:
        f.write(body + hashlib.sha1(body).digest() + _TAG)


def _to_numpy(x):
    if isinstance(x, dict):
        return {k: _to_numpy(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_numpy(v) for v in x)
    if hasattr(x, "detach") and hasattr(x, "cpu"):                       # torch tensor (kept optional: no torch import here)
        return x.detach().cpu().numpy()
    return x


# ---------------------------------------------------------------------------------------------------- parameters
def _linear_to_flat(ws):
    """[W (out,in) ...] of an nn.Sequential(Linear, ReLU, ...) -> flat (out,in)-row-major vector with the last layer padded
    to 16 rows -- FullyFusedMlp_weight.__init__ stores weights (in,out), pads columns to 16 and flattens the transpose
    (fully_fused_mlp.py:26-40); nn.Linear already keeps (out,in)."""
    ws = [np.asarray(w) for w in ws]
    last = ws[-1]
    if last.shape[0] < 16:
        last = np.concatenate([last, np.zeros((16 - last.shape[0], last.shape[1]), last.dtype)], 0)
    return np.concatenate([w.reshape(-1) for w in ws[:-1]] + [last.reshape(-1)])


def _flat_to_linear(flat, n_hidden_matmuls, n_out):
    shapes = [(64, 32)] + [(64, 64)] * n_hidden_matmuls + [(16, 64)]
    out, o = [], 0
    for s in shapes:
        out.append(np.asarray(flat[o:o + s[0] * s[1]]).reshape(s))
        o += s[0] * s[1]
    out[-1] = out[-1][:n_out]
    return out


def model_params_from_reference(model_sd, n_table):
    """-> {'m_grid', 'density', 'rgb'} numpy arrays (dtype as stored) from the reference's model.state_dict()."""
    by_size = {}
    for k, v in model_sd.items():
        v = np.asarray(v)
        by_size.setdefault(v.size, []).append((k, v))
    out = {}
    grid = [kv for kv in by_size.get(n_table, []) if "grid" in kv[0]] or by_size.get(n_table, [])
    if not grid:
        raise KeyError(f"no tensor with {n_table} elements (the hash table for this aabb_scale) in the checkpoint's model dict: "
                       f"{sorted((k, np.asarray(v).shape) for k, v in model_sd.items())[:12]}")
    out["m_grid"] = grid[0][1].reshape(-1)
    for name, n, nhm, n_out in (("density", N_DENSITY, 0, 16), ("rgb", N_RGB, 1, 3)):
        hit = [kv for kv in by_size.get(n, []) if "con_weights" in kv[0]]
        if hit:
            out[name] = hit[0][1].reshape(-1)
            continue
        prefix = "density_mlp." if name == "density" else "rgb_mlp."
        lin = sorted((k, np.asarray(v)) for k, v in model_sd.items() if k.startswith(prefix) and k.endswith(".weight"))
        if len(lin) != nhm + 2:
            raise KeyError(f"neither {prefix}con_weights ({n} elements) nor {nhm + 2} {prefix}<i>.weight matrices in the checkpoint")
        out[name] = _linear_to_flat([w for _, w in lin])
    return out


def _by_size(tensors, sizes):
    """Pick, for every wanted size, the tensor of a list with that element count."""
    got = {}
    for t in tensors:
        t = np.asarray(t)
        if t.size in sizes and t.size not in got:
            got[t.size] = t.reshape(-1)
    missing = [s for s in sizes if s not in got]
    if missing:
        raise KeyError(f"optimizer state without tensors of {missing} elements")
    return [got[s] for s in sizes]


# ---------------------------------------------------------------------------------------------------- whole checkpoints
def reference_to_native(ref, n_table):
    """Reference params.pkl dict -> the dict Runner.load_ckpt consumes (numpy arrays; the loader moves them to the device).
    Parameter order: [m_grid, density_mlp.con_weights, rgb_mlp.con_weights] (NGPNetworks.parameters())."""
    p = model_params_from_reference(ref["model"], n_table)
    sizes = [n_table, N_DENSITY, N_RGB]
    nested = ref["nested_optimizer"]["defaults"]
    pg = nested["param_groups"][0]
    ema_d = ref["ema_optimizer"]["defaults"]
    try:
        if "values_f32" in pg:
            # written by native_to_reference: the fp32 optimizer state next to the parameter-dtype copies the reference reads
            v = _by_size(pg["values_f32"], sizes)
            m = _by_size(pg["m_f32"], sizes)
            ema_vals = _by_size(ema_d["param_groups"][0].get("values_f32", ema_d["param_groups"][0]["values"]), sizes)
        else:
            v = _by_size(pg["values"], sizes)
            m = _by_size(pg["m"], sizes)
            ema_vals = _by_size(ema_d["param_groups"][0]["values"], sizes)
            # a reference-written file holds fp16 moments: second moments below fp16's smallest subnormal (6e-8) are stored as 0
            # while the first moment survives, and lr * m / (sqrt(0) + 1e-15) would throw the entry to inf on the next step.
            # Floor v at m^2 (|update| <= lr, Adam's own bound) for exactly those entries.
            fixed = []
            for mm, vv in zip(m, v):
                mm32, vv32 = np.asarray(mm, np.float32), np.asarray(vv, np.float32).copy()
                bad = (vv32 == 0) & (mm32 != 0)
                vv32[bad] = mm32[bad] * mm32[bad]
                still = bad & (vv32 == 0)                            # m^2 underflows fp32 as well: drop the momentum
                if still.any():
                    mm32 = mm32.copy()
                    mm32[still] = 0
                fixed.append((mm32, vv32))
            m, v = [f[0] for f in fixed], [f[1] for f in fixed]
    except KeyError:
        # nn.Linear checkpoints keep one moment tensor per matrix; the flat fused layout restarts the moments and takes the
        # parameters themselves as EMA values (after every ema_step the two are equal, ema.py:33-36)
        m = [np.zeros(n, np.float32) for n in sizes]
        v = [np.zeros(n, np.float32) for n in sizes]
        ema_vals = [p["m_grid"], p["density"], p["rgb"]]
    steps = int(ema_d["steps"])
    opt = ref["optimizer"]["defaults"]
    s = ref["sampler"]
    mean = np.asarray(s["density_grid_mean"], np.float32).reshape(-1)[:1]
    sampler = {"density_grid": np.asarray(s["density_grid"], np.float32).reshape(-1),
               "density_grid_bitfield": np.asarray(s["density_grid_bitfield"], np.uint8).reshape(-1),
               "density_grid_mean": mean, "density_grid_ema_step": np.asarray(s["density_grid_ema_step"], np.int32).reshape(-1)[:1]}
    if "n_rays_per_batch" in s:
        sampler["n_rays_per_batch"] = int(np.asarray(s["n_rays_per_batch"]).reshape(-1)[0])
    if "rng" in s:
        sampler["rng"] = np.asarray(s["rng"], np.int64).reshape(-1)
    f32 = lambda a: np.asarray(a, np.float32)                                            # noqa: E731
    return {
        "global_step": int(ref["global_step"]),
        "model": {"pos_encoder.m_grid": p["m_grid"], "density_mlp.con_weights": p["density"], "rgb_mlp.con_weights": p["rgb"]},
        "sampler": sampler,
        "optimizer": {"steps": int(opt["steps"]), "m_learning_rate_factor": float(opt["m_learning_rate_factor"])},
        "nested_optimizer": {"n_step": steps, "lr": float(nested.get("lr", 0.1)), "m": [f32(x) for x in m], "v": [f32(x) for x in v],
                             "master": [f32(x) for x in ema_vals]},
        "ema_optimizer": {"steps": steps, "decay": float(ema_d.get("decay", 0.95))},
    }


def native_to_reference(nat, adam_hyper=None, expdecay_hyper=None, param_dtype=np.float16):
    """The dict Runner.save_ckpt builds (tensors or arrays) -> the reference's params.pkl structure (numpy)."""
    nat = _to_numpy(nat)
    model = {k: np.asarray(v) for k, v in nat["model"].items()}
    order = ["pos_encoder.m_grid", "density_mlp.con_weights", "rgb_mlp.con_weights"]
    nested, ema = nat["nested_optimizer"], nat["ema_optimizer"]
    cast = lambda a: np.asarray(a).astype(param_dtype)                                   # noqa: E731  Jittor keeps m / values in the parameter dtype
    f32 = lambda a: np.asarray(a, np.float32)                                            # noqa: E731
    hyper = dict(lr=float(nested["lr"]), eps=1e-15, betas=(0.9, 0.99), weight_decay=0)
    hyper.update(adam_hyper or {})
    dec = dict(base_lr=hyper["lr"], decay_start=20000, decay_interval=10000, decay_base=0.33, decay_end=10000000)
    dec.update(expdecay_hyper or {})
    s = nat["sampler"]
    mean = np.zeros(16384, np.float32)                                                   # density_grid_sampler.py:91-92: div_round_up(128^3, 128) slots
    mean[0] = float(np.asarray(s["density_grid_mean"]).reshape(-1)[0])
    sampler = {"density_grid": np.asarray(s["density_grid"], np.float32), "density_grid_bitfield": np.asarray(s["density_grid_bitfield"], np.uint8),
               "density_grid_mean": mean, "density_grid_ema_step": np.asarray(s["density_grid_ema_step"], np.int32).reshape(1)}
    for k in ("n_rays_per_batch", "rng"):
        if k in s:
            sampler[k] = np.asarray(s[k])
    return {
        "global_step": int(nat["global_step"]),
        "model": {k: model[k] for k in order},
        "sampler": sampler,
        "optimizer": {"defaults": dict(dec, steps=int(nat["optimizer"]["steps"]), m_learning_rate_factor=nat["optimizer"]["m_learning_rate_factor"])},
        # "values"/"m" in the parameter dtype are what the reference's load_ckpt reads (runner.py:133-151); the *_f32 keys are ignored
        # by it and preferred by reference_to_native, so a save -> load in this repo is lossless (fp16 flushes most second moments to 0)
        "nested_optimizer": {"defaults": dict(hyper, param_groups=[{"values": [cast(x) for x in nested["v"]], "m": [cast(x) for x in nested["m"]],
                                                                    "values_f32": [f32(x) for x in nested["v"]], "m_f32": [f32(x) for x in nested["m"]]}])},
        "ema_optimizer": {"defaults": {"lr": 0, "decay": float(ema["decay"]), "steps": int(ema["steps"]),
                                       "param_groups": [{"values": [cast(x) for x in nested["master"]],
                                                         "values_f32": [f32(x) for x in nested["master"]]}]}},
    }


def load_native_from_reference_file(path, n_table, device, model_dtypes=None):
    """params.pkl -> the dict Runner.load_ckpt consumes, with torch tensors on `device` (model tensors cast to model_dtypes[key])."""
    import torch
    nat = reference_to_native(read_reference_ckpt(path), n_table)

    def dev(x):
        if isinstance(x, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(x)).to(device)
        if isinstance(x, list):
            return [dev(t) for t in x]
        if isinstance(x, dict):
            return {k: dev(v) for k, v in x.items()}
        return x

    ck = dev(nat)
    if model_dtypes:
        ck["model"] = {k: v.to(model_dtypes[k]) for k, v in ck["model"].items()}
    return ck
