"""Config-as-python-module loader mirroring JNeRF's utils/config.py:16-163: attribute access returning None on a
miss (:24-27), `_base_` inheritance with `_cover_` override (:61-101), and the process-wide singleton that the
plugin classes use as a service locator (cfg.dataset_obj, cfg.model_obj, cfg.m_training_step, ...)."""
import copy
import importlib.util
import os


class Config(dict):
    def __getattr__(self, name):
        return self.get(name, None)          # missing keys read as None (utils/config.py:24-27)

    def __setattr__(self, name, value):
        self[name] = value

    def load_from_file(self, path):
        self.clear()
        self.update(_wrap(_load(path)))

    def dump(self):
        return {k: v for k, v in self.items() if not k.endswith("_obj")}


def _wrap(d):
    if isinstance(d, dict):
        c = Config()
        for k, v in d.items():
            c[k] = _wrap(v)
        return c
    if isinstance(d, (list, tuple)):
        return type(d)(_wrap(v) for v in d)
    return d


def _merge(base, new):
    out = copy.deepcopy(base)
    for k, v in new.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_cover_", False):
            out[k] = _merge(out[k], v)
        else:
            v = copy.deepcopy(v)
            if isinstance(v, dict):
                v.pop("_cover_", None)
            out[k] = v
    return out


def _load(path):
    path = os.path.abspath(path)
    assert os.path.exists(path), f"{path} not exists"
    if path.endswith((".yaml", ".yml")):
        import yaml
        with open(path) as f:
            d = yaml.safe_load(f)
    else:
        spec = importlib.util.spec_from_file_location("_ngp_cfg", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        d = {k: v for k, v in vars(mod).items() if not k.startswith("__") and not callable(v) and not isinstance(v, type(os))}
    bases = d.pop("_base_", [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, _load(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, d)


_cfg = Config()


def init_cfg(filename):
    _cfg.load_from_file(filename)
    return _cfg


def get_cfg():
    return _cfg


def update_cfg(**kwargs):
    _cfg.update(_wrap(kwargs))
    return _cfg
