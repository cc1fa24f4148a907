"""Name -> class registries and build_from_cfg, the plug-in mechanism of JNeRF (utils/registry.py:1-55).
Same registry names and semantics, so `projects/ngp/configs/*.py` resolve to this package's classes."""


class Registry:
    def __init__(self):
        self._modules = {}

    def register_module(self, name=None, module=None):
        def _register(mod):
            key = name if name is not None else mod.__name__
            assert key not in self._modules, f"{key} is already registered."
            self._modules[key] = mod
            return mod

        if module is not None:
            return _register(module)
        return _register

    def get(self, name):
        assert name in self._modules, f"{name} is not registered."
        return self._modules[name]


def build_from_cfg(cfg, registry, **kwargs):
    if isinstance(cfg, str):
        return registry.get(cfg)(**kwargs)
    if isinstance(cfg, dict):
        args = dict(cfg)
        args.update(kwargs)
        obj_cls = registry.get(args.pop("type"))
        try:
            return obj_cls(**args)
        except TypeError as e:
            raise TypeError(e if "<class" in str(e) else f"{obj_cls}.{e}")
    if isinstance(cfg, list):
        return [build_from_cfg(c, registry, **kwargs) for c in cfg]
    if cfg is None:
        return None
    raise TypeError(f"type {type(cfg)} not support")


DATASETS = Registry()
ENCODERS = Registry()
NETWORKS = Registry()
SAMPLERS = Registry()
LOSSES = Registry()
OPTIMS = Registry()
SCHEDULERS = Registry()
