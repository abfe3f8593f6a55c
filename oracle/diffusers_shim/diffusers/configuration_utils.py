"""ConfigMixin / register_to_config restated from diffusers 0.24.0 (behaviour used by the reference only)."""
import functools
import inspect
import json


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        kwargs.pop("kwargs", None)
        if not hasattr(self, "_internal_dict"):
            self._internal_dict = FrozenDict(kwargs)
        else:
            d = dict(self._internal_dict)
            d.update(kwargs)
            self._internal_dict = FrozenDict(d)

    @property
    def config(self):
        return self._internal_dict

    def __getattr__(self, name):
        # diffusers 0.24 lets `model.in_channels` fall through to `model.config.in_channels` (with a deprecation
        # warning); the reference relies on it (pipeline_pose2vid_long.py:408 `self.denoising_unet.in_channels`).
        is_in_config = "_internal_dict" in self.__dict__ and name in self.__dict__["_internal_dict"]
        if is_in_config and name not in self.__dict__:
            return self.__dict__["_internal_dict"][name]
        sup = super()
        if hasattr(sup, "__getattr__"):
            return sup.__getattr__(name)
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    @classmethod
    def load_config(cls, path, **kwargs):
        with open(path) as f:
            return json.load(f)

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        init = {k: v for k, v in dict(config).items() if k in sig and not k.startswith("_")}
        init.update({k: v for k, v in kwargs.items() if k in sig})
        return cls(**init)


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        init_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("_")}
        init(self, *args, **init_kwargs)
        sig = inspect.signature(init)
        params = {n: p.default for i, (n, p) in enumerate(sig.parameters.items()) if i > 0}
        new_kwargs = {}
        for arg, name in zip(args, params.keys()):
            new_kwargs[name] = arg
        new_kwargs.update({k: init_kwargs.get(k, default) for k, default in params.items() if k not in new_kwargs})
        getattr(self, "register_to_config")(**new_kwargs)

    return inner_init
