import functools
import inspect

from ._restated import AttrDict


class ConfigMixin:
    def register_to_config(self, **kwargs):
        kwargs.pop("kwargs", None)
        if not hasattr(self, "_internal_dict"):
            object.__setattr__(self, "_internal_dict", AttrDict(kwargs))
        else:
            self._internal_dict.update(kwargs)

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        params = {n: p.default for i, (n, p) in enumerate(inspect.signature(init).parameters.items()) if i > 0}
        new = {}
        for a, n in zip(args, params.keys()):
            new[n] = a
        new.update({k: kwargs.get(k, d) for k, d in params.items() if k not in new})
        object.__setattr__(self, "_internal_dict", AttrDict(new))
        init(self, *args, **kwargs)
    return inner_init
