import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields, is_dataclass

USE_PEFT_BACKEND = False


class BaseOutput(OrderedDict):
    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return tuple(self.values())[k]


def deprecate(*a, **kw): return None
def is_invisible_watermark_available(): return False
def is_torch_xla_available(): return False
def is_flax_available(): return False
def is_torch_version(op, v): return True
def scale_lora_layers(*a, **kw): return None
def unscale_lora_layers(*a, **kw): return None


def replace_example_docstring(doc):
    def deco(fn): return fn
    return deco


class logging:  # noqa: N801
    @staticmethod
    def get_logger(name): return _pylogging.getLogger(name)
