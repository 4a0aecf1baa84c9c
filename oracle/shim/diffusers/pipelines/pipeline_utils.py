import torch

from ..configuration_utils import ConfigMixin


class _Bar:
    def __init__(self, total=None): self.n = 0
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def update(self, n=1): self.n += n


class DiffusionPipeline(ConfigMixin):
    def __init__(self):
        pass

    def register_modules(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)
        self._module_names = list(kw)

    def progress_bar(self, iterable=None, total=None):
        return _Bar(total)

    def set_progress_bar_config(self, **kw):
        pass

    def maybe_free_model_hooks(self):
        pass

    def maybe_convert_prompt(self, prompt, tokenizer):
        return prompt

    def to(self, *a, **kw):
        for n in getattr(self, "_module_names", []):
            m = getattr(self, n, None)
            if isinstance(m, torch.nn.Module):
                m.to(*a, **kw)
        return self


class StableDiffusionMixin: pass
