from .._restated import AttnProcessor2_0  # noqa: F401


class FusedAttnProcessor2_0: pass
class LoRAAttnProcessor2_0: pass
class LoRAXFormersAttnProcessor: pass
class XFormersAttnProcessor: pass
