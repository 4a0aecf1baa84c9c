class FromSingleFileMixin: pass
class IPAdapterMixin: pass
class StableDiffusionXLLoraLoaderMixin: pass
class TextualInversionLoaderMixin: pass
class PeftAdapterMixin: pass
class UNet2DConditionLoadersMixin: pass
