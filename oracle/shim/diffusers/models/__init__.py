from .._restated import AutoencoderKL, UNet2DConditionModelBase
from ..configuration_utils import ConfigMixin


class UNet2DConditionModel(UNet2DConditionModelBase):
    pass
