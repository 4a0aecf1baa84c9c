"""faithdiff_dev_b200 — B200-native (sm_100a) implementation of FaithDiff's tiled latent-denoising hot path behind
the reference's Python call boundary.  Importing this package loads libfaithdiff_b200.so; there is no CPU or
PyTorch fallback for the kernels."""
from . import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)
from .pipeline import FaithDiffStableDiffusionXLPipeline, FaithDiffStableDiffusionXLPipelineOutput  # noqa: F401
from .scheduler import DDPMScheduler, EulerDiscreteScheduler  # noqa: F401
from .tiling import LocalAttention  # noqa: F401
from .unet import UNet2DConditionModel  # noqa: F401
from .vae import AutoencoderKL, Encoder  # noqa: F401
from .text_encoder import CLIPTextEncoder, CLIP_L_CONFIG, OPENCLIP_BIGG_CONFIG  # noqa: F401
from .color_fix import adain_color_fix, wavelet_color_fix  # noqa: F401
from .create_model import FaithDiff_pipeline, quantize_8bit_state_dict  # noqa: F401
from .image_process import check_image_size, resize_lanczos  # noqa: F401
from .bsrnet import RRDBNet, create_bsrnet  # noqa: F401
from .png import encode_png, save_png  # noqa: F401
