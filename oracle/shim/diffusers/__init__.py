"""ORACLE / TEST INFRASTRUCTURE ONLY. Minimal stand-in for `diffusers==0.28.0` (absent from this image and
from /root/reference) exposing exactly the names the reference imports; the arithmetic
lives in `_restated.py`. Put `oracle/shim` on sys.path to let the reference's own files import unmodified."""
from ._restated import AutoencoderKL, DDPMScheduler, EulerDiscreteScheduler
from .models import UNet2DConditionModel

__version__ = "0.28.0+oracle-shim"
