"""GPU drop-ins for utils/color_fix.py:14-42 (`adain_color_fix`, `wavelet_color_fix`): PIL in → PIL out, same
signatures; fp32 math in image.cu.  `*_u8` variants take/return HWC uint8 CUDA tensors (no host round trip)."""
import numpy as np
import torch

from . import ops


def _to_dev(img, device):
    a = np.asarray(img.convert("RGB") if hasattr(img, "convert") else img, dtype=np.uint8)
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def color_fix_u8(target_u8, source_u8, mode):
    assert target_u8.shape == source_u8.shape and target_u8.dtype == torch.uint8 and target_u8.is_cuda
    return ops.color_fix(target_u8.contiguous(), source_u8.contiguous(), mode)


def _fix(target, source, mode, device="cuda"):
    from PIL import Image
    out = color_fix_u8(_to_dev(target, device), _to_dev(source, device), mode)
    return Image.fromarray(out.cpu().numpy())


def adain_color_fix(target, source):
    return _fix(target, source, "adain")


def wavelet_color_fix(target, source):
    return _fix(target, source, "wavelet")
