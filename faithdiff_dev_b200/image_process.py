"""Image helpers of the caller.

`check_image_size` (utils/image_process.py:7-35) pads a PIL image at the right / bottom to a multiple of 8 by edge
replication (cv2.BORDER_REPLICATE in the reference); u8 in, u8 out, no GPU work.

`resize_lanczos` is the caller's `image.resize((w*scale, h*scale), Image.LANCZOS)` (test_wo_llava.py:78-81,
test.py:104-108, gradio_demo.py) on the GPU: the tap tables are computed here exactly as Pillow's
`precompute_coeffs` / `normalize_coeffs_8bpc` do (libm `sin`, 22-bit fixed point) and the two integer passes run as
CUDA kernels (fd_resample_u8_*), so the result is bit-identical to PIL and only the LQ image crosses PCIe.
"""
import functools
import math

import numpy as np
from PIL import Image

_PRECISION_BITS = 32 - 8 - 2          # Pillow: 8-bit samples, 2 guard bits for the negative lobes


def check_image_size(x, padder_size=8):
    """utils/image_process.py:7-35.  x: PIL image (as in the reference) or a u8 HWC tensor (e.g. the device result of
    `resize_lanczos`, which then never leaves the GPU); returns (padded image of the same kind, w0, h0, w, h)."""
    if not hasattr(x, "size") or callable(getattr(x, "size")):       # torch tensor: .size is a method, PIL: a (w, h) tuple
        import torch
        assert x.dim() == 3 and x.dtype == torch.uint8, "u8 HWC tensor expected"
        height, width = x.shape[0], x.shape[1]
        mod_pad_h = (padder_size - height % padder_size) % padder_size
        mod_pad_w = (padder_size - width % padder_size) % padder_size
        if mod_pad_w:
            x = torch.cat([x, x[:, -1:].expand(-1, mod_pad_w, -1)], dim=1)
        if mod_pad_h:
            x = torch.cat([x, x[-1:].expand(mod_pad_h, -1, -1)], dim=0)
        return x, width, height, width + mod_pad_w, height + mod_pad_h
    width, height = x.size
    mod_pad_h = (padder_size - height % padder_size) % padder_size
    mod_pad_w = (padder_size - width % padder_size) % padder_size
    a = np.array(x)
    pad = ((0, mod_pad_h), (0, mod_pad_w)) + ((0, 0),) * (a.ndim - 2)
    return Image.fromarray(np.pad(a, pad, mode="edge")), width, height, width + mod_pad_w, height + mod_pad_h


def image2tensor(img):
    import torch
    if img.ndim == 2:
        img = np.expand_dims(img, axis=2)
    return torch.from_numpy(np.ascontiguousarray(img)).permute(2, 0, 1).float().div(255.).unsqueeze(0)


def tensor2image(img):
    img = img.data.squeeze().float().clamp_(0, 1).cpu().numpy()
    if img.ndim == 3:
        img = np.transpose(img, (1, 2, 0))
    return np.uint8((img * 255.0).round())


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    return _sinc(x) * _sinc(x / 3.0) if -3.0 <= x < 3.0 else 0.0


@functools.lru_cache(maxsize=64)
def resample_taps(in_size, out_size):
    """Pillow's LANCZOS tap table for one axis (full-image box).  Returns (bounds int32 [out,2] = first source index and
    tap count, taps int32 [out,ksize]).  Scalar Python floats + math.sin on purpose: the C code evaluates libm `sin` in
    double precision one tap at a time, and a vectorised sin is not bitwise the same function."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), np.int32)
    taps = np.zeros((out_size, ksize), np.int32)
    one = float(1 << _PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        taps[xx, :xmax] = [int(-0.5 + v * one) if v < 0 else int(0.5 + v * one) for v in w]
        bounds[xx] = (xmin, xmax)
    bounds.setflags(write=False)
    taps.setflags(write=False)
    return bounds, taps


@functools.lru_cache(maxsize=16)
def _device_tables(in_w, in_h, out_w, out_h, device):
    import torch
    horiz = vert = None
    row0, rows = 0, in_h
    if out_h != in_h:
        vb, vk = resample_taps(in_h, out_h)
        if out_w != in_w:                      # the horizontal pass only produces the rows the vertical pass reads
            row0 = int(vb[0, 0])
            rows = int(vb[-1, 0] + vb[-1, 1]) - row0
            vb = vb - np.array([row0, 0], np.int32)
        vert = (torch.from_numpy(np.array(vb)).to(device), torch.from_numpy(np.array(vk)).to(device))
    if out_w != in_w:
        hb, hk = resample_taps(in_w, out_w)
        horiz = (torch.from_numpy(np.array(hb)).to(device), torch.from_numpy(np.array(hk)).to(device))
    return horiz, vert, row0, rows


def resize_lanczos(image, size, device="cuda"):
    """`image.resize(size, Image.LANCZOS)` on the GPU, bit-identical to PIL for 8-bit images.
    image: PIL image (L / RGB / CMYK ...; not RGBA: PIL premultiplies alpha), or u8 HWC numpy array / torch tensor (host or device); size = (width, height).
    Returns a u8 HWC tensor on `device` (feed it to the pipeline as `lr_img`, or `Image.fromarray(t.cpu().numpy())`)."""
    import torch
    from . import ops
    if isinstance(image, Image.Image):
        if image.mode not in ("L", "RGB", "CMYK", "YCbCr", "HSV", "LAB"):
            # RGBA / LA are excluded on purpose: PIL premultiplies alpha around the resample; the path feeds RGB
            raise ValueError("resize_lanczos: PIL mode %r is not a plain 8-bit interleaved image" % (image.mode,))
        image = np.array(image)
    t = torch.as_tensor(image)
    if t.dtype != torch.uint8 or t.dim() not in (2, 3):
        raise ValueError("resize_lanczos expects a u8 HW or HWC image, got %s %s" % (t.dtype, tuple(t.shape)))
    squeeze = t.dim() == 2
    if squeeze:
        t = t.unsqueeze(-1)
    if t.shape[-1] > 4:
        raise ValueError("resize_lanczos: at most 4 interleaved channels")
    out_w, out_h = int(size[0]), int(size[1])
    if out_w <= 0 or out_h <= 0:
        raise ValueError("height and width must be > 0")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("resize_lanczos runs on a CUDA device only (there is no CPU path; use PIL on the host)")
    t = t.to(dev, non_blocking=True).contiguous()
    in_h, in_w = t.shape[0], t.shape[1]
    horiz, vert, row0, rows = _device_tables(in_w, in_h, out_w, out_h, str(t.device))
    if horiz is not None:
        t = ops.resample_u8_horizontal(t, horiz[0], horiz[1], row0, rows)
    if vert is not None:
        t = ops.resample_u8_vertical(t, vert[0], vert[1])
    return t.squeeze(-1) if squeeze else t
