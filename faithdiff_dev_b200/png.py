"""PNG output of the caller on the GPU: `out_image.save(path)` (test_wo_llava.py:97, test.py:130).

`encode_png` / `save_png` take the restored image where it already is — a u8 HWC CUDA tensor (`pipe(..., output_type="pt")`
→ crop → `color_fix_u8`) — filter, entropy-code and frame it on the device (csrc/png.cu: five launches, complete file incl.
chunk CRCs and the Adler-32), and copy only the finished file to the host: ~half the bytes of the raw image, and none of the
1-2 s PIL's encoder + the host's zlib spend on a 3840 × 2160 image.  Lossless: any PNG reader returns the input bytes; the
file differs from PIL's (other filter / deflate choices), as two encoders' files always do.
PNG *decode* of the LQ input stays on the host (`Image.open`): an inflate stream is one serial dependency chain, the LQ image
is a quarter of the output, and its decode overlaps the previous image's restoration.
"""
import threading

import numpy as np
import torch

from . import ops


def _as_device_u8(image, device):
    if hasattr(image, "convert") and hasattr(image, "mode"):            # PIL image
        if image.mode not in ("L", "LA", "RGB", "RGBA"):
            image = image.convert("RGB")
        image = np.asarray(image)
    t = torch.as_tensor(image)
    if t.dtype != torch.uint8 or t.dim() not in (2, 3):
        raise ValueError("encode_png expects a u8 HW or HWC image, got %s %s" % (t.dtype, tuple(t.shape)))
    if t.dim() == 2:
        t = t.unsqueeze(-1)
    if not 1 <= t.shape[-1] <= 4:
        raise ValueError("encode_png: 1 (grey), 2 (grey + alpha), 3 (RGB) or 4 (RGBA) interleaved channels, got %d" % t.shape[-1])
    if t.shape[0] == 0 or t.shape[1] == 0:
        raise ValueError("encode_png: empty image")
    if not t.is_cuda:
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("encode_png runs on a CUDA device only (there is no CPU path; use PIL on the host)")
        t = t.to(dev, non_blocking=True)
    return t


_pinned = {}
_pinned_lock = threading.Lock()          # the staging buffer is shared: one encode at a time per process


def _encode_to_pinned(image, device):
    """Runs the encoder and returns the file as a numpy view of a cached pinned host buffer (valid until the next call)."""
    t = _as_device_u8(image, device)
    with torch.cuda.device(t.device):
        out, result = ops.png_encode(t)
        size, overflow = result.tolist()                               # first synchronisation: 16 bytes
        if overflow:
            raise RuntimeError("encode_png: output capacity exceeded (%d bytes needed)" % size)
        host = _pinned.get("buf")
        if host is None or host.numel() < size:
            host = _pinned["buf"] = torch.empty(max(size, out.numel()), dtype=torch.uint8, pin_memory=True)
        host[:size].copy_(out[:size], non_blocking=True)               # only the finished file crosses PCIe, into pinned memory
        torch.cuda.current_stream().synchronize()
        return host.numpy()[:size]


def encode_png(image, device="cuda"):
    """PNG file (bytes) of `image`: u8 HWC / HW CUDA tensor (rows may be a strided view), or a PIL image / numpy array / host
    tensor that is uploaded first."""
    with _pinned_lock:
        return _encode_to_pinned(image, device).tobytes()


def save_png(image, fp, device="cuda"):
    """`image.save(fp)` for PNG: fp is a path or a binary file object."""
    with _pinned_lock:
        data = memoryview(_encode_to_pinned(image, device))
        if hasattr(fp, "write"):
            fp.write(data)
        else:
            with open(fp, "wb") as f:
                f.write(data)
