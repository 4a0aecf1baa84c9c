"""Device PNG encoder through the C-ABI (fd_png_encode): the file is byte-identical to the host emulation of the same
per-thread code (tests/png_host_emul.cpp, itself checked against zlib / PIL in test_cpu_png.py) and PIL — the reference's
own codec for `out_image.save(path)`, test_wo_llava.py:97 — reads back the exact pixels."""
import io

import numpy as np
import pytest
import torch
from PIL import Image

import png_emul

pytestmark = pytest.mark.gpu
IMAGES = png_emul.test_images()


def _decode(data):
    im = Image.open(io.BytesIO(data))
    im.load()
    return np.asarray(im)


@pytest.mark.parametrize("name", sorted(IMAGES))
def test_png_equals_host_emulation_and_decodes(name):
    import faithdiff_dev_b200 as fd
    a = IMAGES[name]
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if name == "crop_view":                                   # a strided device view, not a copy
        full = torch.from_numpy(np.ascontiguousarray(a.base if a.base is not None else a)).cuda()
        t = full[3:515, 5:600]
        assert not t.is_contiguous() and (t.cpu().numpy() == a).all()
    data = fd.encode_png(t)
    want, _ = png_emul.encode(np.ascontiguousarray(a))
    assert data == want
    a3 = a if a.ndim == 3 else a[..., None]
    assert (_decode(data).reshape(a3.shape) == a3).all()


def test_png_4k_roundtrip_and_file(tmp_path, parity_record):
    """The benchmark's output size (3840 x 2160 RGB): lossless through PIL, written with save_png; sizes recorded."""
    import faithdiff_dev_b200 as fd
    g = torch.Generator(device="cuda").manual_seed(0)
    yy, xx = torch.meshgrid(torch.arange(2160, device="cuda"), torch.arange(3840, device="cuda"), indexing="ij")
    base = torch.stack([128 + 100 * torch.sin(xx / 37.0 + k) * torch.cos(yy / 23.0 - k) for k in range(3)], -1)
    img = (base + 4 * torch.randn(base.shape, device="cuda", generator=g)).clamp(0, 255).to(torch.uint8)
    img[:200, :500] = 255                                      # a saturated patch
    path = tmp_path / "out.png"
    fd.save_png(img, str(path))
    back = np.asarray(Image.open(str(path)).convert("RGB"))
    host = img.cpu().numpy()
    assert back.shape == host.shape and (back == host).all()
    buf = io.BytesIO()
    Image.fromarray(host).save(buf, format="PNG")
    ours, pil = path.stat().st_size, len(buf.getvalue())
    parity_record("png_4k", lossless=True, file_bytes=ours, pil_file_bytes=pil, raw_bytes=host.size)
    assert ours <= 1.1 * pil
    assert fd.encode_png(img) == png_emul.encode(host)[0]      # deterministic, and the same bytes as the host emulation


def test_png_accepts_pil_and_host_arrays():
    import faithdiff_dev_b200 as fd
    a = IMAGES["ragged_rgb"]
    want = png_emul.encode(a)[0]
    assert fd.encode_png(Image.fromarray(a)) == want
    assert fd.encode_png(a) == want
    assert fd.encode_png(torch.from_numpy(a)) == want
    grey = IMAGES["one_row_grey"]
    assert (_decode(fd.encode_png(grey[..., 0])) == grey[..., 0]).all()
    with pytest.raises(ValueError):
        fd.encode_png(torch.zeros(4, 4, 5, dtype=torch.uint8, device="cuda"))
