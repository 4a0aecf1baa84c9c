"""Host-side image helpers of the caller (utils/image_process.py:7-35): `check_image_size` pads a PIL image at the right /
bottom to a multiple of 8 by edge replication (cv2.BORDER_REPLICATE in the reference); u8 in, u8 out, no GPU work."""
import numpy as np
from PIL import Image


def check_image_size(x, padder_size=8):
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
