"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain torch fp32, functional) of the reference's BSRNet pre-clean network.

Follows FaithDiff/models/bsrnet_arch.py:
  * dense block (:35-55): five 3x3 convs, each fed the channel concatenation of the block input and all earlier growth
    outputs, LeakyReLU(0.2) after the first four, output = conv5 * 0.2 + input;
  * RRDB (:58-72): three dense blocks, output * 0.2 + input;
  * network (:75-115): zero-pad H, W to a multiple of 4 (right / bottom), conv_first, `nb` RRDBs, trunk_conv, long skip,
    two (nearest x2 -> conv -> LeakyReLU) stages, HRconv + LeakyReLU, conv_last, crop to 4H x 4W;
  * `deg_remove` (:117-175): 512-px tiles with a 16-px halo, each tile run independently, the un-padded centre pasted.
Operates directly on a state dict with the reference's key names (what `create_bsrnet` loads, create_FaithDiff_model.py:39-42).

PARITY PINNED: unlike the diffusers-based modules this file of the reference imports in the build container, so
`oracle/make_golden_bsrnet.py` runs the reference's own RRDBNet (forward and deg_remove) and asserts this restatement
reproduces it bit for bit, then stores the reference's outputs in tests/golden/bsrnet.npz.
Only tests/ may import this module; the product (faithdiff_dev_b200/bsrnet.py) never does.
"""
import math

import torch
import torch.nn.functional as F


def param_shapes(in_nc=3, out_nc=3, nf=64, nb=23, gc=32, sf=4):
    """Key -> shape of the reference state dict (bsrnet_arch.py:38-43, :82-92)."""
    shapes = {"conv_first.weight": (nf, in_nc, 3, 3), "conv_first.bias": (nf,)}
    for i in range(nb):
        for r in (1, 2, 3):
            for k in range(1, 6):
                cin, cout = nf + (k - 1) * gc, (gc if k < 5 else nf)
                shapes["RRDB_trunk.%d.RDB%d.conv%d.weight" % (i, r, k)] = (cout, cin, 3, 3)
                shapes["RRDB_trunk.%d.RDB%d.conv%d.bias" % (i, r, k)] = (cout,)
    names = ["trunk_conv", "upconv1"] + (["upconv2"] if sf == 4 else []) + ["HRconv"]
    for n in names:
        shapes[n + ".weight"] = (nf, nf, 3, 3)
        shapes[n + ".bias"] = (nf,)
    shapes["conv_last.weight"] = (out_nc, nf, 3, 3)
    shapes["conv_last.bias"] = (out_nc,)
    return shapes


def random_state_dict(seed=0, w_std=None, **cfg):
    """Deterministic weights: kaiming-like std (x0.1 inside the dense blocks, as bsrnet_arch.py:46 initialises them) and
    small random biases so that the bias paths are exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(**cfg).items():
        if k.endswith("weight"):
            fan_in = shp[1] * 9
            std = w_std if w_std is not None else math.sqrt(2.0 / fan_in) * (0.1 if ".RDB" in k else 1.0)
            sd[k] = torch.randn(shp, generator=g) * std
        else:
            sd[k] = torch.randn(shp, generator=g) * 0.05
    return sd


def _conv(sd, name, x):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=1)


def _lrelu(x):
    return F.leaky_relu(x, 0.2)


def dense_block(sd, prefix, x):
    feats = [x]
    for k in range(1, 5):
        feats.append(_lrelu(_conv(sd, "%s.conv%d" % (prefix, k), torch.cat(feats, 1))))
    return _conv(sd, prefix + ".conv5", torch.cat(feats, 1)) * 0.2 + x


def rrdb(sd, prefix, x):
    out = x
    for r in (1, 2, 3):
        out = dense_block(sd, "%s.RDB%d" % (prefix, r), out)
    return out * 0.2 + x


def forward(sd, x, sf=4):
    nb = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("RRDB_trunk."))
    _, _, H, W = x.shape
    x = F.pad(x, (0, (4 - W % 4) % 4, 0, (4 - H % 4) % 4))
    fea = _conv(sd, "conv_first", x)
    t = fea
    for i in range(nb):
        t = rrdb(sd, "RRDB_trunk.%d" % i, t)
    fea = fea + _conv(sd, "trunk_conv", t)
    fea = _lrelu(_conv(sd, "upconv1", F.interpolate(fea, scale_factor=2, mode="nearest")))
    if sf == 4:
        fea = _lrelu(_conv(sd, "upconv2", F.interpolate(fea, scale_factor=2, mode="nearest")))
    out = _conv(sd, "conv_last", _lrelu(_conv(sd, "HRconv", fea)))
    return out[:, :, :H * 4, :W * 4]


def tile_plan(height, width, tile_size=512, tile_pad=16):
    """[(y0, y1, x0, x1, y0_pad, y1_pad, x0_pad, x1_pad)] in the reference's loop order (bsrnet_arch.py:131-150)."""
    plan = []
    for ty in range(math.ceil(height / tile_size)):
        for tx in range(math.ceil(width / tile_size)):
            x0, y0 = tx * tile_size, ty * tile_size
            x1, y1 = min(x0 + tile_size, width), min(y0 + tile_size, height)
            plan.append((y0, y1, x0, x1, max(y0 - tile_pad, 0), min(y1 + tile_pad, height), max(x0 - tile_pad, 0),
                         min(x1 + tile_pad, width)))
    return plan


@torch.no_grad()
def deg_remove(sd, x, tile_size=512, tile_pad=16, sf=4):
    b, c, h, w = x.shape
    out = x.new_zeros((b, c, h * 4, w * 4))
    for (y0, y1, x0, x1, y0p, y1p, x0p, x1p) in tile_plan(h, w, tile_size, tile_pad):
        o = forward(sd, x[:, :, y0p:y1p, x0p:x1p], sf)
        oy, ox = (y0 - y0p) * 4, (x0 - x0p) * 4
        out[:, :, y0 * 4:y1 * 4, x0 * 4:x1 * 4] = o[:, :, oy:oy + (y1 - y0) * 4, ox:ox + (x1 - x0) * 4]
    return out
