"""BSRNet pre-clean network (RRDBNet) on the B200 kernels — drop-in for FaithDiff/models/bsrnet_arch.py:75-175 and
`create_bsrnet` (create_FaithDiff_model.py:39-42); the caller is test.py:98-102 (`bsrnet.deg_remove(image_tensor)`).

Same constructor arguments, state-dict key names, `forward` / `deg_remove` signatures and tiling semantics as the reference;
NCHW float tensors in [0,1] in, NCHW float32 out.  Inside, activations are NHWC bf16 and every convolution is one launch of
the path's kernels:

* a dense block (bsrnet_arch.py:35-55) lives in ONE [B,H,W,nf+4*gc] buffer: conv_k reads the first nf+(k-1)*gc channels as a
  strided TMA view and writes its gc outputs (bias + LeakyReLU(0.2) in the epilogue) into the next channel slice, so none of
  the five `torch.cat`s exists; conv5's epilogue does `* 0.2 + x` (alpha, pre-scaled bias, residual) and stores straight into
  the next block's buffer;
* the RRDB skip `out * 0.2 + x` (:72) is one fd_axpby_bf16 on channel-slice views;
* conv_first (3 -> nf) and conv_last (nf -> 3) use the small-channel direct kernels, the two nearest x2 stages
  fd_upsample_nearest_nhwc, the long skip `fea + trunk` is the residual of trunk_conv's epilogue.

There is no CPU path: CPU tensors raise RuntimeError.
"""
import math

import torch

from . import ops
from .blocks import Weights

BF16 = torch.bfloat16
ACT_LRELU = 5


class RRDBNet:
    def __init__(self, in_nc=3, out_nc=3, nf=64, nb=23, gc=32, sf=4):
        if in_nc not in (3, 4, 8) or out_nc > 8 or nf % 8 or gc % 8:
            raise NotImplementedError("RRDBNet on the B200 path: in_nc in (3, 4, 8), out_nc <= 8, nf and gc multiples of 8")
        if sf not in (2, 4):
            raise ValueError("sf must be 2 or 4")
        self.in_nc, self.out_nc, self.nf, self.nb, self.gc, self.sf = in_nc, out_nc, nf, nb, gc, sf
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        self._loaded = False

    # ------------------------------------------------------------------ weights
    def expected_keys(self):
        keys = ["conv_first", "trunk_conv", "upconv1", "HRconv", "conv_last"] + (["upconv2"] if self.sf == 4 else [])
        keys += ["RRDB_trunk.%d.RDB%d.conv%d" % (i, r, k) for i in range(self.nb) for r in (1, 2, 3) for k in range(1, 6)]
        return sorted(k + s for k in keys for s in (".weight", ".bias"))

    def load_state_dict(self, state_dict, strict=True):
        if self.device is None:
            raise RuntimeError("faithdiff_dev_b200.RRDBNet needs a CUDA device (there is no CPU path)")
        missing = [k for k in self.expected_keys() if k not in state_dict]
        unexpected = [k for k in state_dict if k not in set(self.expected_keys())]
        if strict and (missing or unexpected):
            raise RuntimeError("Error(s) in loading state_dict for RRDBNet: missing %s unexpected %s" % (missing[:5], unexpected[:5]))
        W = Weights(state_dict, self.device)
        conv = lambda n: (W.conv3(n + ".weight"), W.f32(n + ".bias"))
        self.conv_first, self.trunk_conv, self.upconv1, self.hrconv, self.conv_last = (
            conv("conv_first"), conv("trunk_conv"), conv("upconv1"), conv("HRconv"), conv("conv_last"))
        self.upconv2 = conv("upconv2") if self.sf == 4 else None
        self.rdbs = []
        for i in range(self.nb):
            for r in (1, 2, 3):
                cs = [conv("RRDB_trunk.%d.RDB%d.conv%d" % (i, r, k)) for k in range(1, 6)]
                cs[4] = (cs[4][0], cs[4][1] * 0.2)            # (conv5 + b) * 0.2 = 0.2*acc + 0.2*b: alpha on the accumulator
                self.rdbs.append(cs)
        self._loaded = True
        return self

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------ forward
    def _dense_block(self, convs, buf, out):
        """buf [B,H,W,nf+4gc] with the block input in [..., :nf]; writes conv5*0.2 + input to `out` (a [..., :nf] view)."""
        nf, gc = self.nf, self.gc
        for k in range(4):
            w, b = convs[k]
            ops.conv3x3(buf[..., :nf + k * gc], w, bias=b, act=ACT_LRELU, out=buf[..., nf + k * gc:nf + (k + 1) * gc])
        w, b = convs[4]
        ops.conv3x3(buf, w, bias=b, residual=buf[..., :nf], alpha=0.2, out=out)

    def forward(self, x):
        if not self._loaded:
            raise RuntimeError("RRDBNet: load_state_dict first")
        if not x.is_cuda:
            raise RuntimeError("faithdiff_dev_b200.RRDBNet runs on CUDA tensors only (there is no CPU path)")
        if x.dim() != 4 or x.shape[1] != self.in_nc:
            raise ValueError("expected a [B,%d,H,W] tensor, got %s" % (self.in_nc, tuple(x.shape)))
        B, _, H, W = x.shape
        ph, pw = (4 - H % 4) % 4, (4 - W % 4) % 4                 # check_image_size (:100-106): zero pad right / bottom
        if ph or pw:
            x = torch.nn.functional.pad(x, (0, pw, 0, ph))
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        h, w = H + ph, W + pw
        nf, gc = self.nf, self.gc
        wide = nf + 4 * gc
        bufs = [torch.empty((B, h, w, wide), device=x.device, dtype=BF16) for _ in range(3)]
        ops.conv3x3_small_in(x, *self.conv_first, out=bufs[0][..., :nf])
        fea = bufs[0][..., :nf].contiguous()                        # long skip (:110-111)
        tmp = torch.empty((B, h, w, nf), device=x.device, dtype=BF16)
        for i in range(self.nb):
            r1, r2, r3 = self.rdbs[3 * i:3 * i + 3]
            self._dense_block(r1, bufs[0], bufs[1][..., :nf])
            self._dense_block(r2, bufs[1], bufs[2][..., :nf])
            self._dense_block(r3, bufs[2], tmp)
            ops.axpby(tmp, bufs[0][..., :nf], 0.2, 1.0, out=bufs[1][..., :nf])      # RRDB: out * 0.2 + x (:72)
            bufs = [bufs[1], bufs[2], bufs[0]]
        trunk_in = bufs[0][..., :nf]
        fea = ops.conv3x3(trunk_in, self.trunk_conv[0], bias=self.trunk_conv[1], residual=fea)
        del bufs, tmp
        fea = ops.conv3x3(ops.upsample_nearest(fea), self.upconv1[0], bias=self.upconv1[1], act=ACT_LRELU)
        if self.sf == 4:
            fea = ops.conv3x3(ops.upsample_nearest(fea), self.upconv2[0], bias=self.upconv2[1], act=ACT_LRELU)
        fea = ops.conv3x3(fea, self.hrconv[0], bias=self.hrconv[1], act=ACT_LRELU)
        out = ops.conv3x3_small_out(fea, *self.conv_last, out_dtype=torch.float32, nchw_out=True)
        return out[:, :, :H * 4, :W * 4]                           # the reference crops to 4x whatever sf is (:115)

    __call__ = forward

    @torch.no_grad()
    def deg_remove(self, input, tile_size=512, tile_pad=16):
        """Tile loop of bsrnet_arch.py:117-175: tiles of `tile_size` with a `tile_pad` halo, the un-padded centre of each
        4x result pasted into the output."""
        if not input.is_cuda:
            raise RuntimeError("faithdiff_dev_b200.RRDBNet runs on CUDA tensors only (there is no CPU path)")
        batch, channel, height, width = input.shape
        output = torch.zeros((batch, channel, height * 4, width * 4), device=input.device, dtype=torch.float32)
        for ty in range(math.ceil(height / tile_size)):
            for tx in range(math.ceil(width / tile_size)):
                x0, y0 = tx * tile_size, ty * tile_size
                x1, y1 = min(x0 + tile_size, width), min(y0 + tile_size, height)
                x0p, x1p = max(x0 - tile_pad, 0), min(x1 + tile_pad, width)
                y0p, y1p = max(y0 - tile_pad, 0), min(y1 + tile_pad, height)
                tile = self.forward(input[:, :, y0p:y1p, x0p:x1p])
                oy, ox = (y0 - y0p) * 4, (x0 - x0p) * 4
                output[:, :, y0 * 4:y1 * 4, x0 * 4:x1 * 4] = tile[:, :, oy:oy + (y1 - y0) * 4, ox:ox + (x1 - x0) * 4]
        return output


def create_bsrnet(bsrnet_path):
    """create_FaithDiff_model.py:39-42."""
    net = RRDBNet(in_nc=3, out_nc=3, nf=64, nb=23, gc=32, sf=4)
    net.load_state_dict(torch.load(bsrnet_path, map_location="cpu"), strict=True)
    return net
