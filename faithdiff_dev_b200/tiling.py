"""Drop-in for `LocalAttention` (FaithDiff/pipelines/pipeline_FaithDiff_tlc.py:101-283): overlapping latent tiles with
a Gaussian-weighted blend.  Tile origins are host integer arithmetic; the weight table is produced on the host with
the SAME numpy float64 expression as the reference (numpy.exp is not bit-identical to libm/CUDA exp, so bit-exact
weights require this) and uploaded once; gather and blend run in the sm_100a kernels of blend.cu."""
import math

import numpy as np
import torch

from . import ops


def tile_origins(h, w, k1, k2, overlap):
    """pipeline…:172-204.  Returns ([(i, j)…] row-major, k1', k2')."""
    k1, k2 = min(k1, h), min(k2, w)
    num_row = (h - 1) // k1 + 1
    num_col = (w - 1) // k2 + 1
    step_j = k2 if num_col == 1 else math.ceil(k2 * overlap)
    step_i = k1 if num_row == 1 else math.ceil(k1 * overlap)
    out = []
    i, last_i = 0, False
    while i < h and not last_i:
        j = 0
        if i + k1 >= h:
            i, last_i = h - k1, True
        last_j = False
        while j < w and not last_j:
            if j + k2 >= w:
                j, last_j = w - k2, True
            out.append((i, j))
            j += step_j
        i += step_i
    return out, k1, k2


def gaussian_weights(tile_width, tile_height):
    """pipeline…:207-228 verbatim arithmetic: python floats + numpy.exp/sqrt/pi, np.outer → float64 [H, W].
    Note the reference's asymmetry: x midpoint (W-1)/2, y midpoint H/2."""
    from numpy import exp, pi, sqrt
    var = 0.01
    midpoint = (tile_width - 1) / 2
    x_probs = [exp(-(x - midpoint) * (x - midpoint) / (tile_width * tile_width) / (2 * var)) / sqrt(2 * pi * var)
               for x in range(tile_width)]
    midpoint = tile_height / 2
    y_probs = [exp(-(y - midpoint) * (y - midpoint) / (tile_height * tile_height) / (2 * var)) / sqrt(2 * pi * var)
               for y in range(tile_height)]
    return np.outer(y_probs, x_probs)


class LocalAttention:
    def __init__(self, kernel_size=None, overlap=0.5):
        self.kernel_size = kernel_size
        self.overlap = overlap
        self._cache = {}

    def _layout(self, h, w, device):
        key = (h, w, str(device))
        if key not in self._cache:
            idx, k1, k2 = tile_origins(h, w, self.kernel_size[0], self.kernel_size[1], self.overlap)
            wts = torch.from_numpy(gaussian_weights(k2, k1)).to(device)
            org = torch.tensor(idx, dtype=torch.int32, device=device).contiguous()
            self._cache[key] = (idx, k1, k2, wts, org)
        return self._cache[key]

    def grids(self, x):
        if not x.is_cuda:
            raise RuntimeError("faithdiff_dev_b200 has no CPU path: input must be a CUDA tensor")
        b, c, h, w = x.shape
        assert b == 1
        self.original_size = (b, c, h, w)
        idx, k1, k2, wts, org = self._layout(h, w, x.device)
        self.idxes = [{"i": i, "j": j} for i, j in idx]
        self.origins = org
        self.weights64 = wts
        self.tile_weights = wts.unsqueeze(0).expand(4, -1, -1)     # reference attribute (4, k1, k2) float64
        self.tile_hw = (k1, k2)
        return ops.tile_gather(x.contiguous(), org, k1, k2)

    def grids_inverse(self, outs):
        b, c, h, w = self.original_size
        canvas, _ = ops.tile_blend(outs.contiguous(), self.weights64, self.origins, h, w)
        return canvas

    def forward(self, x):
        return self.grids_inverse(self.grids(x))
