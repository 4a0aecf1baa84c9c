"""Pins oracle/bsrnet_oracle.py against the reference's OWN BSRNet and writes tests/golden/bsrnet.npz.

Run in the build container (needs /root/reference):  python oracle/make_golden_bsrnet.py
FaithDiff/models/bsrnet_arch.py imports without diffusers, so the unmodified reference class is executed on the CPU:
  * its state-dict keys / shapes must equal oracle.bsrnet_oracle.param_shapes (full nb=23 network and the small fixtures);
  * forward and deg_remove (several tiles, ragged sizes, tile_size 24 / pad 4 so the tiling logic is exercised at fixture
    size) must be reproduced BIT FOR BIT by the restatement;
  * the reference's outputs are stored (weights are regenerated from the seed by the tests).
"""
import importlib.util
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bsrnet_oracle as B  # noqa: E402

REF = "/root/reference/FaithDiff/models/bsrnet_arch.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_bsrnet_arch", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    ref = load_reference()
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        full = ref.RRDBNet(in_nc=3, out_nc=3, nf=64, nb=23, gc=32, sf=4)          # what create_bsrnet builds
    assert {k: tuple(v.shape) for k, v in full.state_dict().items()} == B.param_shapes()
    n_params = sum(v.numel() for v in full.state_dict().values())
    print("full BSRNet parameters:", n_params)
    out = {"n_params_full": np.int64(n_params)}
    cfg = dict(in_nc=3, out_nc=3, nf=64, nb=2, gc=32, sf=4)
    sd = B.random_state_dict(seed=11, **cfg)
    with redirect_stdout(io.StringIO()):
        net = ref.RRDBNet(**cfg).eval()
    net.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(5)
    cases = {"fwd_a": (1, 3, 20, 28), "fwd_ragged": (1, 3, 13, 18), "fwd_batch2": (2, 3, 8, 12)}
    with torch.no_grad():
        for name, shp in cases.items():
            x = torch.rand(shp, generator=g)
            y_ref = net(x)
            y = B.forward(sd, x)
            assert torch.equal(y, y_ref), name
            out[name + "_x"], out[name + "_y"] = x.numpy(), y_ref.numpy()
        for name, shp, ts, tp in [("deg_a", (1, 3, 40, 56), 24, 4), ("deg_ragged", (1, 3, 31, 50), 24, 4), ("deg_one_tile", (1, 3, 16, 20), 512, 16)]:
            x = torch.rand(shp, generator=g)
            y_ref = net.deg_remove(x, tile_size=ts, tile_pad=tp)
            y = B.deg_remove(sd, x, tile_size=ts, tile_pad=tp)
            assert torch.equal(y, y_ref), name
            out[name + "_x"], out[name + "_y"], out[name + "_tile"] = x.numpy(), y_ref.numpy(), np.array([ts, tp])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bsrnet.npz"), **out)
    print("restatement == reference (bit-exact) on", len(cases) + 3, "cases; wrote tests/golden/bsrnet.npz")


if __name__ == "__main__":
    main()
