"""-m gpu: BSRNet pre-clean (RRDBNet) on the B200 kernels against (a) the REAL reference's outputs stored in
tests/golden/bsrnet.npz (FaithDiff/models/bsrnet_arch.py run by oracle/make_golden_bsrnet.py) and (b) the oracle
restatement on fresh inputs.  bf16 activations vs the reference's fp32: rel-L2 <= 1.5e-2 and PSNR >= 45 dB relative to the output's value range
(measured on B200: rel-L2 5.4e-3 at nb=2, 9.6e-3 at nb=23)."""
import os

import numpy as np
import pytest
import torch

from oracle import bsrnet_oracle as B

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bsrnet.npz")
CFG = dict(in_nc=3, out_nc=3, nf=64, nb=2, gc=32, sf=4)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def psnr(a, b):
    """PSNR against the reference's own value range (random-init weights do not produce [0,1] images)."""
    rng = (b.float().max() - b.float().min()).clamp_min(1e-20)
    return (10 * torch.log10(rng ** 2 / ((a.float() - b.float()) ** 2).mean().clamp_min(1e-30))).item()


@pytest.fixture(scope="module")
def net():
    import faithdiff_dev_b200 as fd
    n = fd.RRDBNet(**CFG)
    n.load_state_dict(B.random_state_dict(seed=11, **CFG))
    return n


@pytest.mark.parametrize("case", ["fwd_a", "fwd_ragged", "fwd_batch2"])
def test_forward_matches_reference_golden(net, case):
    g = np.load(GOLD)
    x, y = torch.from_numpy(g[case + "_x"]), torch.from_numpy(g[case + "_y"])
    out = net(x.cuda())
    assert out.shape == y.shape and out.dtype == torch.float32
    assert rel(out.cpu(), y) < 1.5e-2 and psnr(out.cpu(), y) > 45.0, (rel(out.cpu(), y), psnr(out.cpu(), y))


@pytest.mark.parametrize("case", ["deg_a", "deg_ragged", "deg_one_tile"])
def test_deg_remove_matches_reference_golden(net, case):
    g = np.load(GOLD)
    x, y = torch.from_numpy(g[case + "_x"]), torch.from_numpy(g[case + "_y"])
    ts, tp = (int(v) for v in g[case + "_tile"])
    out = net.deg_remove(x.cuda(), tile_size=ts, tile_pad=tp)
    assert out.shape == y.shape
    assert rel(out.cpu(), y) < 1.5e-2 and psnr(out.cpu(), y) > 45.0, (rel(out.cpu(), y), psnr(out.cpu(), y))


def test_full_depth_network_and_default_tiling():
    """nb=23 (what create_bsrnet builds) on a 2-tile input with the default 512 / 16 tiling: oracle restatement in fp32."""
    import faithdiff_dev_b200 as fd
    cfg = dict(CFG, nb=23)
    sd = B.random_state_dict(seed=3, **cfg)
    n = fd.RRDBNet(**cfg)
    n.load_state_dict(sd)
    assert sorted(sd) == n.expected_keys()
    x = torch.rand(1, 3, 40, 530, generator=torch.Generator().manual_seed(1))
    ref = B.deg_remove(sd, x)
    out = n.deg_remove(x.cuda())
    assert rel(out.cpu(), ref) < 2e-2 and psnr(out.cpu(), ref) > 42.0, (rel(out.cpu(), ref), psnr(out.cpu(), ref))


def test_errors():
    import faithdiff_dev_b200 as fd
    n = fd.RRDBNet(**CFG)
    with pytest.raises(RuntimeError):
        n(torch.zeros(1, 3, 8, 8).cuda())                   # weights not loaded
    sd = B.random_state_dict(seed=11, **CFG)
    sd.pop("HRconv.bias")
    with pytest.raises(RuntimeError):
        n.load_state_dict(sd)
    n.load_state_dict(B.random_state_dict(seed=11, **CFG))
    with pytest.raises(RuntimeError):
        n(torch.zeros(1, 3, 8, 8))                          # CPU tensor: no fallback
