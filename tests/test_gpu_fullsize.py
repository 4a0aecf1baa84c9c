"""-m gpu: parity at BASELINE.json's FULL architecture (SDXL-size UNet + FaithDiff stem, 2.6 B parameters, 128x128 latent
tile, CFG batch 2) against the fp32 oracle evaluated on the GPU, plus size-independent properties of the tiled path at the
4K / 8K layouts."""
import copy

import numpy as np
import pytest
import torch

from oracle import faithdiff_oracle as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


@pytest.fixture(scope="module")
def full():
    import faithdiff_dev_b200 as fd
    from faithdiff_dev_b200 import arch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = arch.random_state_dict(arch.unet_param_shapes(arch.SDXL_UNET_CONFIG), seed=1234)
    unet = fd.UNet2DConditionModel(**arch.SDXL_UNET_CONFIG)
    unet.load_state_dict(sd)
    with torch.device("meta"):
        ora = O.FaithDiffUNet(**O.SDXL_UNET_CONFIG)
        ora.load_additional_layers(channel=512)
    ora = ora.to_empty(device="cuda")
    ora.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)      # identical (bf16-representable) weights
    del sd
    return fd, unet, ora.eval()


@pytest.mark.timeout(900)
def test_fullsize_unet_tile_step_matches_fp32_oracle(full):
    fd, unet, ora = full
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(2, 4, 128, 128, generator=g, device="cuda").to(BF)
    lq = torch.randn(2, 512, 128, 128, generator=g, device="cuda").to(BF)
    ehs = torch.randn(2, 77, 2048, generator=g, device="cuda").to(BF)
    te = torch.randn(2, 1280, generator=g, device="cuda").to(BF)
    ids = torch.tensor([[2160., 3840., 142., 352., 1024., 1024.]] * 2, device="cuda")
    for t in (951.0, 1.0):
        with torch.no_grad():
            ref = ora(x.float(), torch.tensor(t), encoder_hidden_states=ehs.float(), input_embedding=lq.float(),
                      added_cond_kwargs={"text_embeds": te.float(), "time_ids": ids}, return_dict=False)[0]
        out = unet(x, torch.tensor(t), encoder_hidden_states=ehs, input_embedding=lq,
                   added_cond_kwargs={"text_embeds": te, "time_ids": ids}, return_dict=False)[0]
        e = rel(out, ref)
        # the same network evaluated by torch in bf16 (the reference's own library path) as the accuracy yardstick
        ob = copy.deepcopy(ora).to(BF)
        with torch.no_grad():
            rb = ob(x, torch.tensor(t), encoder_hidden_states=ehs, input_embedding=lq,
                    added_cond_kwargs={"text_embeds": te, "time_ids": ids.to(BF)}, return_dict=False)[0]
        del ob
        eb = rel(rb, ref)
        print("full-size UNet t=%g: cuda path rel-L2 %.3e | torch bf16 %.3e" % (t, e, eb))
        assert torch.isfinite(out.float()).all()
        assert e < max(1e-2, 1.25 * eb), (e, eb)


@pytest.mark.timeout(900)
def test_fullsize_tiled_step_properties(full):
    """Size-independent properties at the 4K layout (28 tiles): (i) a constant latent field is a fixed point of
    gather → blend; (ii) blending is linear; (iii) N-rank unit accumulation == the fused single-GPU tail within fp32
    summation-order error — on the real 270x480 canvas."""
    fd, unet, ora = full
    from faithdiff_dev_b200 import ops
    from faithdiff_dev_b200.scheduler import EulerDiscreteScheduler
    h, w, k = 270, 480, 128
    la = fd.LocalAttention((k, k), 0.5)
    c = torch.full((1, 4, h, w), 0.7265625, device="cuda", dtype=BF)
    tiles = la.grids(c)
    assert tiles.shape[0] == 28
    assert torch.equal(la.grids_inverse(tiles).to(BF), c)
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randn(tiles.shape, generator=g, device="cuda").to(BF)
    b = torch.randn(tiles.shape, generator=g, device="cuda").to(BF)
    lhs = la.grids_inverse((a.float() + b.float()).to(torch.float32))
    rhs = la.grids_inverse(a.float()) + la.grids_inverse(b.float())
    assert (lhs - rhs).abs().max() < 1e-5
    sch = EulerDiscreteScheduler(); sch.set_timesteps(20)
    eps = torch.randn((28, 2, 4, k, k), generator=g, device="cuda").to(BF)
    x = (torch.randn((28, 4, k, k), generator=g, device="cuda") * 5).to(BF)
    s, sn, _ = sch.step_scalars(5)
    ref, _, _, _ = ops.cfg_euler_blend(eps, x, la.weights64, la.origins, h, w, 5.0, s, sn, sch.step_scalars(6)[2])
    cnt = torch.zeros(1, 4, h, w, device="cuda")
    ops.tile_accumulate(torch.ones(28, 4, k, k, device="cuda"), None, la.weights64, la.origins, cnt)
    num = torch.zeros(1, 4, h, w, device="cuda")
    from faithdiff_dev_b200.dist import shard_units
    for r in range(8):
        units = shard_units(28, r, 8)
        part = torch.zeros_like(num)
        ut = torch.tensor([u // 2 for u in units], device="cuda", dtype=torch.int32)
        ua = torch.tensor([1 - u % 2 for u in units], device="cuda", dtype=torch.int32)
        coef = torch.tensor([(5.0 if u % 2 else -4.0) * (sn - s) for u in units], device="cuda", dtype=torch.float32)
        eu = eps.view(56, 4, k, k)[torch.tensor(units, device="cuda")].contiguous()
        ops.unit_accumulate(x, eu, la.weights64, la.origins, ut, coef, ua, part)
        num += part
    got, _ = ops.canvas_finalize(num, cnt, BF)
    assert rel(got, ref) < 5e-3          # differs only by the bf16 roundings the linearised form skips
