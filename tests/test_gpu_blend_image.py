"""-m gpu: tile layout / Gaussian weights / blend / fused CFG+Euler tail / seam blend / image pre-post / colour fix
against the CPU oracle and the committed golden fixtures.  Integer & layout work and the blend are BIT-EXACT."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import faithdiff_oracle as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def fd():
    import faithdiff_dev_b200 as fd
    return fd


@pytest.mark.parametrize("h,w,k,ov,dtype", [(270, 480, 128, 0.5, BF), (540, 960, 128, 0.5, BF), (150, 131, 128, 0.5, torch.float16),
                                            (37, 53, 16, 0.5, torch.float32), (64, 64, 128, 0.5, BF), (270, 480, 128, 0.25, BF)])
def test_grids_and_blend_bit_exact(fd, h, w, k, ov, dtype):
    g = torch.Generator().manual_seed(11)
    canvas = torch.randn(1, 4, h, w, generator=g).to(dtype)
    lo = O.LocalAttention((k, k), ov)
    tiles_o = lo.grids(canvas)
    la = fd.LocalAttention((k, k), ov)
    tiles = la.grids(canvas.cuda())
    assert la.idxes == lo.idxes
    assert torch.equal(la.tile_weights.cpu(), lo.tile_weights)          # fp64 Gaussian weights: bit-exact
    assert torch.equal(tiles.cpu(), tiles_o)
    outs = torch.randn(tiles_o.shape, generator=g).to(dtype)
    lo.kernel_size = la.kernel_size = tuple(tiles_o.shape[2:])
    assert torch.equal(la.grids_inverse(outs.cuda()).cpu(), lo.grids_inverse(outs))


def test_golden_layout_and_weights(fd, golden_dir):
    from faithdiff_dev_b200.tiling import gaussian_weights, tile_origins
    man = json.load(open(os.path.join(golden_dir, "manifest.json")))
    for name, c in man["layout"].items():
        idx, k1, k2 = tile_origins(c["h"], c["w"], c["k"], c["k"], c["overlap"])
        assert [list(x) for x in idx] == c["idxes"], name
    assert np.array_equal(gaussian_weights(128, 128), np.load(os.path.join(golden_dir, "gaussian_weights_128.npy")))
    assert np.array_equal(gaussian_weights(64, 64), np.load(os.path.join(golden_dir, "gaussian_weights_64.npy")))


def test_fused_cfg_euler_blend_matches_oracle_scheduler(fd):
    """One full per-step tail: CFG combine + EulerDiscreteScheduler.step per tile + grids_inverse + cast + re-grid +
    scale_model_input, against the oracle's scheduler/blend code run op-by-op in bf16 ON THE GPU (torch's CUDA scalar
    semantics — fp32 host scalars, divide-by-scalar as reciprocal multiply — are what the reference executes) — bit-exact."""
    from faithdiff_dev_b200 import ops
    from faithdiff_dev_b200.scheduler import EulerDiscreteScheduler
    g = torch.Generator().manual_seed(12)
    h, w, k = 40, 56, 16
    sch_o = O.EulerDiscreteScheduler(); sch_o.set_timesteps(20)
    sch = EulerDiscreteScheduler(); sch.set_timesteps(20)
    assert torch.equal(sch.sigmas, sch_o.sigmas) and torch.equal(sch.timesteps.cpu(), sch_o.timesteps)
    lo = O.LocalAttention((k, k), 0.5)
    canvas = (torch.randn(1, 4, h, w, generator=g) * 5).to(BF).cuda()
    tiles = lo.grids(canvas)
    T = tiles.shape[0]
    eps = torch.randn(T, 2, 4, k, k, generator=g).to(BF).cuda()
    step = 3
    sch_o._step_index = step
    new = []
    import copy
    for t in range(T):
        s2 = copy.deepcopy(sch_o)
        eu, et = eps[t, 0:1], eps[t, 1:2]
        e = eu + 5.0 * (et - eu)
        new.append(s2.step(e, sch_o.timesteps[step], tiles[t:t + 1], return_dict=False)[0])
    blended = lo.grids_inverse(torch.cat(new, 0))
    lowp = blended.to(BF)
    nxt = lo.grids(lowp)
    s3 = copy.deepcopy(sch_o); s3._step_index = step + 1
    nxs = s3.scale_model_input(nxt, sch_o.timesteps[step + 1])
    la = fd.LocalAttention((k, k), 0.5)
    la.grids(canvas)
    sigma, sigma_next, _ = sch.step_scalars(step)
    next_div = sch.step_scalars(step + 1)[2]
    c32, cl, nt, ns = ops.cfg_euler_blend(eps, tiles, la.weights64, la.origins, h, w, 5.0, sigma, sigma_next, next_div)
    assert torch.equal(c32, blended)
    assert torch.equal(cl, lowp)
    assert torch.equal(nt, nxt)
    assert torch.equal(ns, nxs)
    # multi-GPU form: per-rank accumulate + finalize reproduces the same canvas up to fp32 summation order
    cnt = torch.zeros(1, 4, h, w, device="cuda")
    ops.tile_accumulate(torch.ones(T, 4, k, k, device="cuda"), None, la.weights64, la.origins, cnt)
    num = torch.zeros(1, 4, h, w, device="cuda")
    for r in range(2):
        sel = torch.arange(r, T, 2, device="cuda")
        part = torch.zeros_like(num)
        ops.tile_accumulate(tiles[sel].contiguous(), eps[sel].contiguous(), la.weights64, la.origins[sel].contiguous(),
                            part, 5.0, sigma, sigma_next)
        num += part
    c2, _ = ops.canvas_finalize(num, cnt, BF)
    assert (c2 - blended).abs().max() <= 1e-5 * blended.abs().max()
    # fp32 latent state (the pipeline's default in bf16): the same tail in fp32 arithmetic on the bf16 eps — CFG combine, Euler
    # update and blend without any cast to the model dtype: the fp32 canvas equals the blend of the unrounded per-tile updates,
    # next tiles are its exact re-grid, next scaled tiles its bf16 scaling
    x32 = tiles.float() + 0.001 * torch.randn(tiles.shape, generator=g).cuda()
    new32 = []
    for t in range(T):
        s2 = copy.deepcopy(sch_o)
        eu, et = eps[t, 0:1].float(), eps[t, 1:2].float()
        e = eu + 5.0 * (et - eu)
        sg = s2.sigmas[step]                                            # 0-dim host tensor, as in the scheduler
        pred0 = x32[t:t + 1] - sg * e                                   # EulerDiscreteScheduler.step, gamma = 0, before the cast
        new32.append(x32[t:t + 1] + ((x32[t:t + 1] - pred0) / sg) * (s2.sigmas[step + 1] - sg))
    blended32 = lo.grids_inverse(torch.cat(new32, 0))
    c32b, clb, ntb, nsb = ops.cfg_euler_blend(eps, x32.contiguous(), la.weights64, la.origins, h, w, 5.0, sigma, sigma_next, next_div)
    assert ntb.dtype == torch.float32 and nsb.dtype == BF and clb.dtype == BF
    assert (c32b - blended32).abs().max() <= 2e-6 * blended32.abs().max()          # fp32 op order
    assert torch.equal(ntb, lo.grids(c32b)) and torch.equal(clb, c32b.to(BF))
    assert torch.equal(nsb, (ntb * (1.0 / torch.tensor(next_div, dtype=torch.float32).cuda())).to(BF))


def test_scale_div_and_seam_blend(fd):
    from faithdiff_dev_b200 import ops
    g = torch.Generator().manual_seed(13)
    x = (torch.randn(3, 4, 16, 16, generator=g) * 10).to(BF)
    div = torch.tensor(11.0735, dtype=torch.float32)
    assert torch.equal(ops.scale_div(x.cuda(), float(div)), x.cuda() / div)      # torch CUDA semantics (reciprocal multiply)
    for dtype in (BF, torch.float32):
        a = torch.randn(1, 8, 24, 20, generator=g).to(dtype)
        b = torch.randn(1, 8, 17, 20, generator=g).to(dtype)
        bo = O.AutoencoderKL.blend_v(a.clone(), b.clone(), 6)
        bg = ops.seam_blend(a.cuda(), b.clone().cuda(), 6, True)
        assert torch.equal(bg.cpu(), bo)
        a2 = torch.randn(1, 8, 24, 20, generator=g).to(dtype)
        b2 = torch.randn(1, 8, 24, 9, generator=g).to(dtype)
        assert torch.equal(ops.seam_blend(a2.cuda(), b2.clone().cuda(), 12, False).cpu(), O.AutoencoderKL.blend_h(a2.clone(), b2.clone(), 12))


def test_image_pre_post(fd):
    from faithdiff_dev_b200 import ops
    img = O.synthetic_lq_image(72, 40, seed=4)
    u8 = torch.from_numpy(np.asarray(img)).cuda()
    nhwc, _ = ops.image_preprocess(u8)
    ref = O.VaeImageProcessor().preprocess([img], height=40, width=72).to(BF)
    assert torch.equal(nhwc.cpu().permute(0, 3, 1, 2), ref)
    g = torch.Generator().manual_seed(14)
    dec = (torch.randn(1, 3, 40, 72, generator=g) * 0.7).to(BF)
    pil = O.VaeImageProcessor().postprocess(dec, output_type="pil")[0]
    out = ops.image_postprocess(ops.nchw_to_nhwc(dec.cuda()), True)
    assert np.array_equal(out.cpu().numpy(), np.asarray(pil))


@pytest.mark.parametrize("mode", ["adain", "wavelet"])
def test_color_fix_golden(fd, golden_dir, mode):
    """Golden = the reference's real utils/color_fix.py (fp32 CPU).  u8 outputs may differ by one count where the fp32
    summation order moves a value across a truncation boundary; PSNR must stay far above the 40 dB bar."""
    from PIL import Image
    z = np.load(os.path.join(golden_dir, "color_fix.npz"))
    fn = fd.adain_color_fix if mode == "adain" else fd.wavelet_color_fix
    out = np.asarray(fn(Image.fromarray(z["target"]), Image.fromarray(z["source"]))).astype(np.int32)
    ref = z[mode].astype(np.int32)
    diff = np.abs(out - ref)
    assert diff.max() <= 1, diff.max()
    assert (diff > 0).mean() < 0.02
    mse = (diff.astype(np.float64) ** 2).mean()
    assert mse == 0 or 10 * np.log10(255 ** 2 / mse) > 60
    # and the same at a larger size against the oracle restatement
    big_t, big_s = O.synthetic_lq_image(333, 257, seed=5), O.synthetic_lq_image(333, 257, seed=6)
    o = np.asarray((O.adain_color_fix if mode == "adain" else O.wavelet_color_fix)(big_t, big_s)).astype(np.int32)
    gq = np.asarray(fn(big_t, big_s)).astype(np.int32)
    assert np.abs(o - gq).max() <= 1


@pytest.mark.parametrize("w,h,ow,oh,mode", [(37, 23, 74, 46, "RGB"), (64, 48, 128, 96, "RGB"), (50, 40, 200, 160, "RGB"),
                                            (100, 80, 50, 40, "RGB"), (33, 17, 40, 17, "RGB"), (31, 29, 31, 60, "RGB"),
                                            (120, 90, 37, 29, "CMYK"), (16, 16, 129, 67, "L"), (480, 270, 960, 540, "RGB"),
                                            (1920, 1080, 3840, 2160, "RGB")])
def test_lanczos_resize_bit_exact_vs_pil(fd, w, h, ow, oh, mode):
    """`image.resize((w*s, h*s), Image.LANCZOS)` of the caller (test_wo_llava.py:81) on the GPU: integer work, bit-exact."""
    from PIL import Image
    from faithdiff_dev_b200.image_process import resize_lanczos
    ch = {"L": 1, "RGB": 3, "CMYK": 4}[mode]
    a = np.random.default_rng(w + 7 * oh).integers(0, 256, size=(h, w, ch) if ch > 1 else (h, w), dtype=np.uint8)
    img = Image.fromarray(a, mode)
    ref = np.asarray(img.resize((ow, oh), Image.LANCZOS))
    got = resize_lanczos(img, (ow, oh))
    assert got.is_cuda and got.dtype == torch.uint8
    assert np.array_equal(got.cpu().numpy(), ref)
    # tensor inputs (host and device) take the same path
    assert torch.equal(resize_lanczos(torch.from_numpy(a).cuda(), (ow, oh)), got)
    with pytest.raises(ValueError):                  # PIL premultiplies alpha for RGBA: not this kernel's contract
        resize_lanczos(Image.new("RGBA", (8, 8)), (16, 16))


def test_pipeline_preprocess_resizes_like_pil(fd):
    """VaeImageProcessor.preprocess (pipeline…:1246) resizes a PIL input to (width, height) with LANCZOS before normalising;
    the pipeline does that resize on the GPU: same u8 image as PIL's, same normalised tensor as the oracle's arithmetic."""
    import types
    from PIL import Image
    a = np.random.default_rng(3).integers(0, 256, size=(40, 56, 3), dtype=np.uint8)
    img = Image.fromarray(a)
    fake = types.SimpleNamespace(device=torch.device("cuda"))
    nhwc, dev = fd.FaithDiffStableDiffusionXLPipeline._preprocess(fake, img, 80, 112)
    ref_u8 = np.asarray(img.resize((112, 80), Image.LANCZOS))
    assert np.array_equal(dev.cpu().numpy(), ref_u8) and fake._h2d_bytes == a.size
    ref = (torch.from_numpy(ref_u8.copy()).float() / 255.0) * 2.0 - 1.0
    assert torch.equal(nhwc[0].float().cpu(), ref.to(BF).float())
    # same size: no resize, the u8 image itself is uploaded
    nhwc2, dev2 = fd.FaithDiffStableDiffusionXLPipeline._preprocess(fake, img, 40, 56)
    assert np.array_equal(dev2.cpu().numpy(), a)
