"""-m gpu: parity at BASELINE.json's FULL architecture (SDXL-size UNet + FaithDiff stem, 2.6 B parameters, 128x128 latent
tile, CFG batch 2; full-width LQ encoder and VAE with their d=512 attention) against the fp32 oracle evaluated by torch on the
GPU (test infrastructure: TF32 off), the benchmarked cfg3 layout (270x480 latent canvas, 28 tiles, start_point='lr') end to end,
plus size-independent properties of the tiled path.  ABSOLUTE bars (BASELINE.json north star): per-step latents rel-L2 <= 1e-2,
final image >= 40 dB; raw module outputs rel-L2 <= 2e-2 (UNet eps <= 1e-2).  All numbers -> profiles/r02_parity.json."""
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


@pytest.fixture(scope="module")
def full_vae():
    import faithdiff_dev_b200 as fd
    from faithdiff_dev_b200 import arch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = arch.random_state_dict(arch.vae_param_shapes(), seed=4321)
    vae = fd.AutoencoderKL(**arch.SDXL_VAE_CONFIG)
    vae.load_state_dict(sd)
    vae_o = O.AutoencoderKL(**arch.SDXL_VAE_CONFIG).cuda().eval()
    vae_o.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    return vae, vae_o


def _psnr_u8(a, b):
    a, b = np.asarray(a).astype(np.float64), np.asarray(b).astype(np.float64)
    return float(10 * np.log10(255.0 ** 2 / max(((a - b) ** 2).mean(), 1e-12)))


@pytest.mark.timeout(900)
def test_fullsize_unet_tile_step_matches_fp32_oracle(full, parity_record):
    fd, unet, ora = full
    rec = {}
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
        rec["t=%g" % t] = rel(out, ref)
        assert torch.isfinite(out.float()).all()
    # informational only (no bar depends on it): the same network evaluated by torch in bf16, the reference's own library path
    ob = copy.deepcopy(ora).to(BF)
    with torch.no_grad():
        rb = ob(x, torch.tensor(1.0), encoder_hidden_states=ehs, input_embedding=lq,
                added_cond_kwargs={"text_embeds": te, "time_ids": ids.to(BF)}, return_dict=False)[0]
    del ob
    rec["torch_bf16_t=1_for_comparison"] = rel(rb, ref)
    parity_record("fullsize_unet_eps_rel_l2_vs_fp32_oracle", **rec)
    assert max(rec["t=951"], rec["t=1"]) <= 1e-2, rec


@pytest.mark.timeout(900)
def test_fullsize_lq_encoder_and_vae_tiles_match_fp32_oracle(full, full_vae, parity_record):
    """Real channel widths (128-512) and the d=512 mid-block attention at its real lengths: LQ encoder on a 1024^2 tile
    (L = 16384; unet_2d_condition_vae_extension.py:96-106,192), VAE encode of a 512^2 tile and decode of 64^2 / 128^2 latent
    tiles (L = 4096 / 16384)."""
    fd, unet, ora = full
    vae, vae_o = full_vae
    g = torch.Generator(device="cuda").manual_seed(11)
    errs = {}
    img = (torch.rand(1, 3, 1024, 1024, generator=g, device="cuda") * 2 - 1).to(BF)
    with torch.no_grad():
        ref = ora.denoise_encoder(img.float())
    out = unet.denoise_encoder(img)
    assert out.shape == ref.shape == (1, 512, 128, 128)
    errs["lq_encoder_1024px_tile"] = rel(out, ref)
    del ref, out
    img2 = (torch.rand(2, 3, 512, 512, generator=g, device="cuda") * 2 - 1).to(BF)     # batch 2: the batched tile path
    with torch.no_grad():
        mo = vae_o.encode(img2.float()).latent_dist
    m = vae.encode(img2).latent_dist
    errs["vae_encode_mean_512px_tile"], errs["vae_encode_std_512px_tile"] = rel(m.mean, mo.mean), rel(m.std, mo.std)
    for n, name in ((64, "vae_decode_64_latent_tile"), (128, "vae_decode_128_latent_tile")):
        z = torch.randn(1, 4, n, n, generator=g, device="cuda").to(BF)
        with torch.no_grad():
            ref = vae_o.decode(z.float(), return_dict=False)[0]
        out = vae.decode(z, return_dict=False)[0]
        assert out.shape == ref.shape == (1, 3, 8 * n, 8 * n)
        errs[name] = rel(out, ref)
        u8 = lambda t: ((t.float() / 2 + 0.5).clamp(0, 1) * 255).round().cpu().numpy()
        errs[name + "_psnr_db"] = _psnr_u8(u8(out), u8(ref))
        del ref, out
    parity_record("fullsize_encoder_vae_vs_fp32_oracle", **errs)
    # feature / moment outputs: rel-L2 <= 2e-2; decoded images: >= 40 dB (the north star's image bar), raw rel-L2 <= 3e-2
    assert all(v <= (3e-2 if "decode" in k else 2e-2) for k, v in errs.items() if not k.endswith("_db")), errs
    assert all(v >= 40.0 for k, v in errs.items() if k.endswith("_db")), errs


@pytest.mark.timeout(1800)
def test_cfg3_layout_pipeline_call_matches_fp32_oracle(full, full_vae, parity_record, monkeypatch):
    """The benchmarked configuration itself (bench.py workload 4k = BASELINE cfg3): 1920x1080 synthetic LQ image, LANCZOS x2 to
    3840x2160, the real 270x480 latent canvas with its 28 tiles, full-size UNet / LQ encoder / VAE, tiled encoder + tiled VAE,
    start_point='lr', CFG 5, the benchmark's 20-step Euler schedule — cut after its first three steps with the reference's own
    `denoising_end` (pipeline…:1297-1311; the oracle is 56 fp32 UNet evaluations per step), or run to the end with
    FD_PARITY_FULL=1 (~7 min; scripts/gpu_validate.sh records that run in profiles/r02_parity.json) — against
    oracle.pipeline_call in fp32 on the GPU with the same injected initial / posterior noise."""
    import os
    dend = None if os.environ.get("FD_PARITY_FULL") == "1" else 0.15
    nsteps = 20 if dend is None else 3
    import diffusers._restated as R
    import faithdiff_dev_b200.vae as fvae
    import oracle.faithdiff_oracle as OM
    from PIL import Image
    fd, unet, ora = full
    vae, vae_o = full_vae
    H, W = 2160, 3840
    lr_img = O.synthetic_lq_image(1920, 1080, seed=0).resize((W, H), Image.LANCZOS)
    g = torch.Generator().manual_seed(7)
    pe, npe = torch.randn(1, 77, 2048, generator=g).to(BF), torch.randn(1, 77, 2048, generator=g).to(BF)
    pp, npp = torch.randn(1, 1280, generator=g).to(BF), torch.randn(1, 1280, generator=g).to(BF)
    lat = torch.randn(1, 4, H // 8, W // 8, generator=g).to(BF)
    zn = torch.randn(1, 4, H // 8, W // 8, generator=g).to(BF)
    monkeypatch.setattr(OM, "randn_tensor", lambda shape, generator=None, device=None, dtype=None: lat.to(device=device, dtype=dtype))
    monkeypatch.setattr(R, "randn_tensor", lambda shape, generator=None, device=None, dtype=None, layout=None: zn.to(device=device, dtype=dtype))
    monkeypatch.setattr(fvae.DiagonalGaussianDistribution, "sample", lambda self, generator=None: self.mean + self.std * zn.to(self.mean.device, self.mean.dtype))
    pipe = fd.FaithDiffStableDiffusionXLPipeline(vae, unet.denoise_encoder, None, None, None, None, unet, fd.EulerDiscreteScheduler(),
                                                 fd.DDPMScheduler())
    pipe.set_encoder_tile_settings()
    pipe.enable_vae_tiling()
    ora.denoise_encoder.tile_sample_min_size, ora.denoise_encoder.tile_overlap_factor = 1024, 0.25      # pipeline…:888-896
    ora.denoise_encoder.enable_tiling()
    vae_o.tile_overlap_factor = 0.25
    vae_o.enable_tiling()
    steps = []
    try:
        out = pipe(lr_img=lr_img, prompt_embeds=pe.cuda(), negative_prompt_embeds=npe.cuda(), pooled_prompt_embeds=pp.cuda(),
                   negative_pooled_prompt_embeds=npp.cuda(), num_inference_steps=20, denoising_end=dend, start_point="lr", height=H,
                   width=W, overlap=0.5, guidance_scale=5.0, target_size=(1024, 1024), latents=lat.cuda(),
                   callback_on_step_end=lambda p, i, t, kw: steps.append(kw["latents"].float().clone()))
        img = out.images[0]
        assert len(pipe.tlc_vae_latents.idxes) == 28
        assert [(d["i"], d["j"]) for d in pipe.tlc_vae_latents.idxes] == [(i, j) for i in (0, 64, 128, 142) for j in (0, 64, 128, 192, 256, 320, 352)]
        pipe._graphs = {}                       # release the tile-step graphs' memory before the fp32 oracle runs
        torch.cuda.empty_cache()
        rec = {}
        f = lambda t: t.float().cuda()
        img_o = O.pipeline_call(ora, vae_o, O.EulerDiscreteScheduler(), O.DDPMScheduler(), lr_img, f(pe), f(npe), f(pp), f(npp), H, W,
                                num_inference_steps=20, denoising_end=dend, start_point="lr", overlap=0.5, guidance_scale=5.0,
                                target_size=(1024, 1024), record=rec, device="cuda", dtype=torch.float32)[0]
    finally:
        ora.denoise_encoder.use_tiling = False
        vae_o.disable_tiling()
    errs = [rel(a, b) for a, b in zip(steps, rec["step_latents"])]
    psnr = _psnr_u8(img, img_o)
    parity_record("cfg3_layout_pipeline_call_start_point_lr_%dof20_steps" % nsteps, per_step_latent_rel_l2=errs, image_psnr_db=psnr,
                  oracle_eps_std_uncond_text_cfg_first_last_step=[rec["eps_std_first_tile"][0], rec["eps_std_first_tile"][-1]],
                  tiles=28, steps_run=nsteps, schedule_steps=20, canvas=[270, 480], tolerance={"latent_rel_l2": 1e-2, "psnr_db": 40.0})
    assert len(errs) == nsteps and max(errs) <= 1e-2, errs
    assert psnr >= 40.0, psnr


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
