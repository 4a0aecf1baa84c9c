"""ORACLE / TEST INFRASTRUCTURE ONLY.  Run in the BUILD container (needs /root/reference, which the GPU
box does not have):

    python oracle/make_golden.py

1. imports the reference's OWN files (unmodified, except the three hard-coded 'cuda' device literals at
   pipeline_FaithDiff_tlc.py:228,543,1223 → 'cpu') on top of the diffusers shim,
2. asserts oracle/faithdiff_oracle.py reproduces them EXACTLY on seeded inputs (tile layout, Gaussian
   weights, blend, colour fix, UNet forward incl. FaithDiff stem, LQ encoder tiling, the whole __call__),
3. writes the small fixtures tests/golden/* that pin the oracle (and through it the CUDA path) on the GPU box.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)

from oracle import faithdiff_oracle as O  # noqa: E402


def load_ref_pipeline_module():
    src = open(os.path.join(REF, "FaithDiff/pipelines/pipeline_FaithDiff_tlc.py")).read()
    n0 = src.count("torch.device('cuda')") + src.count("device = 'cuda'")
    assert n0 == 3, n0
    src = src.replace("torch.device('cuda')", "torch.device('cpu')").replace("device = 'cuda'", "device = 'cpu'")
    import FaithDiff.pipelines  # noqa: F401  (package for the relative import)
    mod = types.ModuleType("FaithDiff.pipelines.pipeline_FaithDiff_tlc_cpu")
    mod.__package__ = "FaithDiff.pipelines"
    mod.__file__ = "pipeline_FaithDiff_tlc.py(cpu-patched)"
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    torch.set_num_threads(os.cpu_count())
    P = load_ref_pipeline_module()
    manifest = {}

    # ------------------------------------------------------------------ 1. tile layout + weights (bit-exact)
    cases = {"cfg3_4k": (270, 480, 128, 0.5), "cfg4_8k": (540, 960, 128, 0.5), "cfg1": (64, 64, 128, 0.5),
             "cfg5": (256, 256, 128, 0.5), "ragged": (150, 131, 128, 0.5), "ov25": (270, 480, 128, 0.25),
             "small_k": (37, 53, 16, 0.5), "one_row": (100, 300, 128, 0.5)}
    layout = {}
    for name, (h, w, k, ov) in cases.items():
        la = P.LocalAttention((k, k), ov)
        x = torch.arange(h * w, dtype=torch.float32).reshape(1, 1, h, w).repeat(1, 4, 1, 1)
        tiles = la.grids(x)
        lo = O.LocalAttention((k, k), ov)
        tiles_o = lo.grids(x)
        assert la.idxes == lo.idxes and torch.equal(tiles, tiles_o)
        assert torch.equal(la.tile_weights, lo.tile_weights), name
        layout[name] = dict(h=h, w=w, k=k, overlap=ov, idxes=[[d["i"], d["j"]] for d in la.idxes],
                            weights_sha=sha(la.tile_weights[0].numpy()), k1=tiles.shape[2], k2=tiles.shape[3])
        # blend
        g = torch.Generator().manual_seed(5)
        outs = torch.randn(tiles.shape, generator=g).to(torch.bfloat16)
        la.kernel_size = lo.kernel_size = (tiles.shape[2], tiles.shape[3]) if name in ("cfg1",) else la.kernel_size
        if min(h, w) >= k:
            r, ro = la.grids_inverse(outs.clone()), lo.grids_inverse(outs.clone())
            assert torch.equal(r, ro), name
            layout[name]["blend_sha_bf16_seed5"] = sha(r.numpy())
    w128 = P.LocalAttention((128, 128), 0.5)._gaussian_weights(128, 128)[0].numpy()
    assert sha(w128) == "f0fce21202f1baa7", sha(w128)                      # SURVEY §8a a5 known answer
    np.save(os.path.join(GOLD, "gaussian_weights_128.npy"), w128)
    np.save(os.path.join(GOLD, "gaussian_weights_64.npy"), P.LocalAttention((64, 64), 0.5)._gaussian_weights(64, 64)[0].numpy())
    manifest["layout"] = layout
    print("layout/weights/blend: reference == oracle (bit-exact)")

    # ------------------------------------------------------------------ 2. colour fix (real reference code)
    from utils import color_fix as RC
    from PIL import Image
    rng = np.random.default_rng(3)
    tgt = Image.fromarray(rng.integers(0, 256, (97, 131, 3), dtype=np.uint8))
    src = Image.fromarray(rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)).resize((131, 97), Image.BILINEAR)
    a_ref, w_ref = np.array(RC.adain_color_fix(tgt, src)), np.array(RC.wavelet_color_fix(tgt, src))
    assert np.array_equal(a_ref, np.array(O.adain_color_fix(tgt, src)))
    assert np.array_equal(w_ref, np.array(O.wavelet_color_fix(tgt, src)))
    np.savez_compressed(os.path.join(GOLD, "color_fix.npz"), target=np.array(tgt), source=np.array(src),
                        adain=a_ref, wavelet=w_ref)
    print("color_fix: reference == oracle (bit-exact)")

    # ------------------------------------------------------------------ 3. UNet forward (+ stem) & LQ encoder
    from FaithDiff.models.unet_2d_condition_vae_extension import UNet2DConditionModel as RefUNet
    cfg = dict(O.SDXL_UNET_CONFIG, block_out_channels=(320, 64, 128), attention_head_dim=(5, 1, 2),
               transformer_layers_per_block=(1, 1, 2), cross_attention_dim=256, addition_time_embed_dim=32,
               projection_class_embeddings_input_dim=64 + 192, sample_size=16)
    torch.manual_seed(1234)
    ref_unet = RefUNet(**cfg)
    ref_unet.load_additional_layers(dtype=None)
    O.randomize_zero_modules(ref_unet)
    ref_unet.eval()
    ora_unet = O.FaithDiffUNet(**cfg)
    ora_unet.load_additional_layers(channel=512)
    missing = ora_unet.load_state_dict(ref_unet.state_dict(), strict=True)
    ora_unet.eval()
    assert list(ref_unet.state_dict().keys()) == list(ora_unet.state_dict().keys())
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 4, 16, 16, generator=g)
    lq = torch.randn(2, 512, 16, 16, generator=g)
    ehs = torch.randn(2, 77, 256, generator=g)
    te = torch.randn(2, 64, generator=g)
    ids = torch.tensor([[128., 128., 0., 8., 128., 128.]] * 2)
    with torch.no_grad():
        for t in (torch.tensor(951.), torch.tensor(1.)):
            yr = ref_unet(x, t, encoder_hidden_states=ehs, input_embedding=lq,
                          added_cond_kwargs={"text_embeds": te, "time_ids": ids}, return_dict=False)[0]
            yo = ora_unet(x, t, encoder_hidden_states=ehs, input_embedding=lq,
                          added_cond_kwargs={"text_embeds": te, "time_ids": ids}, return_dict=False)[0]
            assert torch.equal(yr, yo)
        # odd size → forward_upsample_size path (…ext.py:526-529, 665-666)
        xo = torch.randn(2, 4, 18, 14, generator=g)
        lqo = torch.randn(2, 512, 18, 14, generator=g)
        yr = ref_unet(xo, torch.tensor(501.), encoder_hidden_states=ehs, input_embedding=lqo,
                      added_cond_kwargs={"text_embeds": te, "time_ids": ids}, return_dict=False)[0]
        yo = ora_unet(xo, torch.tensor(501.), encoder_hidden_states=ehs, input_embedding=lqo,
                      added_cond_kwargs={"text_embeds": te, "time_ids": ids}, return_dict=False)[0]
        assert torch.equal(yr, yo)
        # LQ encoder, tiled (…ext.py:164-196)
        img = torch.rand(1, 3, 160, 208, generator=g) * 2 - 1
        for enc in (ref_unet.denoise_encoder, ora_unet.denoise_encoder):
            enc.tile_sample_min_size, enc.tile_latent_min_size, enc.use_tiling = 64, 8, True
        er, eo = ref_unet.denoise_encoder(img), ora_unet.denoise_encoder(img)
        assert torch.equal(er, eo) and er.shape == (1, 512, 20, 26), er.shape
    print("UNet forward (+stem, odd sizes) and tiled LQ encoder: reference == oracle (bit-exact)")

    # ------------------------------------------------------------------ 4. whole __call__ (tiled, 3 steps)
    from diffusers import AutoencoderKL, DDPMScheduler, EulerDiscreteScheduler
    torch.manual_seed(4321)
    vae = AutoencoderKL(block_out_channels=(32, 64, 64, 64), sample_size=64).eval()
    vae.enable_tiling()
    for enc in (ref_unet.denoise_encoder, ora_unet.denoise_encoder):
        enc.tile_sample_min_size, enc.tile_latent_min_size, enc.use_tiling = 128, 16, True
    lr_img = O.synthetic_lq_image(208, 160, seed=0)
    pe, npe = torch.randn(1, 77, 256, generator=g), torch.randn(1, 77, 256, generator=g)
    pp, npp = torch.randn(1, 64, generator=g), torch.randn(1, 64, generator=g)

    pipe = P.FaithDiffStableDiffusionXLPipeline(
        vae=vae, denoise_encoder=ref_unet.denoise_encoder, text_encoder=None, text_encoder_2=None, tokenizer=None,
        tokenizer_2=None, unet=ref_unet, scheduler=EulerDiscreteScheduler(), DDPM_scheduler=DDPMScheduler(),
        add_watermarker=False)
    pipe.encode_prompt = lambda *a, **k: (pe, npe, pp, npp)      # prompt encoding is outside the hot path (§8f)
    rec_ref = {}
    orig_inverse = P.LocalAttention.grids_inverse

    def rec_inverse(self, outs):
        r = orig_inverse(self, outs)
        rec_ref.setdefault("step_latents", []).append(r.clone())
        return r
    P.LocalAttention.grids_inverse = rec_inverse
    torch.manual_seed(42)
    out_ref = pipe(lr_img=lr_img, prompt="x", negative_prompt="", num_inference_steps=3, guidance_scale=5.0,
                   generator=torch.Generator().manual_seed(42), start_point="lr", height=160, width=208, overlap=0.5,
                   target_size=(128, 128)).images[0]
    P.LocalAttention.grids_inverse = orig_inverse
    rec = {}
    torch.manual_seed(42)
    out_ora = O.pipeline_call(ora_unet, vae, EulerDiscreteScheduler(), DDPMScheduler(), lr_img, pe, npe, pp, npp,
                              160, 208, num_inference_steps=3, start_point="lr", overlap=0.5, guidance_scale=5.0,
                              generator=torch.Generator().manual_seed(42), target_size=(128, 128), record=rec)[0]
    assert len(rec_ref["step_latents"]) == 3
    for a, b in zip(rec_ref["step_latents"], rec["step_latents"]):
        assert torch.equal(a.to(b.dtype), b)
    assert np.array_equal(np.array(out_ref), np.array(out_ora))
    np.savez_compressed(os.path.join(GOLD, "pipeline_small.npz"), lr=np.array(lr_img), out=np.array(out_ref),
                        step_latents=np.stack([s.numpy() for s in rec["step_latents"]]),
                        pe=pe.numpy(), npe=npe.numpy(), pp=pp.numpy(), npp=npp.numpy())
    manifest["pipeline_small"] = dict(cfg={k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()},
                                      unet_seed=1234, vae_seed=4321, steps=3, height=160, width=208,
                                      out_sha=sha(np.array(out_ref)))
    print("whole __call__ (20 tiles? no: %d tiles, 3 steps): reference == oracle (bit-exact)" % len(
        O.LocalAttention.tile_origins(20, 26, 16, 16, 0.5)[0]))

    # denoising_end (pipeline…:1297-1311): the 20-step schedule cut after the timesteps >= 850 (3 steps) — assertion only
    rec_ref.clear()
    P.LocalAttention.grids_inverse = rec_inverse
    torch.manual_seed(42)
    out_ref = pipe(lr_img=lr_img, prompt="x", negative_prompt="", num_inference_steps=20, denoising_end=0.15, guidance_scale=5.0,
                   generator=torch.Generator().manual_seed(42), start_point="lr", height=160, width=208, overlap=0.5,
                   target_size=(128, 128)).images[0]
    P.LocalAttention.grids_inverse = orig_inverse
    rec = {}
    torch.manual_seed(42)
    out_ora = O.pipeline_call(ora_unet, vae, EulerDiscreteScheduler(), DDPMScheduler(), lr_img, pe, npe, pp, npp,
                              160, 208, num_inference_steps=20, start_point="lr", overlap=0.5, guidance_scale=5.0,
                              generator=torch.Generator().manual_seed(42), target_size=(128, 128), record=rec, denoising_end=0.15)[0]
    assert len(rec_ref["step_latents"]) == 3 and len(rec["step_latents"]) == 3
    for a, b in zip(rec_ref["step_latents"], rec["step_latents"]):
        assert torch.equal(a.to(b.dtype), b)
    assert np.array_equal(np.array(out_ref), np.array(out_ora))
    manifest["pipeline_small"]["denoising_end_0.15_of_20_steps_checked"] = True
    print("__call__ with denoising_end=0.15 of 20 steps (3 steps run): reference == oracle (bit-exact)")

    # ------------------------------------------------------------------ 5. scheduler known answers
    s = EulerDiscreteScheduler()
    s.set_timesteps(20)
    manifest["euler20"] = dict(timesteps=s.timesteps.tolist(), sigmas=[float(v) for v in s.sigmas],
                               init_noise_sigma=float(s.init_noise_sigma))
    assert abs(manifest["euler20"]["sigmas"][0] - 11.02834) < 1e-4 and s.timesteps[0] == 951
    json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(GOLD)))


if __name__ == "__main__":
    main()
