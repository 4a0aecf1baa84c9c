"""-m "not gpu": pins the CPU oracle against the committed golden fixtures (tests/golden/, produced by
oracle/make_golden.py from the reference's OWN files) and against known answers."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import faithdiff_oracle as O


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def manifest(golden_dir):
    return json.load(open(os.path.join(golden_dir, "manifest.json")))


def test_tile_layout_and_weights_golden(manifest, golden_dir):
    for name, c in manifest["layout"].items():
        idx, k1, k2 = O.LocalAttention.tile_origins(c["h"], c["w"], c["k"], c["k"], c["overlap"])
        assert [[d["i"], d["j"]] for d in idx] == c["idxes"], name
        assert (k1, k2) == (c["k1"], c["k2"])
        assert sha(O.LocalAttention.gaussian_weights_np(k2, k1)) == c["weights_sha"], name
    w = O.LocalAttention.gaussian_weights_np(128, 128)
    assert np.array_equal(w, np.load(os.path.join(golden_dir, "gaussian_weights_128.npy")))
    assert sha(w) == "f0fce21202f1baa7" and w.max() == 15.903356381505787 and w[0, 0] == 2.6850345886753826e-10
    assert len(manifest["layout"]["cfg3_4k"]["idxes"]) == 28 and len(manifest["layout"]["cfg4_8k"]["idxes"]) == 112
    assert manifest["layout"]["cfg3_4k"]["idxes"][-1] == [142, 352]


def test_blend_golden(manifest):
    for name, c in manifest["layout"].items():
        if "blend_sha_bf16_seed5" not in c:
            continue
        lo = O.LocalAttention((c["k"], c["k"]), c["overlap"])
        x = torch.arange(c["h"] * c["w"], dtype=torch.float32).reshape(1, 1, c["h"], c["w"]).repeat(1, 4, 1, 1)
        tiles = lo.grids(x)
        outs = torch.randn(tiles.shape, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16)
        assert sha(lo.grids_inverse(outs).numpy()) == c["blend_sha_bf16_seed5"], name


def test_color_fix_golden(golden_dir):
    from PIL import Image
    z = np.load(os.path.join(golden_dir, "color_fix.npz"))
    t, s = Image.fromarray(z["target"]), Image.fromarray(z["source"])
    assert np.array_equal(np.asarray(O.adain_color_fix(t, s)), z["adain"])
    assert np.array_equal(np.asarray(O.wavelet_color_fix(t, s)), z["wavelet"])


def test_scheduler_known_answers(manifest):
    s = O.EulerDiscreteScheduler()
    s.set_timesteps(20)
    assert s.timesteps.tolist() == manifest["euler20"]["timesteps"] == [float(951 - 50 * i) for i in range(20)]
    assert [float(v) for v in s.sigmas] == manifest["euler20"]["sigmas"]
    known = [11.02834, 8.39069, 6.5064, 5.13443, 4.1167, 3.34776, 2.7562, 2.29285, 1.92344, 1.62369, 1.37618, 1.16822,
             0.99041, 0.83565, 0.6984, 0.57405, 0.45833, 0.34619, 0.22815, 0.04131, 0.0]
    assert np.allclose([float(v) for v in s.sigmas], known, atol=6e-6)
    assert abs(float(s.init_noise_sigma) - 11.073581) < 1e-5
    s.set_timesteps(1)
    assert s.timesteps.tolist() == [1.0] and abs(float(s.sigmas[0]) - 0.04131) < 1e-5
    d = O.DDPMScheduler()
    assert abs(float(d.alphas_cumprod[999]) - 0.0046601) < 1e-6
    assert float(d.alphas_cumprod.to(torch.float16)[999]) == pytest.approx(0.00466156, abs=1e-8)


def test_parameter_count_known_answers():
    """Exact public SDXL totals pin every channel width / bias flag / norm placement of the restated architecture."""
    with torch.device("meta"):
        u = O.FaithDiffUNet(**O.SDXL_UNET_CONFIG)
        u.load_additional_layers(channel=512)
        v = O.AutoencoderKL()
    extra = ("denoise_encoder", "information_transformer_layes", "condition_embedding", "spatial_ch_projs")
    core = sum(p.numel() for n, p in u.named_parameters() if n.split(".")[0] not in extra)
    assert core == 2567463684
    assert sum(p.numel() for p in u.parameters()) == 2612197892
    assert sum(p.numel() for p in u.denoise_encoder.parameters()) == 34125696
    assert sum(p.numel() for p in v.parameters()) == 83653863
    keys = set(u.state_dict().keys())
    for k in ("down_blocks.1.attentions.0.transformer_blocks.1.attn2.to_k.weight", "mid_block.attentions.0.proj_in.weight",
              "up_blocks.0.attentions.2.transformer_blocks.9.ff.net.0.proj.bias", "up_blocks.1.upsamplers.0.conv.weight",
              "add_embedding.linear_1.weight", "time_embedding.linear_2.bias", "down_blocks.0.downsamplers.0.conv.bias",
              "denoise_encoder.mid_block.attentions.0.to_q.weight", "information_transformer_layes.1.attn.in_proj_weight",
              "condition_embedding.norm_in.weight", "spatial_ch_projs.weight", "up_blocks.2.resnets.2.conv_shortcut.weight"):
        assert k in keys, k
    assert u.state_dict()["add_embedding.linear_1.weight"].shape == (1280, 2816)


@pytest.mark.timeout(600)
def test_pipeline_small_golden(manifest, golden_dir):
    """Re-runs the oracle __call__ restatement with the seeds recorded in the manifest; must reproduce the image the
    REFERENCE's own pipeline file produced in oracle/make_golden.py, bit for bit."""
    z = np.load(os.path.join(golden_dir, "pipeline_small.npz"))
    m = manifest["pipeline_small"]
    cfg = {k: (tuple(v) if isinstance(v, list) else v) for k, v in m["cfg"].items()}
    unet = O.build_unet(cfg, seed=m["unet_seed"], lq_channels=512)
    torch.manual_seed(m["vae_seed"])
    vae = O.AutoencoderKL(block_out_channels=(32, 64, 64, 64), sample_size=64).eval()
    vae.enable_tiling()
    enc = unet.denoise_encoder
    enc.tile_sample_min_size, enc.tile_latent_min_size, enc.use_tiling = 128, 16, True
    from PIL import Image
    rec = {}
    torch.manual_seed(42)
    out = O.pipeline_call(unet, vae, O.EulerDiscreteScheduler(), O.DDPMScheduler(), Image.fromarray(z["lr"]),
                          torch.from_numpy(z["pe"]), torch.from_numpy(z["npe"]), torch.from_numpy(z["pp"]), torch.from_numpy(z["npp"]),
                          m["height"], m["width"], num_inference_steps=m["steps"], start_point="lr", overlap=0.5, guidance_scale=5.0,
                          generator=torch.Generator().manual_seed(42), target_size=(128, 128), record=rec)[0]
    assert np.array_equal(np.stack([s.numpy() for s in rec["step_latents"]]), z["step_latents"])
    assert np.array_equal(np.asarray(out), z["out"])


def test_check_image_size_and_mha_identity():
    from PIL import Image
    img = Image.fromarray(np.random.default_rng(0).integers(0, 256, (37, 50, 3), dtype=np.uint8))
    p, w0, h0, w1, h1 = O.check_image_size(img)
    assert (w0, h0, w1, h1) == (50, 37, 56, 40) and p.size == (56, 40)
    a = np.asarray(p)
    assert np.array_equal(a[:37, :50], np.asarray(img)) and np.array_equal(a[39, :50], a[36, :50]) and np.array_equal(a[:, 55], a[:, 49])
    # with identical CFG halves the batch-axis MHA is the identity on V
    torch.manual_seed(0)
    blk = O.ResidualAttentionBlock(64, 8).eval()
    x = torch.randn(1, 30, 64)
    y2 = blk(torch.cat([x, x], 0))
    y1 = blk(x)
    assert torch.allclose(y2[0], y1[0], atol=1e-6) and torch.allclose(y2[1], y1[0], atol=1e-6)


def test_bsrnet_oracle_reproduces_reference_golden(golden_dir):
    """oracle/bsrnet_oracle.py against the outputs of the reference's own RRDBNet (bsrnet_arch.py run by
    oracle/make_golden_bsrnet.py): bit-exact on this torch build, forward and tiled deg_remove; parameter count known answer."""
    import numpy as np
    from oracle import bsrnet_oracle as B
    g = np.load(os.path.join(golden_dir, "bsrnet.npz"))
    assert int(g["n_params_full"]) == 16697987 == sum(int(np.prod(s)) for s in B.param_shapes().values())
    cfg = dict(in_nc=3, out_nc=3, nf=64, nb=2, gc=32, sf=4)
    sd = B.random_state_dict(seed=11, **cfg)
    with torch.no_grad():
        for case in ("fwd_a", "fwd_ragged", "fwd_batch2"):
            y = B.forward(sd, torch.from_numpy(g[case + "_x"]))
            assert torch.allclose(y, torch.from_numpy(g[case + "_y"]), rtol=0, atol=2e-6), case
        for case in ("deg_a", "deg_ragged", "deg_one_tile"):
            ts, tp = (int(v) for v in g[case + "_tile"])
            y = B.deg_remove(sd, torch.from_numpy(g[case + "_x"]), tile_size=ts, tile_pad=tp)
            assert torch.allclose(y, torch.from_numpy(g[case + "_y"]), rtol=0, atol=2e-6), case
