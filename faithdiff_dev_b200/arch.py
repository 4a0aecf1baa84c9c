"""Checkpoint key layout of the hot path's models (SURVEY §8b "UNet weights"): (key → shape) specs generated from
the configs, in the reference's naming (diffusers 0.28 + FaithDiff prefixes, unet_2d_condition_vae_extension.py:444-478).
Used to validate incoming state dicts and to draw random-init weights of the named architecture for benchmarks."""
import math

import torch

SDXL_UNET_CONFIG = dict(
    sample_size=128, in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn", up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
    block_out_channels=(320, 640, 1280), layers_per_block=2, downsample_padding=1, act_fn="silu", norm_num_groups=32,
    norm_eps=1e-5, cross_attention_dim=2048, transformer_layers_per_block=(1, 2, 10), attention_head_dim=(5, 10, 20),
    use_linear_projection=True, addition_embed_type="text_time", addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=2816)
SDXL_VAE_CONFIG = dict(block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4, sample_size=512,
                       scaling_factor=0.13025)


def _lin(d, k, o, i, bias=True):
    d[k + ".weight"] = (o, i)
    if bias:
        d[k + ".bias"] = (o,)


def _conv(d, k, o, i, ks=3):
    d[k + ".weight"] = (o, i, ks, ks)
    d[k + ".bias"] = (o,)


def _norm(d, k, c):
    d[k + ".weight"] = (c,)
    d[k + ".bias"] = (c,)


def _resnet(d, k, cin, cout, temb):
    _norm(d, k + ".norm1", cin)
    _conv(d, k + ".conv1", cout, cin)
    if temb:
        _lin(d, k + ".time_emb_proj", cout, temb)
    _norm(d, k + ".norm2", cout)
    _conv(d, k + ".conv2", cout, cout)
    if cin != cout:
        _conv(d, k + ".conv_shortcut", cout, cin, 1)


def _transformer(d, k, c, layers, xdim):
    _norm(d, k + ".norm", c)
    _lin(d, k + ".proj_in", c, c)
    for i in range(layers):
        b = "%s.transformer_blocks.%d" % (k, i)
        _norm(d, b + ".norm1", c)
        for n in ("to_q", "to_k", "to_v"):
            _lin(d, b + ".attn1." + n, c, c, bias=False)
        _lin(d, b + ".attn1.to_out.0", c, c)
        _norm(d, b + ".norm2", c)
        _lin(d, b + ".attn2.to_q", c, c, bias=False)
        _lin(d, b + ".attn2.to_k", c, xdim, bias=False)
        _lin(d, b + ".attn2.to_v", c, xdim, bias=False)
        _lin(d, b + ".attn2.to_out.0", c, c)
        _norm(d, b + ".norm3", c)
        _lin(d, b + ".ff.net.0.proj", 8 * c, c)
        _lin(d, b + ".ff.net.2", c, 4 * c)
    _lin(d, k + ".proj_out", c, c)


def _tup(v, n):
    return (v,) * n if isinstance(v, int) else tuple(v)


def unet_param_shapes(cfg, with_faithdiff=True, lq_channels=512, encoder_channels=(128, 256, 512, 512)):
    d = {}
    ch = cfg["block_out_channels"]
    n = len(ch)
    tl, lpb = _tup(cfg["transformer_layers_per_block"], n), _tup(cfg["layers_per_block"], n)
    temb, xdim = ch[0] * 4, cfg["cross_attention_dim"]
    _conv(d, "conv_in", ch[0], cfg["in_channels"])
    _lin(d, "time_embedding.linear_1", temb, ch[0]); _lin(d, "time_embedding.linear_2", temb, temb)
    _lin(d, "add_embedding.linear_1", temb, cfg["projection_class_embeddings_input_dim"]); _lin(d, "add_embedding.linear_2", temb, temb)
    oc = ch[0]
    for i, t in enumerate(cfg["down_block_types"]):
        ic, oc = oc, ch[i]
        for j in range(lpb[i]):
            _resnet(d, "down_blocks.%d.resnets.%d" % (i, j), ic if j == 0 else oc, oc, temb)
            if t == "CrossAttnDownBlock2D":
                _transformer(d, "down_blocks.%d.attentions.%d" % (i, j), oc, tl[i], xdim)
        if i != n - 1:
            _conv(d, "down_blocks.%d.downsamplers.0.conv" % i, oc, oc)
    _resnet(d, "mid_block.resnets.0", ch[-1], ch[-1], temb)
    _transformer(d, "mid_block.attentions.0", ch[-1], tl[-1], xdim)
    _resnet(d, "mid_block.resnets.1", ch[-1], ch[-1], temb)
    rc, rtl, rl = ch[::-1], tl[::-1], lpb[::-1]
    oc = rc[0]
    for i, t in enumerate(cfg["up_block_types"]):
        pc, oc = oc, rc[i]
        inc = rc[min(i + 1, n - 1)]
        for j in range(rl[i] + 1):
            skip = inc if j == rl[i] else oc
            _resnet(d, "up_blocks.%d.resnets.%d" % (i, j), (pc if j == 0 else oc) + skip, oc, temb)
            if t == "CrossAttnUpBlock2D":
                _transformer(d, "up_blocks.%d.attentions.%d" % (i, j), oc, rtl[i], xdim)
        if i != n - 1:
            _conv(d, "up_blocks.%d.upsamplers.0.conv" % i, oc, oc)
    _norm(d, "conv_norm_out", ch[0])
    _conv(d, "conv_out", cfg["out_channels"], ch[0])
    if with_faithdiff:
        for k, v in encoder_param_shapes(encoder_channels).items():
            d["denoise_encoder." + k] = v
        e = 2 * ch[0]
        for i in range(2):
            b = "information_transformer_layes.%d" % i
            d[b + ".attn.in_proj_weight"] = (3 * e, e); d[b + ".attn.in_proj_bias"] = (3 * e,)
            _lin(d, b + ".attn.out_proj", e, e)
            _norm(d, b + ".ln_1", e)
            _lin(d, b + ".mlp.c_fc", 2 * e, e); _lin(d, b + ".mlp.c_proj", e, 2 * e)
            _norm(d, b + ".ln_2", e)
        _lin(d, "spatial_ch_projs", ch[0], e)
        _conv(d, "condition_embedding.conv_in", lq_channels, lq_channels)
        _norm(d, "condition_embedding.norm_in", lq_channels)
        _conv(d, "condition_embedding.conv_out", ch[0], lq_channels)
    return d


def _vae_attn(d, k, c):
    _norm(d, k + ".group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        _lin(d, k + "." + n, c, c)


def encoder_param_shapes(ch=(128, 256, 512, 512), layers=2, in_channels=3, out_conv=None):
    d = {}
    _conv(d, "conv_in", ch[0], in_channels)
    oc = ch[0]
    for i in range(len(ch)):
        ic, oc = oc, ch[i]
        for j in range(layers):
            _resnet(d, "down_blocks.%d.resnets.%d" % (i, j), ic if j == 0 else oc, oc, None)
        if i != len(ch) - 1:
            _conv(d, "down_blocks.%d.downsamplers.0.conv" % i, oc, oc)
    _resnet(d, "mid_block.resnets.0", ch[-1], ch[-1], None)
    _vae_attn(d, "mid_block.attentions.0", ch[-1])
    _resnet(d, "mid_block.resnets.1", ch[-1], ch[-1], None)
    if out_conv:
        _norm(d, "conv_norm_out", ch[-1])
        _conv(d, "conv_out", out_conv, ch[-1])
    return d


def vae_param_shapes(cfg=SDXL_VAE_CONFIG):
    ch, L, lc = cfg["block_out_channels"], cfg["layers_per_block"], cfg["latent_channels"]
    d = {"encoder." + k: v for k, v in encoder_param_shapes(ch, L, 3, 2 * lc).items()}
    dd = {}
    _conv(dd, "conv_in", ch[-1], lc)
    _resnet(dd, "mid_block.resnets.0", ch[-1], ch[-1], None)
    _vae_attn(dd, "mid_block.attentions.0", ch[-1])
    _resnet(dd, "mid_block.resnets.1", ch[-1], ch[-1], None)
    rc = ch[::-1]
    oc = rc[0]
    for i in range(len(rc)):
        pc, oc = oc, rc[i]
        for j in range(L + 1):
            _resnet(dd, "up_blocks.%d.resnets.%d" % (i, j), pc if j == 0 else oc, oc, None)
        if i != len(rc) - 1:
            _conv(dd, "up_blocks.%d.upsamplers.0.conv" % i, oc, oc)
    _norm(dd, "conv_norm_out", ch[0])
    _conv(dd, "conv_out", 3, ch[0])
    d.update({"decoder." + k: v for k, v in dd.items()})
    _conv(d, "quant_conv", 2 * lc, 2 * lc, 1)
    _conv(d, "post_quant_conv", lc, lc, 1)
    return d


def random_state_dict(shapes, seed=1234, device="cuda", dtype=torch.bfloat16):
    """Random-init weights of the named architecture (torch default-init scale: U(±1/sqrt(fan_in)); norm affine (1, 0))."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        is_norm = any(t in k for t in (".norm", "norm_", "ln_1", "ln_2", "group_norm")) and len(shp) == 1
        if is_norm:
            sd[k] = torch.ones(shp, device=device, dtype=dtype) if k.endswith("weight") else torch.zeros(shp, device=device, dtype=dtype)
            continue
        if len(shp) > 1:
            fan_in = math.prod(shp[1:])
            bound = 1.0 / math.sqrt(fan_in)
        else:
            bound = 0.02
        sd[k] = ((torch.rand(shp, generator=g, device=device, dtype=torch.float32) * 2 - 1) * bound).to(dtype)
    return sd
