"""ORACLE / TEST INFRASTRUCTURE ONLY — CPU (or any-device) torch restatement of the reference's
tiled latent-denoising hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this; the product package never does.

What it restates (reference file:line, all under /root/reference):
  LocalAttention            FaithDiff/pipelines/pipeline_FaithDiff_tlc.py:101-252
  Encoder (LQ encoder)      FaithDiff/models/unet_2d_condition_vae_extension.py:39-196
  ControlNetConditioningEmbedding / QuickGELU / ResidualAttentionBlock   :199-260
  UNet2DConditionModel.forward (FaithDiff stem + SDXL trunk)             :480-697
  pipeline __call__ hot loop                                             pipeline_FaithDiff_tlc.py:1216-1459
  colour fix                utils/color_fix.py:14-118
  check_image_size          utils/image_process.py:7-21
The diffusers 0.28.0 blocks underneath come from oracle/shim/diffusers/_restated.py.

Pinning: the reference has no tests / golden vectors → "parity unpinned" by the reference
itself.  oracle/make_golden.py (run in the build container, where /root/reference exists) executes the
reference's OWN files on top of the shim and asserts this restatement reproduces them exactly on the
same seeded inputs, then writes tests/golden/*.  The [D] layer (diffusers semantics) stays unpinned
except for parameter-count / key-name / scheduler known answers.
"""
import copy
import math
import os
import sys
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)

from diffusers._restated import (  # noqa: E402
    SDXL_UNET_CONFIG, AttrDict, AutoencoderKL, DDPMScheduler, EulerDiscreteScheduler, UNet2DConditionModelBase,
    UNetMidBlock2D, VaeImageProcessor, get_down_block, randn_tensor)


# ============================================================================ tiling (pipeline…:101-252)
class LocalAttention:
    def __init__(self, kernel_size=None, overlap=0.5):
        self.kernel_size = kernel_size
        self.overlap = overlap

    @staticmethod
    def tile_origins(h, w, k1, k2, overlap):
        """pipeline…:172-204 — origins in row-major order; last row/col clamped to dim-k."""
        k1, k2 = min(k1, h), min(k2, w)
        num_row = (h - 1) // k1 + 1
        num_col = (w - 1) // k2 + 1
        step_j = k2 if num_col == 1 else math.ceil(k2 * overlap)
        step_i = k1 if num_row == 1 else math.ceil(k1 * overlap)
        idxes = []
        i, last_i = 0, False
        while i < h and not last_i:
            j = 0
            if i + k1 >= h:
                i, last_i = h - k1, True
            last_j = False
            while j < w and not last_j:
                if j + k2 >= w:
                    j, last_j = w - k2, True
                idxes.append({"i": i, "j": j})
                j += step_j
            i += step_i
        return idxes, k1, k2

    @staticmethod
    def gaussian_weights_np(tile_width, tile_height):
        """pipeline…:207-228 — float64, numpy exp, x midpoint (W-1)/2, y midpoint H/2."""
        from numpy import exp, pi, sqrt
        var = 0.01
        mid = (tile_width - 1) / 2
        xp = [exp(-(x - mid) * (x - mid) / (tile_width * tile_width) / (2 * var)) / sqrt(2 * pi * var)
              for x in range(tile_width)]
        mid = tile_height / 2
        yp = [exp(-(y - mid) * (y - mid) / (tile_height * tile_height) / (2 * var)) / sqrt(2 * pi * var)
              for y in range(tile_height)]
        return np.outer(yp, xp)

    def grids(self, x):
        b, c, h, w = x.shape
        self.original_size = (b, c, h, w)
        assert b == 1
        idxes, k1, k2 = self.tile_origins(h, w, self.kernel_size[0], self.kernel_size[1], self.overlap)
        self.tile_weights = torch.tile(torch.tensor(self.gaussian_weights_np(k2, k1), device=x.device), (4, 1, 1))
        self.idxes = idxes
        return torch.cat([x[:, :, d["i"]:d["i"] + k1, d["j"]:d["j"] + k2] for d in idxes], dim=0)

    def grids_inverse(self, outs):
        """pipeline…:230-252 — fp32 canvas, fp64 products, row-major accumulation order."""
        preds = torch.zeros(self.original_size).to(outs.device)
        b, c, h, w = self.original_size
        count_mt = torch.zeros((b, 4, h, w)).to(outs.device)
        k1, k2 = self.kernel_size
        for cnt, d in enumerate(self.idxes):
            i, j = d["i"], d["j"]
            preds[0, :, i:i + k1, j:j + k2] += outs[cnt, :, :, :] * self.tile_weights
            count_mt[0, :, i:i + k1, j:j + k2] += self.tile_weights
        return preds / count_mt


# ============================================================================ LQ encoder (unet_…ext.py:39-196)
class Encoder(nn.Module):
    def __init__(self, in_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2, norm_num_groups=32):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], 3, 1, 1)
        self.down_blocks = nn.ModuleList()
        self.tile_sample_min_size = 1024
        self.tile_latent_min_size = int(self.tile_sample_min_size / 8)
        self.tile_overlap_factor = 0.25
        self.use_tiling = False
        oc = block_out_channels[0]
        for i in range(len(block_out_channels)):
            ic, oc = oc, block_out_channels[i]
            self.down_blocks.append(get_down_block(
                "DownEncoderBlock2D", num_layers=layers_per_block, in_channels=ic, out_channels=oc,
                add_downsample=i != len(block_out_channels) - 1, resnet_eps=1e-6, downsample_padding=0,
                resnet_groups=norm_num_groups, attention_head_dim=oc, temb_channels=None))
        self.mid_block = UNetMidBlock2D(in_channels=block_out_channels[-1], resnet_eps=1e-6, output_scale_factor=1,
                                        attention_head_dim=block_out_channels[-1], resnet_groups=norm_num_groups,
                                        temb_channels=None, add_attention=True)

    def enable_tiling(self):
        self.use_tiling = True

    def encode(self, sample):
        sample = self.conv_in(sample)
        for b in self.down_blocks:
            sample = b(sample)
        return self.mid_block(sample)

    blend_v = staticmethod(AutoencoderKL.blend_v)
    blend_h = staticmethod(AutoencoderKL.blend_h)

    def tiled_encode(self, x):
        overlap_size = int(self.tile_sample_min_size * (1 - self.tile_overlap_factor))
        blend_extent = int(self.tile_latent_min_size * self.tile_overlap_factor)
        row_limit = self.tile_latent_min_size - blend_extent
        rows = []
        for i in range(0, x.shape[2], overlap_size):
            row = []
            for j in range(0, x.shape[3], overlap_size):
                row.append(self.encode(x[:, :, i:i + self.tile_sample_min_size, j:j + self.tile_sample_min_size]))
            rows.append(row)
        result_rows = []
        for i, row in enumerate(rows):
            rr = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = self.blend_v(rows[i - 1][j], tile, blend_extent)
                if j > 0:
                    tile = self.blend_h(row[j - 1], tile, blend_extent)
                rr.append(tile[:, :, :row_limit, :row_limit])
            result_rows.append(torch.cat(rr, dim=3))
        return torch.cat(result_rows, dim=2)

    def forward(self, sample):
        if self.use_tiling and (sample.shape[-1] > self.tile_latent_min_size
                                or sample.shape[-2] > self.tile_latent_min_size):
            return self.tiled_encode(sample)
        return self.encode(sample)


# ============================================================================ FaithDiff stem (unet_…ext.py:199-260)
class ControlNetConditioningEmbedding(nn.Module):
    def __init__(self, conditioning_embedding_channels, conditioning_channels=4):
        super().__init__()
        self.conv_in = nn.Conv2d(conditioning_channels, conditioning_channels, 3, padding=1)
        self.norm_in = nn.GroupNorm(num_channels=conditioning_channels, num_groups=32, eps=1e-6)
        self.conv_out = nn.Conv2d(conditioning_channels, conditioning_embedding_channels, 3, padding=1)
        for p in self.conv_out.parameters():
            nn.init.zeros_(p)

    def forward(self, c):
        return self.conv_out(F.silu(self.conv_in(self.norm_in(c))))


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        return super().forward(x).type(x.dtype)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)   # batch_first=False: sequence axis = dim 0 (= batch!)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 2)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 2, d_model))]))
        self.ln_2 = LayerNorm(d_model)

    def forward(self, x):
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False, attn_mask=None)[0]
        return x + self.mlp(self.ln_2(x))


class FaithDiffUNet(UNet2DConditionModelBase):
    """unet_…ext.py:269-697 restated. State-dict keys equal the reference's (…ext.py:444-478)."""

    def __init__(self, **cfg):
        super().__init__(**cfg)
        self.denoise_encoder = None
        self.information_transformer_layes = None
        self.condition_embedding = None
        self.spatial_ch_projs = None

    def load_additional_layers(self, dtype=None, channel=512, weight_path=None, trans_channel=None, proj_channel=None,
                               encoder_kwargs=None):
        c0 = self.config.block_out_channels[0]
        proj_channel = proj_channel or c0          # reference: 320
        trans_channel = trans_channel or 2 * c0    # reference: 640 (= cat of two 320-ch maps)
        if self.denoise_encoder is None:
            self.denoise_encoder = Encoder(**(encoder_kwargs or {}))
        if self.information_transformer_layes is None:
            self.information_transformer_layes = nn.Sequential(
                *[ResidualAttentionBlock(trans_channel, 8) for _ in range(2)])
            self.spatial_ch_projs = nn.Linear(trans_channel, proj_channel)
            for p in self.spatial_ch_projs.parameters():
                nn.init.zeros_(p)
        if self.condition_embedding is None:
            self.condition_embedding = ControlNetConditioningEmbedding(proj_channel, channel)
        if dtype is not None:
            self.to(dtype=dtype)

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None, input_embedding=None,
                add_sample=True, return_dict=True, use_condition_embedding=True, **unused):
        up_factor = 2 ** self.num_upsamplers
        forward_upsample_size = any(d % up_factor != 0 for d in sample.shape[-2:])
        upsample_size = None
        # 1. time (…ext.py:543-564)
        t_emb = self.get_time_embed(sample=sample, timestep=timestep)
        emb = self.time_embedding(t_emb, None)
        emb = emb + self.get_aug_embed(emb=emb, encoder_hidden_states=encoder_hidden_states,
                                       added_cond_kwargs=added_cond_kwargs)
        # 2. stem (…ext.py:571-579)
        sample = self.conv_in(sample)
        if input_embedding is not None and self.condition_embedding is not None \
                and self.information_transformer_layes is not None:
            if use_condition_embedding:
                input_embedding = self.condition_embedding(input_embedding)
            b, c, h, w = input_embedding.shape
            feat = torch.cat([sample, input_embedding], dim=1).view(b, 2 * c, h * w).transpose(1, 2)
            feat = self.information_transformer_layes(feat)
            feat_alpha = self.spatial_ch_projs(feat).transpose(1, 2).view(b, c, h, w)
            sample = sample + feat_alpha if add_sample else feat_alpha
        # 3. down (…ext.py:611-630)
        res = (sample,)
        for blk in self.down_blocks:
            if getattr(blk, "has_cross_attention", False):
                sample, r = blk(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states)
            else:
                sample, r = blk(hidden_states=sample, temb=emb)
            res += r
        # 4. mid (…ext.py:642-653)
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states)
        # 5. up (…ext.py:661-684)
        for i, blk in enumerate(self.up_blocks):
            n = len(blk.resnets)
            r, res = res[-n:], res[:-n]
            if i != len(self.up_blocks) - 1 and forward_upsample_size:
                upsample_size = res[-1].shape[2:]
            if getattr(blk, "has_cross_attention", False):
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=r,
                             encoder_hidden_states=encoder_hidden_states, upsample_size=upsample_size)
            else:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=r, upsample_size=upsample_size)
        # 6. post (…ext.py:687-690)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        if not return_dict:
            return (sample,)
        return AttrDict(sample=sample)


# ============================================================================ synthetic weights
TINY_UNET_CONFIG = dict(SDXL_UNET_CONFIG, block_out_channels=(64, 128, 256), attention_head_dim=(1, 2, 4),
                        transformer_layers_per_block=(1, 1, 2), cross_attention_dim=256, sample_size=16,
                        addition_time_embed_dim=32, projection_class_embeddings_input_dim=64 + 6 * 32)
TINY_VAE_CONFIG = dict(block_out_channels=(32, 64, 64, 64), sample_size=64)
TINY_ENCODER_CONFIG = dict(block_out_channels=(32, 64, 64, 64))


def randomize_zero_modules(unet, seed=99, std=0.02):
    """zero_module'd layers (…ext.py:209,399) are re-drawn N(0, std²) so parity tests see the FaithDiff fusion."""
    g = torch.Generator().manual_seed(seed)
    for m in (unet.condition_embedding.conv_out, unet.spatial_ch_projs):
        for p in m.parameters():
            p.data.copy_(torch.randn(p.shape, generator=g) * std)


def build_unet(cfg=None, seed=1234, lq_channels=512, encoder_kwargs=None):
    torch.manual_seed(seed)
    unet = FaithDiffUNet(**(cfg or SDXL_UNET_CONFIG))
    unet.load_additional_layers(channel=lq_channels, encoder_kwargs=encoder_kwargs)
    randomize_zero_modules(unet)
    return unet.eval()


# ============================================================================ pipeline hot path (pipeline…:1216-1459)
@torch.no_grad()
def pipeline_call(unet, vae, scheduler, ddpm_scheduler, lr_img, prompt_embeds, negative_prompt_embeds,
                  pooled_prompt_embeds, negative_pooled_prompt_embeds, height, width, num_inference_steps=20,
                  start_point="lr", overlap=0.5, guidance_scale=5.0, generator=None, target_size=(1024, 1024),
                  original_size=None, add_sample=True, output_type="pil", device="cpu", dtype=torch.float32,
                  record=None, denoising_end=None):
    """Restates FaithDiffStableDiffusionXLPipeline.__call__ from the image preprocess (:1246) to postprocess
    (:1451); prompt encoding (:1231-1243, SURVEY §8f rank 1 'next') is replaced by injected embeddings.
    `record`, if a dict, receives per-step blended latents (the per-step parity quantity)."""
    device = torch.device(device)
    original_size = original_size or (height, width)
    target_size = target_size or (height, width)
    tlc_lat = LocalAttention((target_size[0] // 8, target_size[1] // 8), overlap)
    tlc_img = LocalAttention((target_size[0] // 8, target_size[1] // 8), overlap)
    proc = VaeImageProcessor(vae_scale_factor=8)
    lr = proc.preprocess([lr_img], height=height, width=width).to(device, dtype=dtype)
    scheduler.set_timesteps(num_inference_steps, device=device)
    timesteps = scheduler.timesteps
    image_latents = unet.denoise_encoder(lr)                               # :1250 → :970
    if image_latents.dtype != vae.dtype:
        image_latents = image_latents.to(vae.dtype)
    image_latents = tlc_img.grids(image_latents)                           # :1257
    shape = (1, vae.config.latent_channels, height // 8, width // 8)
    latents = randn_tensor(shape, generator=generator, device=device, dtype=dtype) * scheduler.init_noise_sigma
    if start_point == "lr":                                                # :1271-1276
        z0 = vae.encode(lr).latent_dist.sample() * vae.config.scaling_factor
        t0 = torch.randint(999, 1000, (latents.shape[0],), device=latents.device).long()
        latents = ddpm_scheduler.add_noise(z0[0:1], latents, t0)
    if record is not None:
        record["init_latents"] = latents.clone()
    latents = tlc_lat.grids(latents)                                       # :1278
    status = [copy.deepcopy(scheduler.__dict__)] * image_latents.shape[0]  # :1282
    pe = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0).to(device)
    te = torch.cat([negative_pooled_prompt_embeds, pooled_prompt_embeds], dim=0).to(device)
    T = latents.shape[0]
    if denoising_end is not None and isinstance(denoising_end, float) and 0 < denoising_end < 1:   # :1297-1311
        n_train = scheduler.config.num_train_timesteps
        cutoff = int(round(n_train - denoising_end * n_train))
        timesteps = timesteps[:len(list(filter(lambda ts: ts >= cutoff, timesteps)))]
    for i, t in enumerate(timesteps):                                      # :1324
        if i >= 1:
            latents = tlc_lat.grids(latents).to(dtype=latents.dtype)
        out = []
        for s in range(T):                                                 # :1330
            scheduler.__dict__.update(status[s])
            sub = latents[s].unsqueeze(0)
            img_sub = torch.cat([image_latents[s].unsqueeze(0)] * 2)
            x_in = scheduler.scale_model_input(torch.cat([sub] * 2), t)
            ids = torch.cat([torch.tensor([original_size]),
                             torch.tensor([[tlc_lat.idxes[s]["i"], tlc_lat.idxes[s]["j"]]]),
                             torch.tensor([target_size])], dim=1).to(img_sub.device, dtype=img_sub.dtype)
            ids = ids.repeat(2, 1).to(dtype=img_sub.dtype)
            eps = unet(x_in, t, encoder_hidden_states=pe, input_embedding=img_sub, add_sample=add_sample,
                       added_cond_kwargs={"text_embeds": te, "time_ids": ids}, return_dict=False)[0]
            eu, et = eps.chunk(2)
            eps = eu + guidance_scale * (et - eu)                          # :1364-1365
            if record is not None and s == 0:
                record.setdefault("eps_std_first_tile", []).append((float(eu.float().std()), float(et.float().std()), float(eps.float().std())))
            sub = scheduler.step(eps, t, sub, generator=generator, return_dict=False)[0]
            status[s] = copy.deepcopy(scheduler.__dict__)
            out.append(sub)
        latents = tlc_lat.grids_inverse(torch.cat(out, dim=0)).to(sub.dtype)   # :1409
        if record is not None:
            record.setdefault("step_latents", []).append(latents.clone())
    if output_type == "latent":
        return latents
    image = vae.decode(latents / vae.config.scaling_factor, return_dict=False)[0]   # :1436-1438
    if record is not None:
        record["decoded"] = image.clone()
    return proc.postprocess(image, output_type=output_type)


# ============================================================================ colour fix (utils/color_fix.py)
def calc_mean_std(feat, eps=1e-5):
    b, c = feat.shape[:2]
    var = feat.reshape(b, c, -1).var(dim=2) + eps          # unbiased (color_fix.py:54)
    return feat.reshape(b, c, -1).mean(dim=2).reshape(b, c, 1, 1), var.sqrt().reshape(b, c, 1, 1)


def adaptive_instance_normalization(content, style):
    sm, ss = calc_mean_std(style)
    cm, cs = calc_mean_std(content)
    return (content - cm.expand(content.shape)) / cs.expand(content.shape) * ss.expand(content.shape) + sm.expand(content.shape)


def wavelet_blur(image, radius):
    k = torch.tensor([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]],
                     dtype=image.dtype, device=image.device)[None, None].repeat(3, 1, 1, 1)
    image = F.pad(image, (radius,) * 4, mode="replicate")
    return F.conv2d(image, k, groups=3, dilation=radius)


def wavelet_decomposition(image, levels=5):
    high = torch.zeros_like(image)
    for i in range(levels):
        low = wavelet_blur(image, 2 ** i)
        high += image - low
        image = low
    return high, low


def wavelet_reconstruction(content, style):
    return wavelet_decomposition(content)[0] + wavelet_decomposition(style)[1]


def _to_tensor(pil):
    a = np.asarray(pil, dtype=np.uint8)
    return torch.from_numpy(a.transpose(2, 0, 1).copy()).float().div(255).unsqueeze(0)   # torchvision ToTensor


def _to_pil(t):
    from PIL import Image
    a = t.squeeze(0).clamp_(0.0, 1.0).mul(255).byte().permute(1, 2, 0).numpy()           # torchvision ToPILImage
    return Image.fromarray(a)


def adain_color_fix(target, source):
    return _to_pil(adaptive_instance_normalization(_to_tensor(target), _to_tensor(source)))


def wavelet_color_fix(target, source):
    return _to_pil(wavelet_reconstruction(_to_tensor(target), _to_tensor(source)))


# ============================================================================ image pre (utils/image_process.py:7-21)
def check_image_size(x, padder_size=8):
    from PIL import Image
    width, height = x.size
    ph = (padder_size - height % padder_size) % padder_size
    pw = (padder_size - width % padder_size) % padder_size
    a = np.pad(np.array(x), ((0, ph), (0, pw), (0, 0)), mode="edge")    # == cv2.BORDER_REPLICATE
    return Image.fromarray(a), width, height, width + pw, height + ph


def synthetic_lq_image(w, h, seed=0):
    """low-pass uniform noise (8× box down/up), RGB u8."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, size=((h + 7) // 8, (w + 7) // 8, 3), dtype=np.uint8)
    return Image.fromarray(small).resize((w, h), Image.BILINEAR)
