"""ORACLE / TEST INFRASTRUCTURE ONLY — never imported by the product path.

CPU/torch restatement of the slice of `diffusers==0.28.0` that the reference's hot
path reaches (requirements.txt:2 pins it; the source is NOT in /root/reference and
not installed in this image, so this file restates the published algorithms from
SURVEY.md Appendix B).  PARITY UNPINNED for the [D] parts: the reference ships no
golden vectors; the cross-checks that exist (exact SDXL parameter counts, real
checkpoint key names, scheduler known answers) are asserted in tests/.

Reference call sites this serves:
  FaithDiff/models/unet_2d_condition_vae_extension.py:23,27   (UNet base, get_down_block, UNetMidBlock2D)
  FaithDiff/pipelines/pipeline_FaithDiff_tlc.py:13-53         (AutoencoderKL, schedulers, image processor)
Everything is plain torch.nn with the default parameter layouts so real SDXL
checkpoints would load by key name.
"""
import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- config
class AttrDict(dict):
    """FrozenDict stand-in; attributes are assignable (pipeline_FaithDiff_tlc.py:895 relies on it)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


# ----------------------------------------------------------------------------- embeddings (B5)
def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0,
                           scale=1.0, max_period=10000):
    half_dim = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half_dim, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half_dim - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half_dim:], emb[:, :half_dim]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


# ----------------------------------------------------------------------------- resnet / resampling (B2, B4)
class Downsample2D(nn.Module):
    def __init__(self, channels, out_channels=None, padding=1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, padding=1)

    def forward(self, x, output_size=None):
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels=512, groups=32, eps=1e-6, output_scale_factor=1.0):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.output_scale_factor = output_scale_factor
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb=None):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / self.output_scale_factor


# ----------------------------------------------------------------------------- attention (B3, B6)
class Attention(nn.Module):
    """diffusers Attention + AttnProcessor2_0 (SDPA) semantics."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False, out_bias=True,
                 norm_num_groups=None, eps=1e-5, residual_connection=False, rescale_output_factor=1.0):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.residual_connection = residual_connection
        self.rescale_output_factor = rescale_output_factor
        self.group_norm = nn.GroupNorm(norm_num_groups, query_dim, eps=eps, affine=True) if norm_num_groups else None
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(0.0)])
        self.processor = AttnProcessor2_0()

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **kw):
        residual = hidden_states
        nd = hidden_states.ndim
        if nd == 4:
            b, c, hh, ww = hidden_states.shape
            hidden_states = hidden_states.view(b, c, hh * ww).transpose(1, 2)
        if self.group_norm is not None:
            hidden_states = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        q = self.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = self.to_k(ctx)
        v = self.to_v(ctx)
        B = q.shape[0]
        d = q.shape[-1] // self.heads
        q = q.view(B, -1, self.heads, d).transpose(1, 2)
        k = k.view(B, -1, self.heads, d).transpose(1, 2)
        v = v.view(B, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, -1, self.heads * d).to(q.dtype)
        o = self.to_out[1](self.to_out[0](o))
        if nd == 4:
            o = o.transpose(-1, -2).reshape(b, c, hh, ww)
        if self.residual_connection:
            o = o + residual
        return o / self.rescale_output_factor


class AttnProcessor2_0:
    pass


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states=None, **kw):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states=encoder_hidden_states) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, num_layers, cross_attention_dim, norm_num_groups=32):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, x, encoder_hidden_states=None, return_dict=False, **kw):
        b, c, h, w = x.shape
        res = x
        x = self.norm(x)
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        x = self.proj_in(x)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states=encoder_hidden_states)
        x = self.proj_out(x)
        x = x.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
        return (x + res,)


# ----------------------------------------------------------------------------- UNet blocks (B1)
class DownBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, in_c, out_c, temb, n, eps, groups, add_down, down_pad):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_c if i == 0 else out_c, out_c, temb, groups, eps) for i in range(n)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_c, out_c, down_pad)]) if add_down else None

    def forward(self, hidden_states, temb=None, **kw):
        outs = ()
        for r in self.resnets:
            hidden_states = r(hidden_states, temb)
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class CrossAttnDownBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, in_c, out_c, temb, n, tl, eps, groups, heads, xdim, add_down, down_pad):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_c if i == 0 else out_c, out_c, temb, groups, eps) for i in range(n)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, out_c // heads, out_c, tl, xdim, groups) for _ in range(n)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_c, out_c, down_pad)]) if add_down else None

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, **kw):
        outs = ()
        for r, a in zip(self.resnets, self.attentions):
            hidden_states = r(hidden_states, temb)
            hidden_states = a(hidden_states, encoder_hidden_states=encoder_hidden_states)[0]
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_c, out_c, n, eps, groups, add_down, down_pad):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_c if i == 0 else out_c, out_c, None, groups, eps) for i in range(n)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_c, out_c, down_pad)]) if add_down else None

    def forward(self, x, *a, **kw):
        for r in self.resnets:
            x = r(x, None)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                x = d(x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_c, out_c, n, eps, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_c if i == 0 else out_c, out_c, None, groups, eps) for i in range(n)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_c, out_c)]) if add_up else None

    def forward(self, x, temb=None):
        for r in self.resnets:
            x = r(x, None)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                x = u(x)
        return x


class UNetMidBlock2D(nn.Module):
    """VAE-style mid block: resnet, single-head attention (d = channels), resnet."""

    def __init__(self, in_channels, temb_channels=None, resnet_eps=1e-6, resnet_act_fn="silu", output_scale_factor=1.0,
                 resnet_time_scale_shift="default", attention_head_dim=1, resnet_groups=32, add_attention=True, **kw):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, in_channels, temb_channels, resnet_groups, resnet_eps,
                                                    output_scale_factor) for _ in range(2)])
        if attention_head_dim is None:
            attention_head_dim = in_channels
        self.attentions = nn.ModuleList([
            Attention(in_channels, None, in_channels // attention_head_dim, attention_head_dim, bias=True,
                      norm_num_groups=resnet_groups, eps=resnet_eps, residual_connection=True,
                      rescale_output_factor=output_scale_factor) if add_attention else None])

    def forward(self, x, temb=None):
        x = self.resnets[0](x, temb)
        for a, r in zip(self.attentions, self.resnets[1:]):
            if a is not None:
                x = a(x, temb=temb)
            x = r(x, temb)
        return x


class UNetMidBlock2DCrossAttn(nn.Module):
    has_cross_attention = True

    def __init__(self, c, temb, tl, eps, groups, heads, xdim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, c // heads, c, tl, xdim, groups)])

    def forward(self, x, temb=None, encoder_hidden_states=None, **kw):
        x = self.resnets[0](x, temb)
        for a, r in zip(self.attentions, self.resnets[1:]):
            x = a(x, encoder_hidden_states=encoder_hidden_states)[0]
            x = r(x, temb)
        return x


class UpBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, in_c, prev_c, out_c, temb, n, eps, groups, add_up):
        super().__init__()
        rs = []
        for i in range(n):
            skip = in_c if i == n - 1 else out_c
            rin = prev_c if i == 0 else out_c
            rs.append(ResnetBlock2D(rin + skip, out_c, temb, groups, eps))
        self.resnets = nn.ModuleList(rs)
        self.upsamplers = nn.ModuleList([Upsample2D(out_c, out_c)]) if add_up else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None, **kw):
        for r in self.resnets:
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = r(torch.cat([hidden_states, res], dim=1), temb)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class CrossAttnUpBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, in_c, prev_c, out_c, temb, n, tl, eps, groups, heads, xdim, add_up):
        super().__init__()
        rs, at = [], []
        for i in range(n):
            skip = in_c if i == n - 1 else out_c
            rin = prev_c if i == 0 else out_c
            rs.append(ResnetBlock2D(rin + skip, out_c, temb, groups, eps))
            at.append(Transformer2DModel(heads, out_c // heads, out_c, tl, xdim, groups))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList(at)
        self.upsamplers = nn.ModuleList([Upsample2D(out_c, out_c)]) if add_up else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                upsample_size=None, **kw):
        for r, a in zip(self.resnets, self.attentions):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = r(torch.cat([hidden_states, res], dim=1), temb)
            hidden_states = a(hidden_states, encoder_hidden_states=encoder_hidden_states)[0]
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


def get_down_block(down_block_type, num_layers, in_channels, out_channels, temb_channels=None, add_downsample=True,
                   resnet_eps=1e-5, resnet_act_fn="silu", transformer_layers_per_block=1, num_attention_heads=None,
                   resnet_groups=32, cross_attention_dim=None, downsample_padding=1, attention_head_dim=None, **kw):
    """unet_2d_condition_vae_extension.py:81-93 calls this for DownEncoderBlock2D."""
    if down_block_type == "DownEncoderBlock2D":
        return DownEncoderBlock2D(in_channels, out_channels, num_layers, resnet_eps, resnet_groups, add_downsample,
                                  downsample_padding)
    if down_block_type == "DownBlock2D":
        return DownBlock2D(in_channels, out_channels, temb_channels, num_layers, resnet_eps, resnet_groups,
                           add_downsample, downsample_padding)
    if down_block_type == "CrossAttnDownBlock2D":
        return CrossAttnDownBlock2D(in_channels, out_channels, temb_channels, num_layers, transformer_layers_per_block,
                                    resnet_eps, resnet_groups, num_attention_heads, cross_attention_dim, add_downsample,
                                    downsample_padding)
    raise ValueError(down_block_type)


# ----------------------------------------------------------------------------- UNet base
SDXL_UNET_CONFIG = dict(
    sample_size=128, in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn",
    up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
    block_out_channels=(320, 640, 1280), layers_per_block=2, downsample_padding=1, act_fn="silu",
    norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=2048, transformer_layers_per_block=(1, 2, 10),
    attention_head_dim=(5, 10, 20), use_linear_projection=True, addition_embed_type="text_time",
    addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816,
)


class UNet2DConditionModelBase(nn.Module):
    """The subset of diffusers.UNet2DConditionModel the reference subclass uses
    (unet_2d_condition_vae_extension.py:269-380, :543-690)."""

    def __init__(self, sample_size=None, in_channels=4, out_channels=4, center_input_sample=False,
                 flip_sin_to_cos=True, freq_shift=0,
                 down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
                 mid_block_type="UNetMidBlock2DCrossAttn",
                 up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3,
                 only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 downsample_padding=1, mid_block_scale_factor=1, dropout=0.0, act_fn="silu", norm_num_groups=32,
                 norm_eps=1e-5, cross_attention_dim=1280, transformer_layers_per_block=1,
                 reverse_transformer_layers_per_block=None, encoder_hid_dim=None, encoder_hid_dim_type=None,
                 attention_head_dim=8, num_attention_heads=None, dual_cross_attention=False,
                 use_linear_projection=False, class_embed_type=None, addition_embed_type=None,
                 addition_time_embed_dim=None, num_class_embeds=None, upcast_attention=False,
                 resnet_time_scale_shift="default", resnet_skip_time_act=False, resnet_out_scale_factor=1.0,
                 time_embedding_type="positional", time_embedding_dim=None, time_embedding_act_fn=None,
                 timestep_post_act=None, time_cond_proj_dim=None, conv_in_kernel=3, conv_out_kernel=3,
                 projection_class_embeddings_input_dim=None, attention_type="default", class_embeddings_concat=False,
                 mid_block_only_cross_attention=None, cross_attention_norm=None, addition_embed_type_num_heads=64):
        super().__init__()
        if not hasattr(self, "_internal_dict"):
            cfg = dict(locals())
            cfg.pop("self"); cfg.pop("__class__", None)
            self._internal_dict = AttrDict(cfg)
        assert use_linear_projection and addition_embed_type == "text_time", "oracle restates the SDXL-style config only"
        n = len(block_out_channels)
        heads = num_attention_heads or attention_head_dim
        heads = (heads,) * n if isinstance(heads, int) else tuple(heads)
        tl = (transformer_layers_per_block,) * n if isinstance(transformer_layers_per_block, int) else tuple(transformer_layers_per_block)
        lpb = (layers_per_block,) * n if isinstance(layers_per_block, int) else tuple(layers_per_block)
        temb = block_out_channels[0] * 4
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], conv_in_kernel, padding=(conv_in_kernel - 1) // 2)
        self.time_proj = Timesteps(block_out_channels[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], temb)
        self.add_time_proj = Timesteps(addition_time_embed_dim, flip_sin_to_cos, freq_shift)
        self.add_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, temb)
        self.time_embed_act = None
        self.position_net = None
        self.down_blocks = nn.ModuleList()
        out_c = block_out_channels[0]
        for i, t in enumerate(down_block_types):
            in_c, out_c = out_c, block_out_channels[i]
            self.down_blocks.append(get_down_block(
                t, lpb[i], in_c, out_c, temb, add_downsample=(i != n - 1), resnet_eps=norm_eps,
                transformer_layers_per_block=tl[i], num_attention_heads=heads[i], resnet_groups=norm_num_groups,
                cross_attention_dim=cross_attention_dim, downsample_padding=downsample_padding))
        self.mid_block = UNetMidBlock2DCrossAttn(block_out_channels[-1], temb, tl[-1], norm_eps, norm_num_groups,
                                                 heads[-1], cross_attention_dim)
        self.num_upsamplers = 0
        self.up_blocks = nn.ModuleList()
        rc, rh, rtl, rl = block_out_channels[::-1], heads[::-1], tl[::-1], lpb[::-1]
        out_c = rc[0]
        for i, t in enumerate(up_block_types):
            prev_c, out_c = out_c, rc[i]
            in_c = rc[min(i + 1, n - 1)]
            add_up = i != n - 1
            if add_up:
                self.num_upsamplers += 1
            if t == "CrossAttnUpBlock2D":
                blk = CrossAttnUpBlock2D(in_c, prev_c, out_c, temb, rl[i] + 1, rtl[i], norm_eps, norm_num_groups, rh[i],
                                         cross_attention_dim, add_up)
            elif t == "UpBlock2D":
                blk = UpBlock2D(in_c, prev_c, out_c, temb, rl[i] + 1, norm_eps, norm_num_groups, add_up)
            else:
                raise ValueError(t)
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, block_out_channels[0], eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, conv_out_kernel, padding=(conv_out_kernel - 1) // 2)

    @property
    def config(self):
        return self._internal_dict

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def get_time_embed(self, sample, timestep):
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            dt = torch.float32 if isinstance(timestep, float) else torch.int64
            timesteps = torch.tensor([timesteps], dtype=dt, device=sample.device)
        elif len(timesteps.shape) == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        return self.time_proj(timesteps).to(dtype=sample.dtype)

    def get_class_embed(self, sample, class_labels):
        return None

    def get_aug_embed(self, emb, encoder_hidden_states, added_cond_kwargs):
        text_embeds = added_cond_kwargs.get("text_embeds")
        time_ids = added_cond_kwargs.get("time_ids")
        time_embeds = self.add_time_proj(time_ids.flatten())
        time_embeds = time_embeds.reshape((text_embeds.shape[0], -1))
        add_embeds = torch.concat([text_embeds, time_embeds], dim=-1).to(emb.dtype)
        return self.add_embedding(add_embeds)

    def process_encoder_hidden_states(self, encoder_hidden_states, added_cond_kwargs):
        return encoder_hidden_states


# ----------------------------------------------------------------------------- VAE (B6)
class VAEEncoder(nn.Module):
    def __init__(self, in_channels=3, out_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 norm_num_groups=32, double_z=True):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        oc = block_out_channels[0]
        for i in range(len(block_out_channels)):
            ic, oc = oc, block_out_channels[i]
            self.down_blocks.append(DownEncoderBlock2D(ic, oc, layers_per_block, 1e-6, norm_num_groups,
                                                       i != len(block_out_channels) - 1, 0))
        self.mid_block = UNetMidBlock2D(block_out_channels[-1], None, 1e-6, attention_head_dim=block_out_channels[-1],
                                        resnet_groups=norm_num_groups)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, block_out_channels[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[-1], 2 * out_channels if double_z else out_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class VAEDecoder(nn.Module):
    def __init__(self, in_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 norm_num_groups=32):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[-1], 3, padding=1)
        self.mid_block = UNetMidBlock2D(block_out_channels[-1], None, 1e-6, attention_head_dim=block_out_channels[-1],
                                        resnet_groups=norm_num_groups)
        self.up_blocks = nn.ModuleList()
        rc = block_out_channels[::-1]
        oc = rc[0]
        for i in range(len(rc)):
            pc, oc = oc, rc[i]
            self.up_blocks.append(UpDecoderBlock2D(pc, oc, layers_per_block + 1, 1e-6, norm_num_groups, i != len(rc) - 1))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, block_out_channels[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, 3, padding=1)

    def forward(self, z, latent_embeds=None):
        x = self.conv_in(z)
        x = self.mid_block(x, latent_embeds)
        for b in self.up_blocks:
            x = b(x, latent_embeds)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        noise = randn_tensor(self.mean.shape, generator=generator, device=self.parameters.device,
                             dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class _EncOut:
    def __init__(self, d):
        self.latent_dist = d


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=4, norm_num_groups=32, sample_size=512, scaling_factor=0.13025, force_upcast=False,
                 latents_mean=None, latents_std=None):
        super().__init__()
        cfg = dict(locals()); cfg.pop("self"); cfg.pop("__class__", None)
        self._internal_dict = AttrDict(cfg)
        self.encoder = VAEEncoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.decoder = VAEDecoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.use_tiling = False
        self.tile_sample_min_size = sample_size
        self.tile_latent_min_size = int(sample_size / (2 ** (len(block_out_channels) - 1)))
        self.tile_overlap_factor = 0.25

    @property
    def config(self):
        return self._internal_dict

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def enable_tiling(self, use_tiling=True):
        self.use_tiling = use_tiling

    def disable_tiling(self):
        self.use_tiling = False

    @staticmethod
    def blend_v(a, b, blend_extent):
        blend_extent = min(a.shape[2], b.shape[2], blend_extent)
        for y in range(blend_extent):
            b[:, :, y, :] = a[:, :, -blend_extent + y, :] * (1 - y / blend_extent) + b[:, :, y, :] * (y / blend_extent)
        return b

    @staticmethod
    def blend_h(a, b, blend_extent):
        blend_extent = min(a.shape[3], b.shape[3], blend_extent)
        for x in range(blend_extent):
            b[:, :, :, x] = a[:, :, :, -blend_extent + x] * (1 - x / blend_extent) + b[:, :, :, x] * (x / blend_extent)
        return b

    def _tiled(self, x, fn, tile_in, stride, blend_extent, row_limit):
        rows = []
        for i in range(0, x.shape[2], stride):
            row = []
            for j in range(0, x.shape[3], stride):
                row.append(fn(x[:, :, i:i + tile_in, j:j + tile_in]))
            rows.append(row)
        out_rows = []
        for i, row in enumerate(rows):
            out = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = self.blend_v(rows[i - 1][j], tile, blend_extent)
                if j > 0:
                    tile = self.blend_h(row[j - 1], tile, blend_extent)
                out.append(tile[:, :, :row_limit, :row_limit])
            out_rows.append(torch.cat(out, dim=3))
        return torch.cat(out_rows, dim=2)

    def tiled_encode(self, x):
        stride = int(self.tile_sample_min_size * (1 - self.tile_overlap_factor))
        blend_extent = int(self.tile_latent_min_size * self.tile_overlap_factor)
        row_limit = self.tile_latent_min_size - blend_extent
        return self._tiled(x, lambda t: self.quant_conv(self.encoder(t)), self.tile_sample_min_size, stride,
                           blend_extent, row_limit)

    def tiled_decode(self, z):
        stride = int(self.tile_latent_min_size * (1 - self.tile_overlap_factor))
        blend_extent = int(self.tile_sample_min_size * self.tile_overlap_factor)
        row_limit = self.tile_sample_min_size - blend_extent
        return self._tiled(z, lambda t: self.decoder(self.post_quant_conv(t)), self.tile_latent_min_size, stride,
                           blend_extent, row_limit)

    def encode(self, x, return_dict=True):
        if self.use_tiling and (x.shape[-1] > self.tile_sample_min_size or x.shape[-2] > self.tile_sample_min_size):
            moments = self.tiled_encode(x)
        else:
            moments = self.quant_conv(self.encoder(x))
        return _EncOut(DiagonalGaussianDistribution(moments))

    def decode(self, z, return_dict=True, generator=None):
        if self.use_tiling and (z.shape[-1] > self.tile_latent_min_size or z.shape[-2] > self.tile_latent_min_size):
            dec = self.tiled_decode(z)
        else:
            dec = self.decoder(self.post_quant_conv(z))
        if not return_dict:
            return (dec,)
        return AttrDict(sample=dec)


# ----------------------------------------------------------------------------- schedulers (B7)
def _scaled_linear_alphas_cumprod(beta_start, beta_end, n):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class EulerDiscreteScheduler:
    """SDXL default scheduler config: scaled_linear 0.00085..0.012, leading spacing, steps_offset 1, epsilon."""
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        self.config = AttrDict(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                               steps_offset=steps_offset, timestep_spacing="leading", prediction_type="epsilon")
        self.alphas_cumprod = _scaled_linear_alphas_cumprod(beta_start, beta_end, num_train_timesteps)
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).flip(0)
        ts = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy()
        self.timesteps = torch.from_numpy(ts).to(dtype=torch.float32)
        self.sigmas = torch.cat([sigmas, torch.zeros(1)]).to("cpu")
        self.num_inference_steps = None
        self._step_index = None
        self._begin_index = None
        self.is_scale_input_called = False

    @property
    def init_noise_sigma(self):
        return (self.sigmas.max() ** 2 + 1) ** 0.5

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        n = self.config.num_train_timesteps
        step_ratio = n // num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.float32)
        timesteps += self.config.steps_offset
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32).to("cpu")
        self.timesteps = torch.from_numpy(timesteps.astype(np.float32)).to(device=device)
        self._step_index = None
        self._begin_index = None

    def _init_step_index(self, timestep):
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(self.timesteps.device)
        idx = (self.timesteps == timestep).nonzero()
        pos = 1 if len(idx) > 1 else 0
        self._step_index = idx[pos].item()

    def scale_model_input(self, sample, timestep):
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self._step_index]
        sample = sample / ((sigma ** 2 + 1) ** 0.5)
        self.is_scale_input_called = True
        return sample

    def step(self, model_output, timestep, sample, generator=None, return_dict=True, **kw):
        if self._step_index is None:
            self._init_step_index(timestep)
        sample = sample.to(torch.float32)
        sigma = self.sigmas[self._step_index]
        # gamma == 0 (s_churn = 0); the scheduler still draws an unused noise tensor from `generator`
        randn_tensor(model_output.shape, dtype=model_output.dtype, device=model_output.device, generator=generator)
        sigma_hat = sigma * (0.0 + 1)
        pred_original_sample = sample - sigma_hat * model_output
        derivative = (sample - pred_original_sample) / sigma_hat
        dt = self.sigmas[self._step_index + 1] - sigma_hat
        prev_sample = (sample + derivative * dt).to(model_output.dtype)
        self._step_index += 1
        if not return_dict:
            return (prev_sample,)
        return AttrDict(prev_sample=prev_sample, pred_original_sample=pred_original_sample)


class DDPMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
        self.config = AttrDict(num_train_timesteps=num_train_timesteps)
        self.alphas_cumprod = _scaled_linear_alphas_cumprod(beta_start, beta_end, num_train_timesteps)

    def add_noise(self, original_samples, noise, timesteps):
        self.alphas_cumprod = self.alphas_cumprod.to(device=original_samples.device)
        ac = self.alphas_cumprod.to(dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        a = (ac[timesteps] ** 0.5).flatten()
        while a.ndim < original_samples.ndim:
            a = a.unsqueeze(-1)
        b = ((1 - ac[timesteps]) ** 0.5).flatten()
        while b.ndim < original_samples.ndim:
            b = b.unsqueeze(-1)
        return a * original_samples + b * noise


# ----------------------------------------------------------------------------- misc (B8, B9)
def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    device = torch.device(device) if device is not None else torch.device("cpu")
    rand_device = device
    if generator is not None:
        gdev = generator.device.type
        if gdev != device.type and gdev == "cpu":
            rand_device = torch.device("cpu")
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)


class VaeImageProcessor:
    def __init__(self, vae_scale_factor=8, **kw):
        self.config = AttrDict(vae_scale_factor=vae_scale_factor, resample="lanczos", do_normalize=True)

    def preprocess(self, image, height=None, width=None):
        import PIL.Image
        if not isinstance(image, list):
            image = [image]
        f = self.config.vae_scale_factor
        if height is None:
            height = image[0].height
        if width is None:
            width = image[0].width
        width, height = (x - x % f for x in (width, height))
        image = [i.resize((width, height), resample=PIL.Image.LANCZOS) for i in image]
        arr = np.stack([np.array(i).astype(np.float32) / 255.0 for i in image], axis=0)
        t = torch.from_numpy(arr.transpose(0, 3, 1, 2))
        return 2.0 * t - 1.0

    def postprocess(self, image, output_type="pil", do_denormalize=None):
        import PIL.Image
        if output_type == "latent" or output_type == "pt" and False:
            return image
        image = (image / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return image
        arr = image.cpu().permute(0, 2, 3, 1).float().numpy()
        if output_type == "np":
            return arr
        arr = (arr * 255).round().astype("uint8")
        return [PIL.Image.fromarray(a) for a in arr]
