"""B200-native drop-in for the reference's `UNet2DConditionModel` (FaithDiff extension of the SDXL UNet):
FaithDiff/models/unet_2d_condition_vae_extension.py:269-697.  Same constructor kwargs, `forward` signature,
`load_additional_layers`, `load_state_dict` key layout, `.config`, `.dtype`, `.denoise_encoder`; the arithmetic
runs in hand-written sm_100a kernels over NHWC bf16 activations (no torch.nn compute, no fallback)."""
from dataclasses import dataclass

import torch

from . import ops
from .blocks import BF16, Downsample, ResnetBlock, Transformer2D, Upsample, Weights
from .vae import Encoder


class Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor = None


_DEFAULTS = dict(
    sample_size=None, in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn",
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, downsample_padding=1,
    mid_block_scale_factor=1, dropout=0.0, act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1280,
    transformer_layers_per_block=1, reverse_transformer_layers_per_block=None, encoder_hid_dim=None,
    encoder_hid_dim_type=None, attention_head_dim=8, num_attention_heads=None, dual_cross_attention=False,
    use_linear_projection=False, class_embed_type=None, addition_embed_type=None, addition_time_embed_dim=None,
    num_class_embeds=None, upcast_attention=False, resnet_time_scale_shift="default", resnet_skip_time_act=False,
    resnet_out_scale_factor=1.0, time_embedding_type="positional", time_embedding_dim=None, time_embedding_act_fn=None,
    timestep_post_act=None, time_cond_proj_dim=None, conv_in_kernel=3, conv_out_kernel=3,
    projection_class_embeddings_input_dim=None, attention_type="default", class_embeddings_concat=False,
    mid_block_only_cross_attention=None, cross_attention_norm=None, addition_embed_type_num_heads=64)


class _StemBlock:
    """ResidualAttentionBlock (…extension.py:238-260)."""

    def __init__(self, W, heads):
        self.heads = heads
        self.ln1 = (W.f32("ln_1.weight"), W.f32("ln_1.bias"))
        self.ln2 = (W.f32("ln_2.weight"), W.f32("ln_2.bias"))
        self.win, self.bin = W.mat("attn.in_proj_weight"), W.f32("attn.in_proj_bias")
        self.wout, self.bout = W.mat("attn.out_proj.weight"), W.f32("attn.out_proj.bias")
        self.wfc, self.bfc = W.mat("mlp.c_fc.weight"), W.f32("mlp.c_fc.bias")
        self.wpr, self.bpr = W.mat("mlp.c_proj.weight"), W.f32("mlp.c_proj.bias")
        self.E = self.wout.shape[0]
        self.win_v, self.bin_v = self.win[2 * self.E:].contiguous(), self.bin[2 * self.E:].contiguous()

    def __call__(self, x, groups, seq, hw):
        y = ops.layernorm(x, *self.ln1)
        if seq == 1:
            # one element on the attended (batch) axis — the tile-step evaluates the stem once per tile because its CFG halves are
            # identical: the softmax is exactly 1 and the attention output is V, so only the V third of in_proj is computed
            a = ops.gemm(y, self.win_v, bias=self.bin_v)
        else:
            qkv = ops.gemm(y, self.win, bias=self.bin)
            a = ops.batch_axis_attention(qkv, groups, seq, hw, self.E, self.heads)
        x = ops.gemm(a, self.wout, bias=self.bout, residual=x)
        y = ops.layernorm(x, *self.ln2)
        h = ops.gemm(y, self.wfc, bias=self.bfc, act=3)
        return ops.gemm(h, self.wpr, bias=self.bpr, residual=x)


class UNet2DConditionModel:
    def __init__(self, **kwargs):
        cfg = dict(_DEFAULTS)
        unknown = set(kwargs) - set(cfg)
        if unknown:
            raise TypeError("unknown UNet config keys: %s" % sorted(unknown))
        cfg.update(kwargs)
        self.config = Config(cfg)
        c = self.config
        if not (c.use_linear_projection and c.addition_embed_type == "text_time" and c.act_fn == "silu"
                and c.time_embedding_type == "positional" and c.mid_block_type == "UNetMidBlock2DCrossAttn"
                and c.class_embed_type is None and c.encoder_hid_dim is None and not c.dual_cross_attention
                and c.resnet_time_scale_shift == "default" and c.conv_in_kernel == 3 and c.conv_out_kernel == 3):
            raise NotImplementedError("only the SDXL-style UNet configuration of the reference is implemented")
        self.device = torch.device("cuda")
        self.dtype = BF16
        self.denoise_encoder = None
        self.information_transformer_layes = None
        self.condition_embedding = None
        self.spatial_ch_projs = None
        self.agg_net = None
        self._loaded = False
        self._kv_cache = (None, None, None)
        self.num_upsamplers = len(c.block_out_channels) - 1

    # ------------------------------------------------------------------ weights
    def load_additional_layers(self, dtype=None, channel=512, weight_path=None):
        """…extension.py:405-433.  Modules are materialised by load_state_dict; weight_path is a torch pickle
        (FaithDiff.bin) holding the full UNet + FaithDiff state dict."""
        self._lq_channels = channel
        if weight_path is not None:
            self.load_state_dict(torch.load(weight_path, weights_only=False, map_location="cpu"), strict=True)

    def to(self, *args, **kwargs):
        return self

    def eval(self):
        return self

    def load_state_dict(self, state_dict, strict=True, device="cuda"):
        """Accepts the reference's key layout (…extension.py:444-478): stock SDXL UNet keys plus the prefixes
        denoise_encoder. / information_transformer_layes. / condition_embedding. / spatial_ch_projs."""
        c = self.config
        dev = torch.device(device)
        self.device = dev
        W = Weights(state_dict, dev)
        n = len(c.block_out_channels)
        heads = c.num_attention_heads or c.attention_head_dim
        heads = (heads,) * n if isinstance(heads, int) else tuple(heads)
        tl = c.transformer_layers_per_block
        tl = (tl,) * n if isinstance(tl, int) else tuple(tl)
        lpb = c.layers_per_block
        lpb = (lpb,) * n if isinstance(lpb, int) else tuple(lpb)
        g, eps = c.norm_num_groups, c.norm_eps
        self.temb_dim = c.block_out_channels[0] * 4
        # embeddings
        self.te1 = (W.mat("time_embedding.linear_1.weight"), W.f32("time_embedding.linear_1.bias"))
        self.te2 = (W.mat("time_embedding.linear_2.weight"), W.f32("time_embedding.linear_2.bias"))
        self.ae1 = (W.mat("add_embedding.linear_1.weight"), W.f32("add_embedding.linear_1.bias"))
        self.ae2 = (W.mat("add_embedding.linear_2.weight"), W.f32("add_embedding.linear_2.bias"))
        self.conv_in = (W.conv3("conv_in.weight"), W.f32("conv_in.bias"))
        # trunk; every resnet gets a slice of one stacked time_emb_proj linear
        self._resnets = []

        def resnet(prefix):
            r = ResnetBlock(W.sub(prefix), g, eps)
            self._resnets.append(r)
            return r

        self.down = []
        for i, t in enumerate(c.down_block_types):
            blk = {"resnets": [], "attn": [], "down": None}
            for j in range(lpb[i]):
                blk["resnets"].append(resnet("down_blocks.%d.resnets.%d." % (i, j)))
                if t == "CrossAttnDownBlock2D":
                    blk["attn"].append(Transformer2D(W.sub("down_blocks.%d.attentions.%d." % (i, j)), heads[i], g))
            if i != n - 1:
                blk["down"] = Downsample(W.sub("down_blocks.%d.downsamplers.0." % i), c.downsample_padding)
            self.down.append(blk)
        self.mid = {"resnets": [resnet("mid_block.resnets.0."), resnet("mid_block.resnets.1.")],
                    "attn": [Transformer2D(W.sub("mid_block.attentions.0."), heads[-1], g)]}
        self.up = []
        rh = heads[::-1]
        for i, t in enumerate(c.up_block_types):
            blk = {"resnets": [], "attn": [], "up": None}
            for j in range(lpb[::-1][i] + 1):
                blk["resnets"].append(resnet("up_blocks.%d.resnets.%d." % (i, j)))
                if t == "CrossAttnUpBlock2D":
                    blk["attn"].append(Transformer2D(W.sub("up_blocks.%d.attentions.%d." % (i, j)), rh[i], g))
            if i != n - 1:
                blk["up"] = Upsample(W.sub("up_blocks.%d.upsamplers.0." % i))
            self.up.append(blk)
        off, tws, tbs = 0, [], []
        for r in self._resnets:
            r.temb_offset = off
            off += r.cout
            tws.append(r.tw)
            tbs.append(r.tb)
            r.tw = True
        self.temb_w, self.temb_b = torch.cat(tws, 0).contiguous(), torch.cat(tbs, 0).contiguous()
        self.norm_out = (W.f32("conv_norm_out.weight"), W.f32("conv_norm_out.bias"))
        self.conv_out = (W.conv3("conv_out.weight"), W.f32("conv_out.bias"))
        self._transformers = [a for b in self.down for a in b["attn"]] + self.mid["attn"] + [a for b in self.up for a in b["attn"]]
        # FaithDiff stem
        if W.has("spatial_ch_projs.weight"):
            cw = W.sub("condition_embedding.")
            self.condition_embedding = dict(ng=cw.f32("norm_in.weight"), nb=cw.f32("norm_in.bias"),
                                            w1=cw.conv3("conv_in.weight"), b1=cw.f32("conv_in.bias"),
                                            w2=cw.conv3("conv_out.weight"), b2=cw.f32("conv_out.bias"))
            self.information_transformer_layes = []
            i = 0
            while W.has("information_transformer_layes.%d.ln_1.weight" % i):
                self.information_transformer_layes.append(_StemBlock(W.sub("information_transformer_layes.%d." % i), 8))
                i += 1
            self.spatial_ch_projs = (W.mat("spatial_ch_projs.weight"), W.f32("spatial_ch_projs.bias"))
        if W.has("denoise_encoder.conv_in.weight"):
            self.denoise_encoder = Encoder()
            self.denoise_encoder.load_state_dict({k[len("denoise_encoder."):]: v for k, v in state_dict.items()
                                                  if k.startswith("denoise_encoder.")}, device=device)
            W.used.update(k for k in state_dict if k.startswith("denoise_encoder."))
        if strict:
            missing = [k for k in state_dict if k not in W.used and not k.startswith("agg_net.")]
            if missing:
                raise RuntimeError("unexpected keys in state_dict: %s…" % missing[:5])
        self._loaded = True
        self._kv_cache = (None, None, None)
        self._weights_version = getattr(self, "_weights_version", 0) + 1     # pipelines key their graphs / text K/V caches on it
        return self

    # ------------------------------------------------------------------ pieces
    def time_embeddings(self, timestep, text_embeds, time_ids, batch):
        """…extension.py:543-564 → per-resnet time biases [batch, ΣC_out] fp32 (silu(emb) @ time_emb_proj)."""
        c = self.config
        dev = self.device
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([float(timestep)], dtype=torch.float32, device=dev)
        t = timestep.to(dev, torch.float32).reshape(-1)
        if t.numel() == 1:
            t = t.expand(batch)
        t = t.contiguous()
        t_sin = ops.timestep_embedding(t, c.block_out_channels[0])
        e = ops.linear_smallm(t_sin, *self.te1, silu_out=True)
        emb = ops.linear_smallm(e, *self.te2)
        # text_time augmentation; time_ids are rounded through the working dtype first as the reference does
        # (pipeline_FaithDiff_tlc.py:1344-1345)
        ids = time_ids.to(dev).to(torch.float32).reshape(-1).contiguous()
        nid = ids.numel() // batch
        add_in = torch.empty((batch, self.ae1[0].shape[1]), device=dev, dtype=torch.float32)
        tdim = text_embeds.shape[-1]
        add_in[:, :tdim] = text_embeds.to(dev, torch.float32)
        ops.timestep_embedding(ids, c.addition_time_embed_dim, out=add_in[:, tdim:], inner=nid, outer_ld=add_in.stride(0))
        a = ops.linear_smallm(add_in, *self.ae1, silu_out=True)
        emb = ops.linear_smallm(a, *self.ae2, addend=emb)
        return ops.linear_smallm(emb, self.temb_w, self.temb_b, silu_in=True)

    def time_embedding_table(self, timesteps, text_embeds, time_ids):
        """All per-resnet time biases of an image at once: timesteps [S], text_embeds [P,·] (P = CFG branches),
        time_ids [T,6] (one row per tile) → fp32 [S, T, P, ΣC_out].  Same arithmetic as time_embeddings()."""
        c = self.config
        dev = self.device
        S, T, P = timesteps.numel(), time_ids.shape[0], text_embeds.shape[0]
        t = timesteps.to(dev, torch.float32).reshape(-1).contiguous()
        e = ops.linear_smallm(ops.timestep_embedding(t, c.block_out_channels[0]), *self.te1, silu_out=True)
        t_emb = ops.linear_smallm(e, *self.te2)                                       # [S, D]
        ids = time_ids.to(dev).to(torch.float32).contiguous()
        tdim = text_embeds.shape[-1]
        add_in = torch.empty((T, P, self.ae1[0].shape[1]), device=dev, dtype=torch.float32)
        add_in[:, :, :tdim] = text_embeds.to(dev, torch.float32)[None]
        sin = ops.timestep_embedding(ids.reshape(-1), c.addition_time_embed_dim).reshape(T, 1, -1)
        add_in[:, :, tdim:] = sin
        a = ops.linear_smallm(add_in.reshape(T * P, -1), *self.ae1, silu_out=True)
        aug = ops.linear_smallm(a, *self.ae2)                                         # [T*P, D]
        emb = (t_emb[:, None, :] + aug[None, :, :]).reshape(S * T * P, -1).contiguous()
        return ops.linear_smallm(emb, self.temb_w, self.temb_b, silu_in=True).reshape(S, T, P, -1)

    def text_kv(self, encoder_hidden_states):
        """Cross-attention K/V of the text tokens for all 70 transformer blocks; identical for every tile and
        step, so cached on the tensor's identity + version counter."""
        c = self._kv_cache
        if c[0] is encoder_hidden_states and c[1] == encoder_hidden_states._version:
            return c[2]                 # the cache holds a reference, so the tensor's storage cannot be recycled
        ehs = encoder_hidden_states.to(self.device, BF16).contiguous()
        B, lk, D = ehs.shape
        t2 = ehs.reshape(B * lk, D)
        kv = [[blk.text_kv(t2) for blk in tr.blocks] for tr in self._transformers]
        self._kv_cache = (encoder_hidden_states, encoder_hidden_states._version, (kv, lk))
        return self._kv_cache[2]

    def condition_embed(self, lq_nhwc):
        """ControlNetConditioningEmbedding (…extension.py:213-219): step-invariant per tile."""
        ce = self.condition_embedding
        g = ops.groupnorm(lq_nhwc, ce["ng"], ce["nb"], 32, 1e-6, False)
        e = ops.conv3x3(g, ce["w1"], bias=ce["b1"], act=1)
        return ops.conv3x3(e, ce["w2"], bias=ce["b2"])

    def stem(self, sample_nchw, cond_nhwc, add_sample=True, groups=1, cat=None):
        """conv_in + FaithDiff fusion (…extension.py:571-579).  sample NCHW [B,4,h,w]; cond NHWC [B,h,w,320] or
        None.  The information transformer attends across the batch axis (seq = B // groups per pixel).
        `cat` (optional, [B,h,w,2*c0]): caller-owned stem input whose upper channel half already holds the condition
        embedding (the tile-step graph writes it there directly); conv_in fills the lower half."""
        B, _, h, w = sample_nchw.shape
        c0 = self.conv_in[0].shape[0]
        if cond_nhwc is None and cat is None:
            return ops.conv3x3_small_in(sample_nchw.contiguous(), *self.conv_in)
        if cat is None:
            cat = torch.empty((B, h, w, 2 * c0), device=self.device, dtype=BF16)
            cat[..., c0:] = cond_nhwc        # strided D2D copy of the cached, step-invariant condition embedding
        ops.conv3x3_small_in(sample_nchw.contiguous(), *self.conv_in, out=cat[..., :c0])
        x = cat.view(B * h * w, 2 * c0)
        for blk in self.information_transformer_layes:
            x = blk(x, groups, B // groups, h * w)
        res = cat.view(B * h * w, 2 * c0)[:, :c0] if add_sample else None
        out = ops.gemm(x, *self.spatial_ch_projs[:1], bias=self.spatial_ch_projs[1], residual=res, colstats_rows_per_sample=h * w)
        return ops.tag_stats(out.view(B, h, w, c0), getattr(out, "_fd_colstats", None))

    def trunk(self, x, temb_all, kv, lk, out_dtype):
        """down / mid / up / out (…extension.py:611-690) on NHWC bf16."""
        c = self.config
        up_factor = 2 ** self.num_upsamplers
        fwd_up = any(d % up_factor != 0 for d in x.shape[1:3])
        ti = 0
        skips = [x]
        for blk in self.down:
            for j, r in enumerate(blk["resnets"]):
                x = r(x, temb_all)
                if blk["attn"]:
                    x = blk["attn"][j](x, kv[ti], lk)
                    ti += 1
                skips.append(x)
            if blk["down"] is not None:
                x = blk["down"](x)
                skips.append(x)
        x = self.mid["resnets"][0](x, temb_all)
        x = self.mid["attn"][0](x, kv[ti], lk)
        ti += 1
        x = self.mid["resnets"][1](x, temb_all)
        for i, blk in enumerate(self.up):
            for j, r in enumerate(blk["resnets"]):
                x = r(x, temb_all, x2=skips.pop())
                if blk["attn"]:
                    x = blk["attn"][j](x, kv[ti], lk)
                    ti += 1
            if blk["up"] is not None:
                size = skips[-1].shape[1:3] if fwd_up else None
                x = blk["up"](x, size)
        n = ops.groupnorm(x, *self.norm_out, c.norm_num_groups, c.norm_eps, True)
        return ops.conv3x3_small_out(n, *self.conv_out, out_dtype=out_dtype, nchw_out=True)

    # ------------------------------------------------------------------ reference-facing forward
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None, attention_mask=None,
                cross_attention_kwargs=None, added_cond_kwargs=None, down_block_additional_residuals=None,
                mid_block_additional_residual=None, down_intrablock_additional_residuals=None,
                encoder_attention_mask=None, input_embedding=None, add_sample=True, return_dict=True,
                use_condition_embedding=True):
        if not self._loaded:
            raise RuntimeError("UNet2DConditionModel: load_state_dict() has not been called")
        for name, v in (("class_labels", class_labels), ("timestep_cond", timestep_cond), ("attention_mask", attention_mask),
                        ("down_block_additional_residuals", down_block_additional_residuals),
                        ("mid_block_additional_residual", mid_block_additional_residual),
                        ("down_intrablock_additional_residuals", down_intrablock_additional_residuals),
                        ("encoder_attention_mask", encoder_attention_mask)):
            if v is not None:
                raise NotImplementedError("%s is not supported on the B200 hot path" % name)
        if cross_attention_kwargs:
            raise NotImplementedError("cross_attention_kwargs (LoRA scale / GLIGEN) are not supported")
        if not sample.is_cuda:
            raise RuntimeError("faithdiff_dev_b200 has no CPU path: `sample` must be a CUDA tensor")
        B = sample.shape[0]
        temb_all = self.time_embeddings(timestep, added_cond_kwargs["text_embeds"], added_cond_kwargs["time_ids"], B)
        kv, lk = self.text_kv(encoder_hidden_states)
        cond = None
        if input_embedding is not None and self.condition_embedding is not None \
                and self.information_transformer_layes is not None:
            lq = ops.nchw_to_nhwc(input_embedding.contiguous())
            cond = self.condition_embed(lq) if use_condition_embedding else lq
        x = self.stem(sample, cond, add_sample=add_sample, groups=1)
        out = self.trunk(x, temb_all, kv, lk, sample.dtype)
        if not return_dict:
            return (out,)
        return UNet2DConditionOutput(sample=out)

    __call__ = forward
