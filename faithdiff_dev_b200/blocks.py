"""Host-side building blocks shared by the UNet, the LQ encoder and the VAE.  Each block owns weights already
packed into the kernel layouts (bf16 K-major GEMM operands, fp32 biases / norm affine) and drives the C-ABI
kernels over NHWC bf16 activations.  State-dict keys are the reference's (diffusers 0.28 naming)."""
import os

import torch

from . import ops

BF16 = torch.bfloat16


class Weights:
    """Reads tensors of a reference-layout state dict and packs them for the kernels."""

    def __init__(self, sd, device, prefix=""):
        self.sd, self.device, self.prefix = sd, device, prefix
        self.used = set()

    def sub(self, p):
        w = Weights(self.sd, self.device, self.prefix + p)
        w.used = self.used
        return w

    def has(self, k):
        return self.prefix + k in self.sd

    def raw(self, k):
        self.used.add(self.prefix + k)
        return self.sd[self.prefix + k]

    def f32(self, k):
        return self.raw(k).detach().to(self.device, torch.float32).contiguous()

    def mat(self, k):
        return self.raw(k).detach().to(self.device, BF16).contiguous()

    def conv3(self, k):
        """[O,I,3,3] -> [O, 9*I] (tap-major, channel-minor)."""
        w = self.raw(k).detach().to(self.device)
        return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(BF16).contiguous()

    def conv1(self, k):
        w = self.raw(k).detach().to(self.device)
        return w.reshape(w.shape[0], w.shape[1]).to(BF16).contiguous()


class ResnetBlock:
    """diffusers ResnetBlock2D: GN→SiLU→conv3x3(+temb)→GN→SiLU→conv3x3, 1x1 shortcut when channels change."""

    def __init__(self, W, groups, eps, temb_offset=None):
        self.groups, self.eps = groups, eps
        self.n1g, self.n1b = W.f32("norm1.weight"), W.f32("norm1.bias")
        self.c1w, self.c1b = W.conv3("conv1.weight"), W.f32("conv1.bias")
        self.n2g, self.n2b = W.f32("norm2.weight"), W.f32("norm2.bias")
        self.c2w, self.c2b = W.conv3("conv2.weight"), W.f32("conv2.bias")
        self.cout = self.c1w.shape[0]
        self.temb_offset = temb_offset
        if W.has("time_emb_proj.weight"):
            self.tw, self.tb = W.mat("time_emb_proj.weight"), W.f32("time_emb_proj.bias")
        else:
            self.tw = None
        if W.has("conv_shortcut.weight"):
            self.sw, self.sb = W.conv1("conv_shortcut.weight"), W.f32("conv_shortcut.bias")
        else:
            self.sw = None

    def __call__(self, x, temb_all=None, x2=None):
        B, H, Wd = x.shape[:3]
        n1 = ops.groupnorm(x, self.n1g, self.n1b, self.groups, self.eps, True, x2=x2)
        rb = None
        if self.tw is not None:
            rb = temb_all[:, self.temb_offset:self.temb_offset + self.cout]
        h = ops.conv3x3(n1, self.c1w, bias=self.c1b, rowbias=rb, colstats=True)      # statistics for norm2 from the epilogue
        n2 = ops.groupnorm(h, self.n2g, self.n2b, self.groups, self.eps, True)
        if self.sw is not None:
            M = B * H * Wd
            res = ops.gemm(x.reshape(M, -1), self.sw, a2=None if x2 is None else x2.reshape(M, -1), bias=self.sb)
            res = res.view(B, H, Wd, -1)
        else:
            assert x2 is None
            res = x
        return ops.conv3x3(n2, self.c2w, bias=self.c2b, residual=res, colstats=True)   # ... for whichever GroupNorm reads the block's output


class VaeAttention:
    """diffusers Attention in the VAE / LQ-encoder mid block (unet_2d_condition_vae_extension.py:96-106): GroupNorm, 1 head with
    d = C, q/k/v/out with bias, residual.  C = 512 (the real networks): one fused QKV GEMM over all samples, the flash-style
    fd_attention_d512 kernel (no [L, L] matrix in HBM, no per-sample Python loop), out-projection + residual in the GEMM epilogue.
    Other widths (the tests' small configurations): scores and P·V as tensor-core GEMMs around a row-softmax kernel."""

    def __init__(self, W, groups, eps):
        self.groups, self.eps = groups, eps
        self.ng, self.nb = W.f32("group_norm.weight"), W.f32("group_norm.bias")
        self.wq, self.bq = W.mat("to_q.weight"), W.f32("to_q.bias")
        self.wk, self.bk = W.mat("to_k.weight"), W.f32("to_k.bias")
        self.wv, self.bv = W.mat("to_v.weight"), W.f32("to_v.bias")
        self.wo, self.bo = W.mat("to_out.0.weight"), W.f32("to_out.0.bias")
        self.wqkv = torch.cat([self.wq, self.wk, self.wv], 0).contiguous()
        self.bqkv = torch.cat([self.bq, self.bk, self.bv], 0).contiguous()

    def __call__(self, x):
        B, H, Wd, Cc = x.shape
        L = H * Wd
        n = ops.groupnorm(x, self.ng, self.nb, self.groups, self.eps, False)
        if Cc == 512 and os.environ.get("FD_VAE_ATTN_GEMM", "0") != "1":
            qkv = ops.gemm(n.reshape(B * L, Cc), self.wqkv, bias=self.bqkv)
            a = ops.attention_d512(qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:], B, L, Cc ** -0.5)
            o = ops.gemm(a, self.wo, bias=self.bo, residual=x.reshape(B * L, Cc), colstats_rows_per_sample=L)
            return ops.tag_stats(o.view(B, H, Wd, Cc), getattr(o, "_fd_colstats", None))
        out = torch.empty_like(x)
        Lp = (L + 7) // 8 * 8            # token count padded to the 16-byte vector width of the GEMM epilogue
        for b in range(B):
            nb = n[b].reshape(L, Cc)
            if Lp != L:
                pad = torch.zeros((Lp, Cc), device=x.device, dtype=BF16)
                pad[:L] = nb
                nb = pad
            q = ops.gemm(nb, self.wq, bias=self.bq)
            k = ops.gemm(nb, self.wk, bias=self.bk)
            vt = ops.gemm(self.wv, nb, bias=self.bv, bias_along_m=True)           # V^T [C, Lp] = Wv · X^T
            s = ops.gemm(q, k, out_fp32=True)                                      # [Lp, Lp] fp32 scores
            p = ops.softmax_rows(s, alpha=Cc ** -0.5, cols=L)                      # pad key columns get p = 0
            del s
            o = ops.gemm(p, vt)                                                    # [Lp, C]
            ops.gemm(o[:L], self.wo, bias=self.bo, residual=x[b].reshape(L, Cc), out=out[b].reshape(L, Cc))
        return out


class Downsample:
    def __init__(self, W, pad_lo):
        self.w, self.b, self.pad_lo = W.conv3("conv.weight"), W.f32("conv.bias"), pad_lo

    def __call__(self, x):
        return ops.conv3x3(x, self.w, bias=self.b, stride=2, pad_lo=self.pad_lo, colstats=True)


def upsample_phase_weights(w):
    """[O,I,3,3] fp32 -> four [O, 4*I] bf16 operands, one per output phase (a, b) of conv3x3(nearest_upsample_2x(x)): along each
    axis the three taps of phase 0 fall on input pixels {i-1, i, i} and of phase 1 on {i, i, i+1}, so the two taps that share a
    pixel are summed (in fp32, one bf16 rounding of the sum); tap order r'*2 + s', channel-minor."""
    def fold(t, axis, a):
        t0, t1, t2 = t.select(axis, 0), t.select(axis, 1), t.select(axis, 2)
        return torch.stack((t0, t1 + t2) if a == 0 else (t0 + t1, t2), dim=axis)
    wf = w.float()
    out = []
    for a in range(2):
        for b in range(2):
            ph = fold(fold(wf, 2, a), 3, b)                                        # [O, I, 2, 2]
            out.append(ph.permute(0, 2, 3, 1).reshape(ph.shape[0], -1).to(BF16).contiguous())
    return out


class Upsample:
    """diffusers Upsample2D: nearest x2 (or to an explicit size) + conv3x3.  The x2 case runs as four phase convs straight from
    the low-resolution tensor (ops.conv3x3_upsampled2x): no 4x tensor in HBM and 4/9 of the conv's FLOPs; FD_UPSAMPLE_FOLD=0 or
    an explicit odd `size` (forward_upsample_size, …extension.py:665-666) takes the materialising path."""

    def __init__(self, W):
        self.w, self.b = W.conv3("conv.weight"), W.f32("conv.bias")
        self.wp = upsample_phase_weights(W.raw("conv.weight").detach().to(W.device)) if os.environ.get("FD_UPSAMPLE_FOLD", "1") == "1" else None

    def __call__(self, x, size=None):
        if size is None and self.wp is not None:
            return ops.conv3x3_upsampled2x(x, self.wp, bias=self.b, colstats=True)
        u = ops.upsample_nearest(x) if size is None else ops.upsample_nearest(x, size[0], size[1])
        return ops.conv3x3(u, self.w, bias=self.b, colstats=True)


def interleave_geglu(w, b):
    """ff.net.0.proj [8C, C]: rows [0,4C) = value, [4C,8C) = gate → interleaved (v0,g0,v1,g1,…) so one accumulator
    tile holds both halves and the GEMM epilogue emits value·gelu(gate)."""
    half = w.shape[0] // 2
    wi = torch.stack([w[:half], w[half:]], dim=1).reshape(w.shape[0], w.shape[1]).contiguous()
    bi = torch.stack([b[:half], b[half:]], dim=1).reshape(-1).contiguous()
    return wi, bi


# The three LayerNorms of a transformer block are folded into the GEMMs around them (fd_gemm_ln): no LayerNorm launch, no
# normalised copy of the residual stream.  Default since round 2 (GPU-validated: kernel, stack and full-UNet
# parity; 4K restoration 8806 -> 8750 ms); FD_LN_FOLD=0 restores the LayerNorm launches for A/B runs.
LN_FOLD = os.environ.get("FD_LN_FOLD", "1") == "1"


def fold_layernorm(w, bias, gamma, beta):
    """(W [N,C] bf16, bias [N] fp32 or None, gamma, beta [C] fp32) -> (W' = bf16(W*gamma), colsum(W') fp32, bias + W beta)."""
    wf = w.float()
    wp = (wf * gamma[None, :]).to(BF16).contiguous()
    bp = wf @ beta
    if bias is not None:
        bp = bp + bias
    return wp, wp.float().sum(1).contiguous(), bp.float().contiguous()


class TransformerBlock:
    """diffusers BasicTransformerBlock: LN→self-attn→+, LN→cross-attn(text)→+, LN→GEGLU FF→+."""

    def __init__(self, W, heads):
        self.heads = heads
        self.ln = [(W.f32("norm%d.weight" % i), W.f32("norm%d.bias" % i)) for i in (1, 2, 3)]
        self.wqkv = torch.cat([W.mat("attn1.to_q.weight"), W.mat("attn1.to_k.weight"), W.mat("attn1.to_v.weight")], 0).contiguous()
        self.wo1, self.bo1 = W.mat("attn1.to_out.0.weight"), W.f32("attn1.to_out.0.bias")
        self.wq2 = W.mat("attn2.to_q.weight")
        self.wkv2 = torch.cat([W.mat("attn2.to_k.weight"), W.mat("attn2.to_v.weight")], 0).contiguous()
        self.wo2, self.bo2 = W.mat("attn2.to_out.0.weight"), W.f32("attn2.to_out.0.bias")
        self.wff1, self.bff1 = interleave_geglu(W.mat("ff.net.0.proj.weight"), W.f32("ff.net.0.proj.bias"))
        self.wff2, self.bff2 = W.mat("ff.net.2.weight"), W.f32("ff.net.2.bias")
        self.C = self.wo1.shape[0]
        self.scale = (self.C // heads) ** -0.5
        self.fold = LN_FOLD
        if self.fold:
            self.f_qkv = fold_layernorm(self.wqkv, None, *self.ln[0])
            self.f_q2 = fold_layernorm(self.wq2, None, *self.ln[1])
            self.f_ff1 = fold_layernorm(self.wff1, self.bff1, *self.ln[2])
            self.wqkv = self.wq2 = self.wff1 = self.bff1 = None      # only the folded copies stay resident

    def text_kv(self, text2d):
        return ops.gemm(text2d, self.wkv2)

    def call_folded(self, h, st, B, L, kv, lk):
        """Same block with the LayerNorms folded away: `st` = RowStats of h from the GEMM that produced it; returns (h, st)."""
        Cc = self.C
        w, cs, b = self.f_qkv
        qkv = ops.gemm(h, w, bias=b, ln_in=st, ln_colsum=cs)
        a = ops.attention(qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:], B, self.heads, L, L, self.scale)
        h, st = ops.gemm(a, self.wo1, bias=self.bo1, residual=h, stats_out=True)
        w, cs, b = self.f_q2
        q = ops.gemm(h, w, bias=b, ln_in=st, ln_colsum=cs)
        a = ops.attention(q, kv[:, :Cc], kv[:, Cc:], B, self.heads, L, lk, self.scale)
        h, st = ops.gemm(a, self.wo2, bias=self.bo2, residual=h, stats_out=True)
        w, cs, b = self.f_ff1
        g = ops.gemm(h, w, bias=b, act=2, ln_in=st, ln_colsum=cs)
        return ops.gemm(g, self.wff2, bias=self.bff2, residual=h, stats_out=True)

    def __call__(self, h, B, L, kv, lk):
        Cc = self.C
        y = ops.layernorm(h, *self.ln[0])
        qkv = ops.gemm(y, self.wqkv)
        a = ops.attention(qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:], B, self.heads, L, L, self.scale)
        h = ops.gemm(a, self.wo1, bias=self.bo1, residual=h)
        y = ops.layernorm(h, *self.ln[1])
        q = ops.gemm(y, self.wq2)
        a = ops.attention(q, kv[:, :Cc], kv[:, Cc:], B, self.heads, L, lk, self.scale)
        h = ops.gemm(a, self.wo2, bias=self.bo2, residual=h)
        y = ops.layernorm(h, *self.ln[2])
        g = ops.gemm(y, self.wff1, bias=self.bff1, act=2)
        return ops.gemm(g, self.wff2, bias=self.bff2, residual=h)


class Transformer2D:
    """diffusers Transformer2DModel (use_linear_projection): GN(eps 1e-6) → Linear in → N blocks → Linear out → + input."""

    def __init__(self, W, heads, groups):
        self.groups = groups
        self.ng, self.nb = W.f32("norm.weight"), W.f32("norm.bias")
        self.wi, self.bi = W.mat("proj_in.weight"), W.f32("proj_in.bias")
        self.wo, self.bo = W.mat("proj_out.weight"), W.f32("proj_out.bias")
        self.blocks = []
        i = 0
        while W.has("transformer_blocks.%d.norm1.weight" % i):
            self.blocks.append(TransformerBlock(W.sub("transformer_blocks.%d." % i), heads))
            i += 1

    def __call__(self, x, kv_list, lk):
        B, H, Wd, Cc = x.shape
        L = H * Wd
        n = ops.groupnorm(x, self.ng, self.nb, self.groups, 1e-6, False)
        if self.blocks and self.blocks[0].fold:
            h, st = ops.gemm(n.reshape(B * L, Cc), self.wi, bias=self.bi, stats_out=True)
            for blk, kv in zip(self.blocks, kv_list):
                h, st = blk.call_folded(h, st, B, L, kv, lk)
        else:
            h = ops.gemm(n.reshape(B * L, Cc), self.wi, bias=self.bi)
            for blk, kv in zip(self.blocks, kv_list):
                h = blk(h, B, L, kv, lk)
        out = ops.gemm(h, self.wo, bias=self.bo, residual=x.reshape(B * L, Cc), colstats_rows_per_sample=L)
        return ops.tag_stats(out.view(B, H, Wd, Cc), getattr(out, "_fd_colstats", None))
