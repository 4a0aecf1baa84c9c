"""-m gpu: LayerNorm folded into the GEMMs around it (fd_gemm_ln; the DEFAULT path of the UNet's transformer blocks since round 2,
FD_LN_FOLD=0 switches back to LayerNorm launches).  Producer statistics == row sums of the stored bf16 output; consumer ==
layer_norm(x) W^T + b (linear and GEGLU epilogues, single-CTA and CTA-pair tiles); a transformer stack with the fold on == the
same stack with LayerNorm launches."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize("M,C,N,act,mean", [(4096, 640, 1920, 0, 0.0), (2048, 1280, 3840, 0, 0.5), (1000, 1280, 1280, 0, 3.0),
                                            (4096, 640, 5120, 2, 0.2), (6144, 1280, 10240, 2, 1.0)])
def test_producer_statistics_and_folded_consumer(M, C, N, act, mean):
    from faithdiff_dev_b200 import ops
    from faithdiff_dev_b200.blocks import fold_layernorm, interleave_geglu
    torch.manual_seed(M + N)
    a = torch.randn(M, C, device="cuda").to(BF)
    wo = (torch.randn(C, C, device="cuda") / C ** 0.5).to(BF)
    bo = torch.randn(C, device="cuda") * 0.1 + mean
    res = (torch.randn(M, C, device="cuda") * 2).to(BF)
    plain = ops.gemm(a, wo, bias=bo, residual=res)
    h, st = ops.gemm(a, wo, bias=bo, residual=res, stats_out=True)
    assert torch.equal(h, plain)                                            # the producer's output is untouched
    s = st.slots[:, :st.n].double().sum(1)
    assert st.n >= 2 and st.n <= st.slots.shape[1]
    assert torch.allclose(s[:, 0], h.double().sum(1), rtol=1e-5, atol=1e-2)
    assert torch.allclose(s[:, 1], (h.double() ** 2).sum(1), rtol=1e-5, atol=1e-2)
    gamma, beta = 1 + 0.3 * torch.randn(C, device="cuda"), 0.2 * torch.randn(C, device="cuda")
    w = (torch.randn(N, C, device="cuda") / C ** 0.5).to(BF)
    b = torch.randn(N, device="cuda") * 0.1
    if act == 2:
        w, b = interleave_geglu(w, b)
    # fp32 truth WITHOUT the bf16 rounding of the LayerNorm output: that rounding belongs to the default path only, and a
    # reference that contains it is biased towards it (first GPU run: folded 3.7e-3 vs default 1.7e-3 against the rounded
    # reference for the GEGLU case, where the CPU study against fp64 has the folded form ahead, 2.3e-3 vs 2.9e-3)
    ln = F.layer_norm(h.float(), (C,), gamma, beta, 1e-5)
    lin = ln @ w.float().t() + b
    ref = lin[:, 0::2] * F.gelu(lin[:, 1::2]) if act == 2 else lin
    base = ops.gemm(ops.layernorm(h, gamma, beta), w, bias=b, act=act)      # the default path
    wp, cs, bp = fold_layernorm(w, b, gamma, beta)
    out = ops.gemm(h, wp, bias=bp, act=act, ln_in=st, ln_colsum=cs)
    assert out.shape == base.shape
    assert rel(out, ref) < 6e-3 and rel(out, ref) < 1.5 * rel(base, ref) + 1e-3, (rel(out, ref), rel(base, ref))


def test_transformer_stack_with_fold_matches_layernorm_path(monkeypatch):
    from faithdiff_dev_b200 import blocks
    from faithdiff_dev_b200.blocks import Transformer2D, Weights
    torch.manual_seed(7)
    Cc, heads, depth, B, H = 640, 10, 2, 2, 32
    g = torch.Generator().manual_seed(1)
    sd = {"norm.weight": torch.ones(Cc), "norm.bias": torch.zeros(Cc),
          "proj_in.weight": torch.randn(Cc, Cc, generator=g) / Cc ** 0.5, "proj_in.bias": torch.zeros(Cc),
          "proj_out.weight": torch.randn(Cc, Cc, generator=g) / Cc ** 0.5, "proj_out.bias": torch.zeros(Cc)}
    for i in range(depth):
        p = "transformer_blocks.%d." % i
        for n in (1, 2, 3):
            sd[p + "norm%d.weight" % n] = 1 + 0.2 * torch.randn(Cc, generator=g)
            sd[p + "norm%d.bias" % n] = 0.1 * torch.randn(Cc, generator=g)
        for a, kin in (("attn1", Cc), ("attn2", 2048)):
            sd[p + a + ".to_q.weight"] = torch.randn(Cc, Cc, generator=g) / Cc ** 0.5
            sd[p + a + ".to_k.weight"] = torch.randn(Cc, kin, generator=g) / kin ** 0.5
            sd[p + a + ".to_v.weight"] = torch.randn(Cc, kin, generator=g) / kin ** 0.5
            sd[p + a + ".to_out.0.weight"] = torch.randn(Cc, Cc, generator=g) / Cc ** 0.5
            sd[p + a + ".to_out.0.bias"] = 0.05 * torch.randn(Cc, generator=g)
        sd[p + "ff.net.0.proj.weight"] = torch.randn(8 * Cc, Cc, generator=g) / Cc ** 0.5
        sd[p + "ff.net.0.proj.bias"] = 0.05 * torch.randn(8 * Cc, generator=g)
        sd[p + "ff.net.2.weight"] = torch.randn(Cc, 4 * Cc, generator=g) / (4 * Cc) ** 0.5
        sd[p + "ff.net.2.bias"] = 0.05 * torch.randn(Cc, generator=g)
    x = torch.randn(B, H, H, Cc, device="cuda").to(BF)
    text = torch.randn(B * 77, 2048, device="cuda").to(BF)
    outs = {}
    for fold in (False, True):
        monkeypatch.setattr(blocks, "LN_FOLD", fold)
        t = Transformer2D(Weights(sd, torch.device("cuda")), heads, 32)
        kv = [blk.text_kv(text) for blk in t.blocks]
        outs[fold] = t(x, kv, 77)
    assert rel(outs[True], outs[False]) < 8e-3, rel(outs[True], outs[False])
