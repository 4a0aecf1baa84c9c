"""CPU study for DESIGN.md §10 item 1 (LayerNorm folded into the consumer GEMM): error of
    rstd_r * (x @ W'^T - mean_r * s) + b'      with W' = bf16(W * gamma), s = rowsum(W'), b' = b + W @ beta, fp32 accumulation
against the reference order  bf16(LN(x)) @ bf16(W)^T + b  — both measured against an fp64 evaluation — as a function of the
row's |mean| / std ratio (the cancellation risk).  x is bf16 (the residual stream as stored), C = 1280, N = 512 outputs."""
import torch

torch.manual_seed(0)
C, N, M = 1280, 512, 512
bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
W = torch.randn(N, C) / C ** 0.5
gamma, beta, b = 1 + 0.3 * torch.randn(C), 0.2 * torch.randn(C), 0.1 * torch.randn(N)
print("%10s %14s %14s %14s" % ("|mean|/std", "reference bf16", "folded", "folded, stats from bf16 x (fp32 sums)"))
for ratio in (0.0, 0.3, 1.0, 3.0, 10.0, 30.0, 100.0):
    x = bf(torch.randn(M, C) * (1 + torch.rand(M, 1)) + ratio * (1 + torch.rand(M, 1)) * torch.sign(torch.randn(M, 1)))
    xd = x.double()
    mean, var = xd.mean(1, keepdim=True), xd.var(1, unbiased=False, keepdim=True)
    truth = ((xd - mean) / (var + 1e-5).sqrt() * gamma.double() + beta.double()) @ W.double().t() + b.double()
    ln32 = torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-5)
    ref = bf(bf(ln32) @ bf(W).t() + b)                                  # what the reference computes (fp32 accumulate, bf16 out)
    Wp = bf(W * gamma)
    s = Wp.sum(1)
    bp = b + W @ beta
    def folded(mean_r, rstd_r):
        return bf(rstd_r * (x @ Wp.t() - mean_r * s) + bp)
    f_exact = folded(mean.float(), (1 / (var + 1e-5).sqrt()).float())
    s1, s2 = x.sum(1, keepdim=True), (x * x).sum(1, keepdim=True)       # one-pass fp32 sums, as an epilogue would accumulate them
    m32 = s1 / C
    v32 = (s2 / C - m32 * m32).clamp_min(0)
    f_onepass = folded(m32, torch.rsqrt(v32 + 1e-5))
    rel = lambda a: ((a.double() - truth).norm() / truth.norm()).item()
    print("%10.1f %14.3e %14.3e %14.3e" % (ratio, rel(ref), rel(f_exact), rel(f_onepass)))
