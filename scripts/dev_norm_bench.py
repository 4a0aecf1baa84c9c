"""Dev (GPU): isolated LayerNorm / GroupNorm timings vs HBM roofline (events over back-to-back launches inside a graph)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faithdiff_dev_b200 import ops
BF = torch.bfloat16
def bench(fn, n=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for rows, cols in [(6144, 1280), (4096, 1280), (24576, 640), (49152, 640)]:
    xs = [torch.randn(rows, cols, device="cuda").to(BF) for _ in range(8)]   # rotate buffers (> L2 in total for the big ones)
    gm, bt = torch.ones(cols, device="cuda"), torch.zeros(cols, device="cuda")
    out = torch.empty_like(xs[0])
    i = [0]
    def f():
        ops.layernorm(xs[i[0] % 8], gm, bt, 1e-5, out=out); i[0] += 1
    us = bench(f)
    print("layernorm %6d x %4d: %6.1f us  %.0f GB/s" % (rows, cols, us, rows * cols * 4 / us / 1e3))
for B, H, C in [(6, 32, 1280), (6, 64, 640), (6, 128, 320), (6, 32, 2560)]:
    xs = [torch.randn(B, H, H, C, device="cuda").to(BF) for _ in range(4)]
    gm, bt = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    i = [0]
    out = torch.empty_like(xs[0])
    def f():
        ops.groupnorm(xs[i[0] % 4], gm, bt, 32, 1e-5, True, out=out); i[0] += 1
    us = bench(f, 20)
    print("groupnorm %d x %d^2 x %4d: %6.1f us  %.0f GB/s" % (B, H, C, us, B * H * H * C * 4 / us / 1e3))
