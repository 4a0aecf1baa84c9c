"""Dev (GPU): the FaithDiff stem's MLP GEMMs at the tile-step's launch size (7 tiles x 128^2 tokens, K = 640 / 1280), CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faithdiff_dev_b200 import ops
BF = torch.bfloat16
M = 7 * 16384
x = torch.randn(M, 640, device="cuda").to(BF)
w1 = (torch.randn(1280, 640, device="cuda") / 25).to(BF); b1 = torch.zeros(1280, device="cuda")
w2 = (torch.randn(640, 1280, device="cuda") / 36).to(BF); b2 = torch.zeros(640, device="cuda")
for name, fn, fl in [("c_fc + quick-GELU  M=%d N=1280 K=640" % M, lambda: ops.gemm(x, w1, bias=b1, act=3), 2.0 * M * 1280 * 640),
                     ("c_fc, no activation", lambda: ops.gemm(x, w1, bias=b1), 2.0 * M * 1280 * 640),
                     ("c_fc + SiLU", lambda: ops.gemm(x, w1, bias=b1, act=1), 2.0 * M * 1280 * 640)]:
    for _ in range(3):
        h = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        h = fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-44s %7.1f us  %7.1f TFLOP/s" % (name, ms * 1e3, fl / ms / 1e9))
ref = torch.nn.functional.linear(x[:4096].float(), w1.float(), b1)
got = ops.gemm(x[:4096].contiguous(), w1, bias=b1, act=3)
print("quick-GELU rel-L2 vs torch fp32: %.2e" % ((got.float() - ref * torch.sigmoid(1.702 * ref)).norm() / (ref * torch.sigmoid(1.702 * ref)).norm()).item())
got = ops.gemm(x[:4096].contiguous(), w1, bias=b1, act=1)
print("SiLU rel-L2 vs torch fp32: %.2e" % ((got.float() - torch.nn.functional.silu(ref)).norm() / torch.nn.functional.silu(ref).norm()).item())
