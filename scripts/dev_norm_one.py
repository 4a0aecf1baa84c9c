"""Dev (GPU): a handful of GroupNorm / LayerNorm launches at tile-step shapes, for `ncu --set full -k regex:...`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faithdiff_dev_b200 import ops
BF = torch.bfloat16
for B, H, C in [(14, 128, 320), (14, 64, 640), (14, 32, 1280)]:
    x = torch.randn(B, H, H, C, device="cuda").to(BF)
    gm, bt = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    out = torch.empty_like(x)
    flush = torch.empty(64 << 20, device="cuda", dtype=torch.float32).zero_()     # 256 MB: x leaves L2
    ops.groupnorm(x, gm, bt, 32, 1e-5, True, out=out)
for rows, cols in [(14336, 1280), (57344, 640)]:
    x = torch.randn(rows, cols, device="cuda").to(BF)
    gm, bt = torch.ones(cols, device="cuda"), torch.zeros(cols, device="cuda")
    out = torch.empty_like(x)
    flush = torch.empty(64 << 20, device="cuda", dtype=torch.float32).zero_()
    ops.layernorm(x, gm, bt, 1e-5, out=out)
torch.cuda.synchronize()
print("ok")
