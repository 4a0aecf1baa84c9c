"""Dev (GPU, for ncu): encodes one synthetic 3840 x 2160 RGB image a few times with the device PNG encoder."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faithdiff_dev_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = torch.Generator(device="cuda").manual_seed(0)
yy, xx = torch.meshgrid(torch.arange(2160, device="cuda"), torch.arange(3840, device="cuda"), indexing="ij")
base = torch.stack([128 + 100 * torch.sin(xx / 37.0 + k) * torch.cos(yy / 23.0 - k) for k in range(3)], -1)
img = (base + 4 * torch.randn(base.shape, device="cuda", generator=g)).clamp(0, 255).to(torch.uint8)
torch.cuda.synchronize()
for _ in range(n):
    out, result = ops.png_encode(img)
torch.cuda.synchronize()
print("file bytes", result.tolist())
