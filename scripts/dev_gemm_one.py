"""Dev (GPU): run one fd_gemm shape a few times (for ncu). usage: dev_gemm_one.py M N K [reps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faithdiff_dev_b200 import ops
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
kind = sys.argv[5] if len(sys.argv) > 5 else "bias"      # bias | res | geglu | plain
use_res = kind == "res"
act = 2 if kind == "geglu" else 0
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
out = torch.empty(M, N // 2 if act == 2 else N, device="cuda", dtype=torch.bfloat16)
res = torch.randn(M, N, device="cuda").to(torch.bfloat16) if use_res else None
bias = None if kind == "plain" else torch.randn(N, device="cuda")
for _ in range(reps):
    ops.gemm(a, w, out=out, residual=res, bias=bias, act=act)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(reps): ops.gemm(a, w, out=out, residual=res, bias=bias, act=act)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("M=%d N=%d K=%d: %.3f ms %.1f TFLOP/s" % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9))
