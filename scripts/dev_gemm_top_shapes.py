"""Dev / profiling target (GPU): the GEMM / conv shapes that carry >= 70 % of a tile-step's FLOPs AT THE LAUNCH SIZE THE 4K BENCH USES
(7 tiles = 14 samples per launch), each launched once after a warm-up with the L2 flushed in between, for `ncu --set full`
(scripts/gpu_profile_r02.sh); without ncu it prints CUDA-event TFLOP/s per shape.  usage: dev_gemm_top_shapes.py [samples=14]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faithdiff_dev_b200 import ops
from faithdiff_dev_b200.blocks import interleave_geglu
BF = torch.bfloat16
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 14
torch.manual_seed(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def mk(*shape, s=1.0):
    return (torch.randn(*shape, device="cuda") * s).to(BF)


M2 = ns * 32 * 32          # 1280-channel level (32x32 per sample)
M1 = ns * 64 * 64          # 640-channel level
cases = []
x = mk(M2, 1280); w = mk(10240, 1280, s=0.03); b = torch.zeros(10240, device="cuda")
wi, bi = interleave_geglu(w, b)
cases.append(("FF1 GEGLU  M=%d N=10240 K=1280" % M2, 2.0 * M2 * 10240 * 1280, lambda: ops.gemm(x, wi, bias=bi, act=2)))
g = mk(M2, 5120); w2 = mk(1280, 5120, s=0.02); r = mk(M2, 1280); b2 = torch.zeros(1280, device="cuda")
cases.append(("FF2 + residual  M=%d N=1280 K=5120" % M2, 2.0 * M2 * 1280 * 5120, lambda: ops.gemm(g, w2, bias=b2, residual=r)))
wq = mk(3840, 1280, s=0.03)
cases.append(("QKV  M=%d N=3840 K=1280" % M2, 2.0 * M2 * 3840 * 1280, lambda: ops.gemm(x, wq)))
wo = mk(1280, 1280, s=0.03)
cases.append(("attn-out + residual  M=%d N=1280 K=1280" % M2, 2.0 * M2 * 1280 * 1280, lambda: ops.gemm(x, wo, bias=b2, residual=r)))
xc = mk(ns, 32, 32, 1280); wc = mk(1280, 9 * 1280, s=0.01)
cases.append(("conv3x3 1280->1280 @32x32 x%d" % ns, 2.0 * M2 * 1280 * 9 * 1280, lambda: ops.conv3x3(xc, wc, bias=b2)))
xc1 = mk(ns, 64, 64, 640); wc1 = mk(640, 9 * 640, s=0.01); b640 = torch.zeros(640, device="cuda")
cases.append(("conv3x3 640->640 @64x64 x%d" % ns, 2.0 * M1 * 640 * 9 * 640, lambda: ops.conv3x3(xc1, wc1, bias=b640)))
xc0 = mk(ns, 128, 128, 320); wc0 = mk(320, 9 * 320, s=0.02); b320 = torch.zeros(320, device="cuda")
cases.append(("conv3x3 320->320 @128x128 x%d" % ns, 2.0 * ns * 128 * 128 * 320 * 9 * 320, lambda: ops.conv3x3(xc0, wc0, bias=b320)))
x1 = mk(M1, 640); w1 = mk(5120, 640, s=0.04); b1 = torch.zeros(5120, device="cuda"); w1i, b1i = interleave_geglu(w1, b1)
cases.append(("FF1 GEGLU  M=%d N=5120 K=640" % M1, 2.0 * M1 * 5120 * 640, lambda: ops.gemm(x1, w1i, bias=b1i, act=2)))
for name, fl, fn in cases:          # warm-up (kernel attributes, tensor-map cache)
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, fl, fn in cases:
    flush.zero_()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print("%-46s %8.1f us  %7.1f TFLOP/s (cold L2)" % (name, e0.elapsed_time(e1) * 1e3, fl / e0.elapsed_time(e1) / 1e9), flush=True)
torch.cuda.profiler.stop()
