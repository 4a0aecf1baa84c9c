"""Dev (GPU): isolated LayerNorm / GroupNorm timings vs the HBM roofline (CUDA events around a CUDA graph of back-to-back
launches over rotating buffers larger than L2 in total).  GroupNorm: one-launch persistent kernel vs the three-launch path."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faithdiff_dev_b200 import ops
from faithdiff_dev_b200._lib import lib, check
BF = torch.bfloat16
PEAK = 6454.9
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", PEAK)
except Exception:
    pass


def bench(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


if "ln" in sys.argv or len(sys.argv) == 1:
    for rows, cols in [(6144, 1280), (14336, 1280), (24576, 640), (57344, 640)]:
        nb = max(2, int(300e6 // (rows * cols * 2)))
        xs = [torch.randn(rows, cols, device="cuda").to(BF) for _ in range(nb)]
        gm, bt = torch.ones(cols, device="cuda"), torch.zeros(cols, device="cuda")
        out = torch.empty_like(xs[0])
        i = [0]
        def f():
            ops.layernorm(xs[i[0] % nb], gm, bt, 1e-5, out=out); i[0] += 1
        us = bench(f)
        gbs = rows * cols * 4 / us / 1e3
        print("layernorm %6d x %4d: %7.1f us  %5.0f GB/s  %.2f of HBM peak" % (rows, cols, us, gbs, gbs / PEAK), flush=True)

if "gn" in sys.argv or len(sys.argv) == 1:
    p = lambda t: None if t is None else t.data_ptr()
    for B, H, C in [(2, 128, 320), (6, 128, 320), (14, 128, 320), (14, 128, 640), (14, 64, 640), (14, 64, 1280), (14, 32, 1280), (14, 32, 2560),
                    (6, 32, 1280), (6, 64, 640)]:
        nb = max(2, min(6, int(400e6 // (B * H * H * C * 2))))
        xs = [torch.randn(B, H, H, C, device="cuda").to(BF) for _ in range(nb)]
        gm, bt = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        out = torch.empty_like(xs[0])
        work = torch.empty(lib.fd_groupnorm_work_floats(B, 32, C), device="cuda", dtype=torch.float32)
        i = [0]
        def fused():
            ops.groupnorm(xs[i[0] % nb], gm, bt, 32, 1e-5, True, out=out); i[0] += 1
        def three():
            check(lib.fd_groupnorm_nhwc(p(xs[i[0] % nb]), C, None, 0, p(gm), p(bt), p(out), B, H * H, 32, 1e-5, 1, p(work),
                                        torch.cuda.current_stream().cuda_stream), "gn"); i[0] += 1
        byts = B * H * H * C * 4
        u1, u3 = bench(fused), bench(three)
        print("groupnorm %2d x %3d^2 x %4d (%6.1f MB): one-launch %7.1f us %5.0f GB/s %.2f of peak | three-launch %7.1f us %5.0f GB/s %.2f"
              % (B, H, C, byts / 2e6, u1, byts / u1 / 1e3, byts / u1 / 1e3 / PEAK, u3, byts / u3 / 1e3, byts / u3 / 1e3 / PEAK), flush=True)

if "gns" in sys.argv or len(sys.argv) == 1:
    # GroupNorm from producer statistics (finalize + streaming apply) vs the one-launch two-pass kernel vs a plain copy
    for B, H, C in [(14, 128, 320), (14, 128, 640), (14, 64, 640), (14, 64, 1280), (14, 32, 1280), (14, 32, 2560), (8, 512, 128), (8, 256, 256)]:
        nb = max(2, min(6, int(400e6 // (B * H * H * C * 2))))
        xs = [torch.randn(B, H, H, C, device="cuda").to(BF) for _ in range(nb)]
        for x in xs:     # fabricate the producer's statistics: per (32-row box, channel pair) sum / sum of squares, mode-0 box layout
            v = x.view(B * H * H // 32, 32, C // 2, 2).float()
            ops.tag_stats(x, ops.ColStats(torch.stack([v.sum((1, 3)), (v * v).sum((1, 3))], -1).contiguous(), C // 2, H * H // 32))
        gm, bt = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        out = torch.empty_like(xs[0])
        i = [0]
        def stats():
            ops.groupnorm(xs[i[0] % nb], gm, bt, 32, 1e-5, True, out=out); i[0] += 1
        plain = [x.clone() for x in xs]
        def fused():
            ops.groupnorm(plain[i[0] % nb], gm, bt, 32, 1e-5, True, out=out); i[0] += 1
        def copy():
            out.copy_(plain[i[0] % nb]); i[0] += 1
        ref = ops.groupnorm(plain[0], gm, bt, 32, 1e-5, True)
        got = ops.groupnorm(xs[0], gm, bt, 32, 1e-5, True)
        err = ((got.float() - ref.float()).norm() / ref.float().norm()).item()
        byts = B * H * H * C * 4
        us, uf, uc = bench(stats), bench(fused), bench(copy)
        print("groupnorm %2d x %3d^2 x %4d (%6.1f MB): from producer stats %7.1f us %5.0f GB/s %.2f of peak | one-launch two-pass %7.1f us %.2f | "
              "torch copy %7.1f us %.2f | rel diff %.1e" % (B, H, C, byts / 2e6, us, byts / us / 1e3, byts / us / 1e3 / PEAK, uf, byts / uf / 1e3 / PEAK,
                                                            uc, byts / uc / 1e3 / PEAK, err), flush=True)
