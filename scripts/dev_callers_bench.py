"""Dev (GPU): timings of the callers either side of the path: GPU LANCZOS upscale vs PIL on the host
cores, and the BSRNet pre-clean (nb=23) on one 512x512 tile."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image
import faithdiff_dev_b200 as fd
from faithdiff_dev_b200 import ops
from oracle import bsrnet_oracle as B


def ev(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


a = np.random.default_rng(0).integers(0, 256, size=(1080, 1920, 3), dtype=np.uint8)
img = Image.fromarray(a)
t0 = time.perf_counter(); ref = np.asarray(img.resize((3840, 2160), Image.LANCZOS)); t_pil = (time.perf_counter() - t0) * 1e3
dev = torch.from_numpy(a).cuda()
fd.resize_lanczos(dev, (3840, 2160))
ms = ev(lambda: fd.resize_lanczos(dev, (3840, 2160)))
host = torch.from_numpy(a).pin_memory()
ms_h = ev(lambda: fd.resize_lanczos(host, (3840, 2160)))
same = bool((fd.resize_lanczos(dev, (3840, 2160)).cpu().numpy() == ref).all())
traffic = a.size + 2 * 1080 * 3840 * 3 + 2160 * 3840 * 3
print("LANCZOS 1920x1080 -> 3840x2160: GPU %.3f ms (device input; %.0f GB/s of u8 traffic), %.3f ms incl. H2D of the 6.2 MB LQ image; "
      "PIL on the host %.1f ms; bit-identical: %s" % (ms, traffic / ms / 1e6, ms_h, t_pil, same), flush=True)

cfg = dict(in_nc=3, out_nc=3, nf=64, nb=23, gc=32, sf=4)
net = fd.RRDBNet(**cfg)
net.load_state_dict(B.random_state_dict(seed=3, **cfg))
x = torch.rand(1, 3, 512, 512, device="cuda")
n0 = ops.launch_count()
net.deg_remove(x)
launches = ops.launch_count() - n0
ms = ev(lambda: net.deg_remove(x), 3)
flop = 2 * 9 * (512 * 512 * (3 * 64 + 69 * (64 * 32 + 96 * 32 + 128 * 32 + 160 * 32 + 192 * 64) + 64 * 64) + 1024 * 1024 * 64 * 64
                + 2048 * 2048 * (2 * 64 * 64 + 64 * 3))
print("BSRNet deg_remove 512x512 -> 2048x2048 (nb=23): %.1f ms, %d launches, %.1f TFLOP -> %.0f TFLOP/s" % (ms, launches, flop / 1e12, flop / ms / 1e9), flush=True)
