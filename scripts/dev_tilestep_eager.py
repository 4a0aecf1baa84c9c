"""Dev / profiling target (GPU): ONE eager UNet evaluation of an n-tile launch group at full size (the kernels a CUDA-graph replay of
the 4K bench executes: 7 tiles = 14 (tile, CFG-branch) samples per launch) followed by the fused per-step tail, with nothing else
in the process after the warm-up.  Used under ncu for the unfiltered launch list (time + DRAM bytes of EVERY kernel, ours and
torch's) — scripts/gpu_profile_r02.sh.   usage: dev_tilestep_eager.py [tiles=7] [reps=1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import faithdiff_dev_b200 as fd
from faithdiff_dev_b200 import arch, ops
from faithdiff_dev_b200.pipeline import TileStepGraph
BF = torch.bfloat16
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 7
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
unet = fd.UNet2DConditionModel(**arch.SDXL_UNET_CONFIG)
unet.load_state_dict(arch.random_state_dict(arch.unet_param_shapes(arch.SDXL_UNET_CONFIG), seed=1234))
g = torch.Generator(device="cuda").manual_seed(0)
pe = torch.randn(2, 77, 2048, generator=g, device="cuda").to(BF)
kv, lk = unet.text_kv(pe)
slot = tuple(i // 2 for i in range(2 * nt))
branch = tuple(i % 2 for i in range(2 * nt))
ts = TileStepGraph(unet, kv, lk, slot, branch, 128, 128, BF, True, use_graph=False)      # constructor = one warm-up evaluation
scaled = torch.randn(nt, 4, 128, 128, generator=g, device="cuda").to(BF)
cond = torch.randn(nt, 128, 128, 320, generator=g, device="cuda").to(BF)
temb = torch.randn(2 * nt, unet.temb_w.shape[0], generator=g, device="cuda")
eps = torch.empty(2 * nt, 4, 128, 128, device="cuda", dtype=BF)
la = fd.LocalAttention((128, 128), 0.5)
canvas = torch.randn(1, 4, 128, 64 * (nt + 1), generator=g, device="cuda").to(BF)
tiles = la.grids(canvas)
assert tiles.shape[0] == nt, tiles.shape
# algorithmic bytes of every fd_gemm launch (each operand read once + result written once at its storage dtype), for
# roofline.traffic's denominator (host-side bookkeeping only: no extra kernels)
import json
import faithdiff_dev_b200.ops as OP
alg = {"n": 0, "bytes": 0.0, "flops": 0.0}
_g, _c = OP.gemm, OP.conv3x3


def gemm(a, w, a2=None, residual=None, act=0, out_fp32=False, **kw):
    r = _g(a, w, a2=a2, residual=residual, act=act, out_fp32=out_fp32, **kw)
    o = r[0] if isinstance(r, tuple) else r
    alg["n"] += 1
    alg["bytes"] += 2.0 * (a.numel() + (a2.numel() if a2 is not None else 0) + w.numel() + (residual.numel() if residual is not None else 0)) \
        + o.numel() * o.element_size()
    alg["flops"] += 2.0 * a.shape[0] * w.shape[0] * w.shape[1]
    return r


def conv(x, w, x2=None, residual=None, **kw):
    r = _c(x, w, x2=x2, residual=residual, **kw)
    alg["n"] += 1
    alg["bytes"] += 2.0 * (x.numel() + (x2.numel() if x2 is not None else 0) + w.numel() + (residual.numel() if residual is not None else 0) + r.numel())
    alg["flops"] += 2.0 * r.numel() * w.shape[1]
    return r


OP.gemm, OP.conv3x3 = gemm, conv
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(reps):
    ts(scaled, cond, temb, eps)
    ops.cfg_euler_blend(eps.view(nt, 2, 4, 128, 128), tiles, la.weights64, la.origins, 128, 64 * (nt + 1), 5.0, 11.03, 8.39, 8.45)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
json.dump({"launch": "all %d fd_gemm launches of one %d-tile (%d-sample) UNet evaluation, the launch size of the 4K bench" % (alg["n"] // reps, nt, 2 * nt),
           "algorithmic_bytes_per_launch": alg["bytes"] / alg["n"], "algorithmic_flops_per_launch": alg["flops"] / alg["n"],
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of every kernel of scripts/dev_tilestep_eager.py %d "
                     "(profiles/r02_launches_tilestep7_unfiltered.csv; cold-cache, serialised)" % nt},
          open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_gemm_alg.json"), "w"), indent=1)
print("done: %d-tile eager evaluation x%d, fd launches total %d" % (nt, reps, ops.launch_count()))
