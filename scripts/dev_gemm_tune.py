"""Dev (GPU): sweep (CTA pair?, N tile) for every distinct GEMM/conv shape of one UNet tile-step; prints the measured time of
each option next to the launcher's own choice.  usage: dev_gemm_tune.py [tiles_per_launch]"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import faithdiff_dev_b200 as fd
from faithdiff_dev_b200 import ops, arch
import faithdiff_dev_b200.ops as O
from faithdiff_dev_b200.pipeline import TileStepGraph
BF = torch.bfloat16
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = arch.SDXL_UNET_CONFIG
unet = fd.UNet2DConditionModel(**cfg); unet.load_state_dict(arch.random_state_dict(arch.unet_param_shapes(cfg)))
ehs = torch.randn(2, 77, 2048, device="cuda").to(BF)
kv, lk = unet.text_kv(ehs)
ts = TileStepGraph(unet, kv, lk, [i // 2 for i in range(2 * nb)], [i % 2 for i in range(2 * nb)], 128, 128, BF, True, use_graph=False)
ts.scaled.normal_(); ts.cond.normal_(); ts.temb.normal_()
calls = collections.OrderedDict()
count = collections.Counter()
real = {}
def wrap(name):
    f = getattr(O, name); real[name] = f
    def g(*a, **k):
        r = f(*a, **k)
        shp = tuple(tuple(v.shape) for v in a[:2] if torch.is_tensor(v))
        key = (name, shp, k.get("act", 0), k.get("stride", 1), k.get("residual") is not None, k.get("x2", k.get("a2")) is not None)
        count[key] += 1
        if key not in calls:
            kk = dict(k); kk["out"] = r
            calls[key] = (a, kk)
        return r
    setattr(O, name, g)
wrap("gemm"); wrap("conv3x3")
ts._run()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=8):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()                       # evict L2: in the pipeline the weights always come from HBM
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps * 1e3
total_auto = total_best = 0.0
for key, (a, k) in calls.items():
    name = key[0]
    f = real[name]
    auto = timeit(lambda: f(*a, **k))
    res = {}
    for pair in (0, 1):
        for bn in range(64, 257, 32):
            if key[2] == 2 and bn % 32: continue
            kk = dict(k); kk["block_n"] = bn | (0x1000 if pair else 0x2000)
            try:
                res[(pair, bn)] = timeit(lambda: f(*a, **kk), 5)
            except Exception as ex:
                res[(pair, bn)] = float("inf")
    best = min(res, key=res.get)
    n = count[key]
    total_auto += auto * n; total_best += res[best] * n
    print("%-8s %-44s act=%d s=%d res=%d two=%d x%3d auto %7.1f us | best pair=%d bn=%3d %7.1f us | " % (
        name, str(key[1]), key[2], key[3], key[4], key[5], n, auto, best[0], best[1], res[best]) +
        " ".join("%d/%d:%.0f" % (p_, b_, t) for (p_, b_), t in sorted(res.items())), flush=True)
print("sum over launches: auto %.2f ms, best %.2f ms" % (total_auto / 1e3, total_best / 1e3))
