"""Dev (GPU): accurate per-shape GEMM/conv/attention timings of one UNet evaluation (events + spin head start)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import faithdiff_dev_b200 as fd
from faithdiff_dev_b200 import ops, arch
import faithdiff_dev_b200.ops as O
from faithdiff_dev_b200.pipeline import TileStepGraph
BF = torch.bfloat16
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = arch.SDXL_UNET_CONFIG
unet = fd.UNet2DConditionModel(**cfg); unet.load_state_dict(arch.random_state_dict(arch.unet_param_shapes(cfg)))
ehs = torch.randn(2, 77, 2048, device="cuda").to(BF)
kv, lk = unet.text_kv(ehs)
ns = int(os.environ.get("FD_SAMPLES", 2 * nb))        # odd sample counts occur in the multi-GPU unit split
ts = TileStepGraph(unet, kv, lk, [i // 2 for i in range(ns)], [i % 2 for i in range(ns)], 128, 128, BF, True, use_graph=True)
ts.scaled.normal_(); ts.cond.normal_(); ts.temb.normal_()
for _ in range(3): ts.graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10): ts.graph.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
nb = ns / 2
print("graph replay samples=%d: %.2f ms per replay = %.2f ms per tile (%.0f TF/s ref-alg)" % (ns, ms, ms / nb, 14.216 * nb / ms * 1e3), flush=True)
if os.environ.get("FD_REPLAY_ONLY"): sys.exit(0)
rec = collections.defaultdict(list)
order = []
def wrap(name, flop):
    f = getattr(O, name)
    def g(*a, **k):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); r = f(*a, **k); e.record()
        shp = tuple(tuple(v.shape) for v in a[:2] if torch.is_tensor(v))
        rec[(name, shp, k.get("act", 0), k.get("stride", 1), k.get("residual") is not None)].append((s, e, flop(a, k, r)))
        order.append((name, shp, k.get('residual') is not None, (k.get('bias') is not None) or (len(a) > 2 and torch.is_tensor(a[2])), s, e))
        return r
    setattr(O, name, g)
wrap("gemm", lambda a, k, r: 2.0 * a[0].shape[0] * a[1].shape[0] * a[1].shape[1])
wrap("conv3x3", lambda a, k, r: 2.0 * r.numel() * a[1].shape[1])
wrap("attention", lambda a, k, r: 4.0 * a[3] * a[4] * a[5] * a[6] * 64)
wrap("layernorm", lambda a, k, r: 0)
wrap("groupnorm", lambda a, k, r: 0)
torch.cuda._sleep(int(4e8))
ts._run()
torch.cuda.synchronize()
rows = []
for key, evs in rec.items():
    t = sum(s.elapsed_time(e) for s, e, _ in evs); f = sum(x[2] for x in evs)
    rows.append((t, len(evs), key, f))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
cls = collections.defaultdict(float)
for t, n, k, f in rows: cls[k[0]] += t
print("instrumented total %.2f ms: " % tot + ", ".join("%s %.2f" % kv for kv in sorted(cls.items(), key=lambda kv: -kv[1])))
for t, n, k, f in rows[:45]:
    print("  %7.3f ms x%3d (%6.1f us each) %-9s %-46s act=%d s=%d res=%d %s" % (t, n, t / n * 1e3, k[0], str(k[1]), k[2], k[3], k[4], ("%.0f TF/s" % (f / t / 1e9)) if f else ""))

if os.environ.get("FD_SEQ"):
    lo, hi = [int(v) for v in os.environ["FD_SEQ"].split(",")]
    for i, (name, shp, res, bias, s_, e_) in enumerate(order[lo:hi]):
        print("  #%4d %-9s %-44s res=%d bias=%d %7.1f us" % (lo + i, name, str(shp), res, bias, s_.elapsed_time(e_) * 1e3))
