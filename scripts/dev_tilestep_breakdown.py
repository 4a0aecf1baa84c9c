"""Dev (GPU): per-op, per-shape breakdown of one eager 7-tile UNet evaluation (CUDA events around every ops.* call while a spin
kernel keeps the GPU behind the host): where the tile-step's time is and which shapes run below par.  usage: [tiles=7]"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import faithdiff_dev_b200 as fd
from faithdiff_dev_b200 import arch, ops
from faithdiff_dev_b200.pipeline import TileStepGraph
BF = torch.bfloat16
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 7
unet = fd.UNet2DConditionModel(**arch.SDXL_UNET_CONFIG)
unet.load_state_dict(arch.random_state_dict(arch.unet_param_shapes(arch.SDXL_UNET_CONFIG), seed=1234))
g = torch.Generator(device="cuda").manual_seed(0)
kv, lk = unet.text_kv(torch.randn(2, 77, 2048, generator=g, device="cuda").to(BF))
ts = TileStepGraph(unet, kv, lk, tuple(i // 2 for i in range(2 * nt)), tuple(i % 2 for i in range(2 * nt)), 128, 128, BF, True, use_graph=False)
rec = collections.defaultdict(list)


def wrap(name, flops):
    f = getattr(ops, name)

    def w(*a, **k):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); r = f(*a, **k); e.record()
        o = r[0] if isinstance(r, tuple) else r
        key = (name,) + tuple(tuple(t.shape) for t in a[:2] if torch.is_tensor(t)) + (k.get("act", 0), k.get("stride", 1), "res" if k.get("residual") is not None else "",
                                                                                       "ln" if k.get("ln_in") is not None else ("st" if k.get("stats_out") else ""))
        rec[key].append((s, e, flops(a, k, o)))
        return r
    setattr(ops, name, w)


wrap("gemm", lambda a, k, o: 2.0 * a[0].shape[0] * a[1].shape[0] * a[1].shape[1])
wrap("conv3x3", lambda a, k, o: 2.0 * o.numel() * a[1].shape[1])
wrap("conv3x3_upsampled2x", lambda a, k, o: 2.0 * o.numel() * a[1][0].shape[1])
wrap("attention", lambda a, k, o: 4.0 * a[3] * a[4] * a[5] * a[6] * 64)
for n in ("groupnorm", "layernorm", "conv3x3_small_in", "conv3x3_small_out", "batch_axis_attention", "nchw_to_nhwc"):
    wrap(n, lambda a, k, o: 0.0)
torch.cuda._sleep(int(3e8))
ts._run()
torch.cuda.synchronize()
rows = []
for key, evs in rec.items():
    t = sum(s.elapsed_time(e) for s, e, _ in evs)
    rows.append((t, len(evs), key, sum(f for _, _, f in evs)))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("%d-tile evaluation: %.2f ms in %d timed calls" % (nt, tot, sum(r[1] for r in rows)))
for t, n, key, fl in rows[:45]:
    print("  %7.3f ms %5.1f%% x%3d %7.1f us %s %s" % (t, 100 * t / tot, n, t / n * 1e3, ("%6.0f TF/s" % (fl / t / 1e9)) if fl else "           ", str(key)[:150]))
