"""Aggregates an ncu --csv launch list that carries gpu__time_duration.sum + dram__bytes_read.sum + dram__bytes_write.sum per launch
(scripts/gpu_profile_r02.sh) into (a) the unfiltered per-kernel share table and (b) profiles/r02_gemm_traffic.json for bench.py.
usage: agg_traffic.py <launches.csv> <algorithmic_bytes_json_or_-> [out.json]"""
import collections, csv, json, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per = collections.OrderedDict()
for r in rows:
    key = r["ID"]
    d = per.setdefault(key, {"name": re.sub(r"\(.*", "", r["Kernel Name"])})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"].lower()
    m = r["Metric Name"]
    if m.startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u if u in ("ns", "us", "ms", "s") else ("ns" if u.startswith("n") else ("us" if u.startswith("u") else "ms")), 1.0)
        d["us"] = v
    else:
        v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
        d["rd" if "read" in m else "wr"] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in per.values():
    a = agg[d["name"]]
    a[0] += 1; a[1] += d.get("us", 0.0); a[2] += d.get("rd", 0.0) + d.get("wr", 0.0)
S = sum(a[1] for a in agg.values())
ours = sum(a[1] for k, a in agg.items() if k.startswith("fd::") or "fd::" in k)
print("%d launches, %.1f us kernel time (cold-cache, serialised under ncu); fd:: kernels %.2f %%, other (torch glue) %.2f %%" % (len(per), S, 100 * ours / S, 100 * (S - ours) / S))
for k, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %6.2f %%  x%5d  %10.1f us  DRAM %9.1f MB  %s" % (100 * t / S, c, t, b / 1e6, k[:110]))
g = [0, 0.0, 0.0]
for k, a in agg.items():
    if "gemm_tcgen05" in k:
        g = [g[0] + a[0], g[1] + a[1], g[2] + a[2]]
if g[0] and len(sys.argv) > 3:
    alg = json.load(open(sys.argv[2])) if sys.argv[2] != "-" else {}
    out = {"kernel": "gemm_tcgen05_kernel (all instantiations)", "launches": g[0], "dram_bytes_per_launch": g[2] / g[0],
           "avg_launch_us_under_ncu": g[1] / g[0], "share_of_kernel_time": g[1] / S}
    out.update(alg)
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print("wrote", sys.argv[3])
