import csv, collections, re, sys
with open(sys.argv[1]) as f:
    lines=[l for l in f if not l.startswith('==')]
rows=list(csv.DictReader(lines))
seq=[]
for row in rows:
    n=row['Kernel Name']; v=float(row['Metric Value'].replace(',',''))
    unit=row['Metric Unit']
    if unit.startswith('n'): v/=1e3
    elif unit.startswith('m'): v*=1e3
    seq.append((re.sub(r'\(.*','',n),v,row['Grid Size']))
# stage split: encode = up to first scale_div; denoise = until first conv1x1 after tile_blend
names=[s[0] for s in seq]
def agg(lo,hi,title):
    tot=collections.defaultdict(lambda:[0,0.0]); S=sum(s[1] for s in seq[lo:hi])
    for n,v,g in seq[lo:hi]: tot[n][0]+=1; tot[n][1]+=v
    print("== %s: %d launches, %.1f us"%(title,hi-lo,S))
    for n,(c,t) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:14]:
        print("   %-32s x%5d %10.1f us %5.1f%%"%(n[:32],c,t,100*t/S))
i0=names.index('scale_div_kernel') if 'scale_div_kernel' in names else 0
i1=names.index('tile_blend_kernel', i0)+1 if 'tile_blend_kernel' in names[i0:] else len(seq)
agg(0,i0,'before the first scheduler step (LQ encoder, VAE encode, condition embedding, eager warm-up of the tile-step before graph capture)'); agg(i0,i1,'one denoise step (1 tile: one UNet tile-step + fused CFG/Euler/blend tail)'); agg(i1,len(seq),'rest of the capture window (further denoise steps up to the ncu -c cap)')
# gemm by grid size buckets in denoise
g=[(v,gs) for n,v,gs in seq[i0:i1] if n=='gemm_tcgen05_kernel']
print("gemm launches in denoise:",len(g),"mean us %.1f"%(sum(v for v,_ in g)/max(1,len(g))))
h=collections.Counter()
for v,_ in g: h[int(v//5)*5]+=1
print(sorted(h.items()))
