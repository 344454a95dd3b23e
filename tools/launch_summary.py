import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
tot = 0; agg = {}
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    n = r[ki].split('(')[0]
    agg.setdefault(n, []).append(v); tot += v
for n, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
    if sum(v) / tot < 0.002: continue
    print(f"{n:40s} n={len(v):3d} sum={sum(v)/1e6:9.3f} ms share={sum(v)/tot:.3f} each={[round(x/1e6,2) for x in v][:10]}")
