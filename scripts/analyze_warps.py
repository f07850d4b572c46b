import sys, csv
rows=[list(map(int,r)) for r in csv.reader(open(sys.argv[1]))]
t0=min(r[1] for r in rows); t1=max(r[2] for r in rows)
ends=sorted(r[2]-t0 for r in rows); starts=sorted(r[1]-t0 for r in rows)
n=len(rows)
print(f"warps {n} span {(t1-t0)/1e3:.1f} us; start p50 {starts[n//2]/1e3:.1f} p99 {starts[int(n*.99)]/1e3:.1f} max {starts[-1]/1e3:.1f}")
print("end percentiles us:", {p: round(ends[min(n-1,int(n*p/100))]/1e3,1) for p in (1,10,50,90,99,100)})
busy=sum(r[2]-r[1] for r in rows)/n
print(f"mean busy {busy/1e3:.1f} us = {100*busy/(t1-t0):.1f}% of span; items/warp mean {sum(r[3] for r in rows)/n:.1f} max {max(r[3] for r in rows)}; levelB misses total {sum(r[4] for r in rows)}")
late=[r for r in rows if r[2]-t0 > 0.9*(t1-t0)]
print("warps ending in last 10% of span:", len(late), "their misses:", sum(r[4] for r in late), "items:", sum(r[3] for r in late))
