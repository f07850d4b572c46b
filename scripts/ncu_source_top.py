"""Per-source-line instruction counts from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python scripts/ncu_source_top.py in.csv [out.csv] [topN]"""
import csv, sys, collections, os
src = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
agg = collections.defaultdict(lambda: [0, 0, 0, ''])
cur = None; hdr = None
with open(src, newline='') as f:
    for row in csv.reader(f):
        if not row: continue
        if row[0] == 'File Path': cur = os.path.basename(row[1]); hdr = None; continue
        if row[0] == 'Function Name': continue
        if row[0] == 'Line No': hdr = row; continue
        if hdr is None or len(row) < len(hdr): continue
        line, text = row[0], row[1]
        if not line.isdigit(): continue      # SASS rows repeat what their source line already sums
        ie = hdr.index('Instructions Executed'); te = hdr.index('Thread Instructions Executed'); ss = hdr.index('# Samples')
        def num(x):
            try: return int(float(x.replace(',', '')))
            except Exception: return 0
        a = agg[(cur, int(line) if line.isdigit() else -1)]
        a[0] += num(row[ie]); a[1] += num(row[te]); a[2] += num(row[ss])
        if text: a[3] = text.strip()
tot = sum(a[0] for a in agg.values()) or 1
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]
w = csv.writer(open(out, 'w', newline='') if out else sys.stdout)
w.writerow(['inst_executed', 'share_pct', 'thread_inst', 'stall_samples', 'file', 'line', 'source'])
for (fn, ln), a in rows:
    w.writerow([a[0], round(100.0 * a[0] / tot, 2), a[1], a[2], fn, ln, a[3][:150]])
print('total warp instructions', tot, file=sys.stderr)
