"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by source line and by code region.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda > both.csv; python scripts/ncu_source_regions.py both.csv file.cuh name:a-b ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
fname = sys.argv[2]
regions = [(a.split(':')[0], *map(int, a.split(':')[1].split('-'))) for a in sys.argv[3:]]
cur = None; hdr = None; agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 2 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        try: samples, inst = int(r[4] or 0), int(r[7] or 0)
        except ValueError: continue
        st = {hdr[i]: int(r[i] or 0) for i in range(len(hdr)) if hdr[i].startswith('stall_') and 'Not Issued' not in hdr[i]}
        agg[(cur, int(r[0]))] = (samples, inst, st, r[1])
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print("total samples", tot, "warp instructions", toti)
for name, a, b in regions:
    sel = [v for (f, l), v in agg.items() if f == fname and a <= l <= b]
    s = sum(v[0] for v in sel); i = sum(v[1] for v in sel); st = {}
    for v in sel:
        for k, x in v[2].items(): st[k] = st.get(k, 0) + x
    top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
    print(f"{name:16s} samples {100*s/tot:5.1f}%  inst {100*i/toti:5.1f}%  ", [(k[6:], round(100 * x / max(s, 1))) for k, x in top])
oth = [v for (f, l), v in agg.items() if f != fname]
print("other files: samples %.1f%% inst %.1f%%" % (100 * sum(v[0] for v in oth) / tot, 100 * sum(v[1] for v in oth) / toti))
print("top lines:")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(__import__('os').environ.get('TOP', 30))]:
    top = sorted(v[2].items(), key=lambda kv: -kv[1])[:3]
    print(f"{f}:{l:4d} s {100*v[0]/tot:4.1f}% i {100*v[1]/toti:4.1f}% {[(k[6:], x) for k, x in top]} | {v[3].strip()[:80]}")
