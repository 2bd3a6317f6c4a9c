"""Segment the stall samples of an `ncu --page source --csv` dump of a kernel by its barrier / mbarrier instructions:
python tools/ncu_segments.py dump.csv   (how the per-pass shares in profiles/*_summary.md are obtained)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
acc, accst = 0, {s: 0 for s in stalls}
allst = {s: 0 for s in stalls}
for k, r in enumerate(data):
    src = r[ix['Source']].strip()
    acc += int(r[ix['# Samples']])
    for s in stalls:
        v = int(r[ix[s]] or 0)
        accst[s] += v
        allst[s] += v
    if any(t in src for t in ('BAR.', 'SYNCS.PHASECHK', 'UBLKCP', 'EXIT')) and acc > 0:
        top = sorted(accst.items(), key=lambda kv: -kv[1])[:5]
        print(f"{k:5d} {src[:52]:52s} {acc:5d} ({100 * acc / tot:4.1f}%) " + ' '.join(f"{a[6:]}={b}" for a, b in top))
        acc, accst = 0, {s: 0 for s in stalls}
print('all:', ' '.join(f"{a[6:]}={100 * b / max(1, sum(allst.values())):.0f}%" for a, b in sorted(allst.items(), key=lambda kv: -kv[1])[:9]))
