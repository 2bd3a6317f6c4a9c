"""Instruction mix + top stall lines from `ncu --page source --csv` output of a report."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
iS, iN, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
cnt, tot = collections.Counter(), 0
lines = []
for r in rows[hi + 1:]:
    if len(r) <= iN or not r[iN].isdigit():
        continue
    n = int(r[iN]); toks = r[iS].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    cnt[op.split(".")[0]] += n; tot += n
    lines.append((int(r[iSamp] or 0), n, r[iS].strip()))
print("total warp instructions", tot)
for k, v in cnt.most_common(22):
    print(f"  {k:10s} {v:10d} {100*v/tot:5.1f}%")
print("top sampled instructions:")
for s, n, src in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 15]:
    print(f"  {s:5d} {n:9d} {src}")
