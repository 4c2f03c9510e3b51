#!/usr/bin/env python
"""Per-source-line hot spots of one ncu capture (needs -lineinfo and --import-source on).
    python tools/ncu_source_hot.py <report.ncu-rep> [top]
Prints, per (file, line): warp instructions executed, share, stall samples -- the `cuda,sass` source page aggregated."""
import csv, io, subprocess, sys, os, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = []
for ln in out.splitlines():      # ncu does not escape quotes inside the source column: split by hand, merge the surplus into it
    if not ln.startswith('"'):
        continue
    rows.append(ln.strip().strip('"').split('","'))
cur = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(r)}; continue
    if r[0] == "": continue   # SASS rows
    try:
        line = int(r[0])
    except ValueError:
        continue
    if len(r) > len(hdr):
        extra = len(r) - len(hdr)
        r = [r[0], '","'.join(r[1:2 + extra])] + r[2 + extra:]
    if len(r) != len(hdr):
        continue
    inst = int(r[ix["Instructions Executed"]] or 0); samp = int(r[ix["# Samples"]] or 0)
    k = (cur, line)
    a = agg.setdefault(k, [0, 0, r[1]])
    a[0] += inst; a[1] += samp
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print(f"total warp instructions {tot:,}  stall samples {tots:,}")
byfile = collections.Counter(); byfile_s = collections.Counter()
for (f, l), a in agg.items(): byfile[f] += a[0]; byfile_s[f] += a[1]
for f, v in byfile.most_common(): print(f"  {f:28s} {100*v/tot:5.1f}% inst  {100*byfile_s[f]/max(tots,1):5.1f}% samples")
print(f"{'file:line':32s} {'inst%':>6s} {'samp%':>6s}  source")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f+':'+str(l):32s} {100*a[0]/tot:6.2f} {100*a[1]/max(tots,1):6.2f}  {a[2].strip()[:110]}")
