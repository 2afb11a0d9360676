#!/usr/bin/env python
"""SASS listing of one kernel from an ncu report with per-instruction executed counts (address order), optionally only
instructions executed at least --min times.  python tools/ncu_sass.py rep 'raster_kernel<(int)0' --min 1000000"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
mn = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = {}, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = blocks.setdefault(r[1], [])
    elif cur is not None:
        cur.append(r)
for name, rs in blocks.items():
    if pat not in name:
        continue
    hdr = rs[0]
    ia, isrc, ii = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    isamp = hdr.index("# Samples")
    print("==", name[:100])
    for r in rs[1:]:
        try:
            n = int(r[ii])
        except (ValueError, IndexError):
            continue
        if n >= mn:
            print("%6s %11d %6s  %s" % (r[ia][-5:], n, r[isamp], r[isrc].strip()[:110]))
