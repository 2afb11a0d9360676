#!/usr/bin/env python
"""Prints the per-launch table of an `ncu --csv --metrics ...` log (time, instructions, DRAM bytes, issue utilisation)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iM, iV, iID = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = {}
for r in rows[1:]:
    d.setdefault((int(r[iID]), r[iK].split('(')[0]), {})[r[iM]] = float(r[iV].replace(',', ''))
for k in sorted(d):
    m = d[k]
    print(k[0], k[1][:44], "%.1f us" % (m['gpu__time_duration.sum'] / 1e3), "inst %.0fk" % (m.get('smsp__inst_executed.sum', 0) / 1e3),
          "rd %.2f MB wr %.2f MB" % (m.get('dram__bytes_read.sum', 0) / 1e6, m.get('dram__bytes_write.sum', 0) / 1e6),
          "issue %.0f%%" % m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0))
