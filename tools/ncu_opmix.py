#!/usr/bin/env python
"""Dynamic opcode mix of one kernel from an ncu report (--set full --import-source on): warp-level
instructions executed per SASS opcode, and per source line range.

  python tools/ncu_opmix.py gpurun_out/prof.ncu-rep 'raster_kernel<(int)1' [--top 30]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
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
        isrc, ii = hdr.index("Source"), hdr.index("Instructions Executed")
        agg, total = {}, 0
        for r in rs[1:]:
            try:
                n = int(r[ii])
            except (ValueError, IndexError):
                continue
            s = r[isrc].strip()
            toks = s.split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            op = op.split(".")[0]
            agg[op] = agg.get(op, 0) + n
            total += n
        print("== %s\n   warp instructions %d" % (name[:90], total))
        for op, n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
            print("  %6.2f%%  %12d  %s" % (100.0 * n / max(total, 1), n, op))


if __name__ == "__main__":
    main()
