#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel from an ncu report.

ncu's CSV source page is per SASS instruction; this joins it with `nvdisasm -g` line markers of the
shipped library so the hot spots can be read against hana_kernels.cuh / hana_core.cuh.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep 'raster_kernel<(int)0' [--top 40]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hana-softwarerenderer_b200", "libhana_b200.so")


def line_map():
    """{mangled kernel: {address: 'file:line'}} from nvdisasm -g of the library's cubin."""
    tmp = tempfile.mkdtemp(prefix="hana_cub")
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    maps = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                             text=True).stdout
        cur, loc = None, "?"
        for ln in txt.splitlines():
            m = re.search(r"\.text\.(\S+)\s+-+", ln)
            if m:
                cur = maps.setdefault(m.group(1), {})
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                loc = "%s:%s" % (os.path.basename(m.group(1)), m.group(2))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
            if m and cur is not None:
                cur[int(m.group(1), 16)] = loc
    return maps


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = {}, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = blocks.setdefault(r[1], [])
        elif cur is not None:
            cur.append(r)
    maps = line_map()
    for name, rs in blocks.items():
        if pat not in name:
            continue
        # demangled -> mangled: match on template args
        t = re.search(r"raster_kernel<\(int\)(\d+), \(int\)(\d+)>", name)
        key = None
        for k in maps:
            if t and ("raster_kernelILi%sELi%sE" % t.groups()) in k:
                key = k
            elif not t and name.split("(")[0].split("::")[-1].split("<")[0] in k:
                key = key or k
        amap = maps.get(key, {})
        hdr = rs[0]
        ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        base = None
        agg = {}
        for r in rs[1:]:
            try:
                addr = int(r[ia], 16)
            except (ValueError, IndexError):
                continue
            base = addr if base is None else base
            loc = amap.get(addr - base, "?")
            a = agg.setdefault(loc, [0, 0, {}])
            a[0] += int(r[ii] or 0)
            a[1] += int(r[isamp] or 0)
            for i, h in stall_cols:
                v = int(r[i] or 0)
                if v:
                    a[2][h] = a[2].get(h, 0) + v
        ti = sum(a[0] for a in agg.values()) or 1
        ts = sum(a[1] for a in agg.values()) or 1
        print("== %s\n   warp instructions %d, samples %d" % (name[:90], ti, ts))
        for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0 if "--by-inst" in sys.argv else 1])[:top]:
            st = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
            print("  %5.1f%% inst  %5.1f%% samples  %-24s %s" % (100.0 * a[0] / ti, 100.0 * a[1] / ts, loc,
                                                              " ".join("%s=%d" % (h[6:], v) for h, v in st)))


if __name__ == "__main__":
    main()
