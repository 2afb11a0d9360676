#!/usr/bin/env python
"""One markdown table row per kernel launch of an ncu report (--set full): duration, DRAM traffic, instructions,
issue utilisation, occupancy, registers. The source of the tables under profiles/.

  python tools/ncu_summary.py gpurun_out/prof_all.ncu-rep > profiles/rNN_step_kernels.md
  python tools/ncu_summary.py gpurun_out/prof_all.ncu-rep --traffic 256 profiles/raster_main_dram.json
      also writes the DRAM traffic (read + write) of the raster_main launch, per frame (the capture rendered 256 frames
      per launch), to the file bench.py reads `roofline.traffic` from
"""
import csv
import io
import json
import os
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    idx = [(hdr.index(c), t) for c, t in COLS if c in hdr]
    print("| # | kernel | " + " | ".join(t for _, t in idx) + " |")
    print("|---|---|" + "---|" * len(idx))
    for k, r in enumerate(data):
        cells = []
        for i, _ in idx:
            v = r[i]
            try:
                f = float(v)
                v = ("%.0f" % f) if f >= 1000 or f == int(f) else ("%.2f" % f)
            except ValueError:
                pass
            u = units[i]
            cells.append(v + ((" " + u) if u and u not in ("%", "inst", "register/thread", "") else ""))
        print("| %d | `%s` | %s |" % (k, r[name_i].split("(")[0][:48], " | ".join(cells)))


    if "--traffic" in sys.argv:
        k = sys.argv.index("--traffic")
        frames, dst = int(sys.argv[k + 1]), sys.argv[k + 2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        best = None
        for r in data:
            if "raster_kernel<1, 0" in r[name_i] or "raster_kernel<(int)1, (int)0" in r[name_i]:
                best = r  # the last such launch (steady state)
        if best is None:
            raise SystemExit("no raster_kernel<BLINN, CLEAR_FOLD> launch in the report")
        rd, wr = float(best[ri]) * scale[units[ri]], float(best[wi]) * scale[units[wi]]
        json.dump({"kernel": "raster_kernel<BLINN, CLEAR_FOLD>", "dram_read_bytes_per_launch": rd, "dram_write_bytes_per_launch": wr,
                   "frames_per_launch": frames, "bytes_per_frame": (rd + wr) / frames,
                   "source": "ncu --set full capture %s (dram__bytes_read.sum + dram__bytes_write.sum of one launch of %d frames)" %
                             (os.path.basename(rep), frames)}, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main()
