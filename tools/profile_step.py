#!/usr/bin/env python
"""One profiled step of the bench workload for ncu: 3 warm-up submissions, then ONE submission of --frames frames
(african_head, Shadow + Blinn, 1080p orbit) between cudaProfilerStart/Stop.

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/rNN_step \
      python tools/profile_step.py --frames 256
Read the report on the CPU box with tools/ncu_summary.py / ncu_lines.py / ncu_opmix.py."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--scene", default="african_head")
    ap.add_argument("--shader", default="BLINN")
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--tga", action="store_true", help="also encode the frames as RLE TGA files (hana_sweep_encode_tga) inside the profiled region")
    a = ap.parse_args()
    import torch
    hana = ge.load_package()
    W, H = (int(x) for x in a.size.split("x"))
    ctx = hana.Context(0)
    sc = hana.load_bundled(a.scene, None, 3)
    model, dtex, ntex = sc.upload(ctx)
    sw = ctx.sweep(W, H, a.frames)
    arr = hana.orbit_sweep_uniforms(W, H, 0, a.frames, frames_per_turn=1024)
    shader = getattr(hana, a.shader)
    for _ in range(3):
        sw.render(model, shader, arr, dtex, ntex)
        if a.tga:
            assert ctx.L.hana_sweep_encode_tga(sw.h, 0, a.frames) == 0
        ctx.sync()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    sw.render(model, shader, arr, dtex, ntex)
    if a.tga:
        assert ctx.L.hana_sweep_encode_tga(sw.h, 0, a.frames) == 0
    ctx.sync()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    for o in (sw, model, dtex, ntex):
        o.close()
    ctx.close()


if __name__ == "__main__":
    main()
