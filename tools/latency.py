#!/usr/bin/env python
"""Single-frame latency through the drop-in boundary, measured separately and honestly (SURVEY.md §7 hard part 2): one
frame of the bundled scenes is launch- and copy-latency-bound, nowhere near the batched throughput the bench reports.

  Level 1  the reference's own scene code (DrawModel::draw, scene.h:53-99, compiled from its unmodified sources) linked with
           this repo's graphics_draw_triangle(DrawData*) shim instead of graphics.cpp (oracle/_ref/libhana_ref_dropin.so):
           per pass the shim gathers the a2v stream, uploads the host RenderBuffer (+ the shadow map's colour plane),
           draws, downloads the RenderBuffer — the price of keeping the reference's host buffers the source of truth.
  Level 2  ONE hana_draw_model_host call per frame (uniform upload, both passes, colour + depth back to host memory).
  ref      the unmodified reference on one host core, same call (DrawModel::draw), for scale.

configs[0] = african_head, Blinn, 800x600, shadow off; configs[1] = same, shadow on, 1920x1080; plus the README workload
(diablo3_pose, NormalMap + shadow, 1000x600). Writes profiles/r02_latency.json; prints the same JSON.
This tool drives the reference's compiled scene code, which is test infrastructure: it is not part of bench.py."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def stats(ms):
    ms = sorted(ms)
    return {"median_ms": ms[len(ms) // 2], "min_ms": ms[0], "p90_ms": ms[int(len(ms) * 0.9)], "n": len(ms)}


def main():
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 30
    hana = ge.load_package()
    from oracle import horacle as H
    assets = os.path.join(ROOT, "assets")
    cases = [("configs[0] african_head Blinn 800x600 shadow off", "african_head", H.BLINN, hana.BLINN, 800, 600, False),
             ("configs[1] african_head Blinn 1920x1080 shadow on", "african_head", H.BLINN, hana.BLINN, 1920, 1080, True),
             ("README diablo3_pose NormalMap 1000x600 shadow on", "diablo3_pose", H.NORMALMAP, hana.NORMALMAP, 1000, 600, True)]
    out = {"unit": "ms per frame, wall clock of one synchronous call", "cases": []}
    ctx = hana.Context(0)
    for name, scene, hshader, shader, W, Hh, shadow in cases:
        obj = os.path.join(assets, scene, scene + ".obj")
        row = {"case": name}
        # Level 1: reference scene code + the shim
        if os.path.exists(H.REF_DROPIN_SO):
            g = H.Reference(obj, W, Hh, hshader, dropin=True)
            for _ in range(3):
                g.render_time(shadow)
            row["level1_shim"] = stats([g.render_time(shadow) * 1e3 for _ in range(reps)])
            g.close()
        # the reference itself, one core
        if os.path.exists(H.REF_SO):
            r = H.Reference(obj, W, Hh, hshader, instrumented=False)
            r.warmup(shadow)
            row["reference_one_core"] = stats([r.render_time(shadow) * 1e3 for _ in range(max(3, reps // 10))])
            r.close()
        # Level 2: one C-ABI call with host buffers
        sc = hana.load_bundled(scene, assets, 3)
        objs = sc.upload(ctx)
        u = hana.default_uniforms(W, Hh, shadow)
        col = np.zeros((Hh, W, 4), np.uint8)
        dep = np.full((Hh, W), hana.FLT_MAX, np.float32)
        ts = []
        for i in range(reps + 3):
            t0 = time.perf_counter()
            ctx.draw_model_host(col, dep, objs[0], shader, u, objs[1], objs[2], assume_cleared=True)
            ts.append((time.perf_counter() - t0) * 1e3)
        row["level2_hana_draw_model_host"] = stats(ts[3:])
        for o in objs:
            o.close()
        out["cases"].append(row)
    ctx.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r02_latency.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
