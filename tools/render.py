#!/usr/bin/env python
"""Headless render-to-TGA driver on the CUDA path: the replacement of the reference's Win32 main loop for batch output
(main.cpp:56-161 builds the scene, renders, presents; TGAImage::write_tga_file tgaimage.cpp:145-246 writes images).

  python tools/render.py --model african_head --size 800x600 --shader blinn --no-shadow --out out/c1        (configs[0])
  python tools/render.py --model african_head --size 1920x1080 --shader blinn --shadow --frames 1024 --gpus 8 --out out/orbit

--model   a bundled scene name (assets/<name>/<name>.obj) or a path to an OBJ file; textures are found by suffix next to
          it as Model does (model.cpp:45-47: _diffuse.tga, _nm_tangent.tga)
--frames  N frames of the orbit: frame k = the default camera advanced k times by Camera::update_transform with
          orbit = (--orbit-step, 0) (camera.cpp:63-70); 1 frame = the default camera (0,0,2) -> origin
--gpus    frames are dealt to the GPUs in contiguous blocks (one context per GPU in this process, no collective)
--rle     RLE-compressed files (the reference's default, rle = true); encoded on the device for multi-frame renders
Single frames go through hana_draw_model_host + hana_tga_write (the host-buffer boundary), batches through the sweep API
and the device-side encoder. Output: <out>_<frame:04d>.tga (or <out>.tga for one frame). Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

SHADERS = {"shadow": 0, "blinn": 1, "normalmap": 2, "ground": 3, "toon": 4, "texture": 5, "texture_light": 6}


def load_scene(hana, model, normal_pass):
    if os.path.exists(model):
        d, base = os.path.dirname(model), os.path.splitext(os.path.basename(model))[0]
        tex = []
        for suffix in ("_diffuse.tga", "_nm_tangent.tga"):
            p = os.path.join(d, base + suffix)
            tex.append(hana.tga_load(p, model_flip=True) if os.path.exists(p) else None)
        return hana.Scene(base, hana.obj_load(model, normal_pass), tex[0], tex[1])
    return hana.load_bundled(model, None, normal_pass)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model", default="african_head")
    ap.add_argument("--size", default="800x600")
    ap.add_argument("--shader", default="blinn", choices=sorted(SHADERS))
    ap.add_argument("--shadow", dest="shadow", action="store_true", default=True)
    ap.add_argument("--no-shadow", dest="shadow", action="store_false")
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--orbit-step", type=float, default=1.0 / 1024.0)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--rle", action="store_true", default=True)
    ap.add_argument("--raw", dest="rle", action="store_false")
    ap.add_argument("--normal-pass", type=int, default=1, help="which walk of the reference's draw loop the normals correspond to "
                    "(Model::normal re-normalises in place, model.cpp:108-111); 1 = first frame's shadow pass... 3 = after one warm-up frame")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    W, H = (int(x) for x in a.size.lower().split("x"))
    hana = ge.load_package()
    scene = load_scene(hana, a.model, a.normal_pass)
    shader = SHADERS[a.shader]
    os.makedirs(os.path.dirname(os.path.abspath(a.out)) or ".", exist_ok=True)
    t0 = time.perf_counter()
    written = []
    if a.frames == 1:
        ctx = hana.Context(0)
        model, dtex, ntex = scene.upload(ctx)
        u = hana.default_uniforms(W, H, a.shadow)
        col = np.zeros((H, W, 4), np.uint8)
        dep = np.full((H, W), hana.FLT_MAX, np.float32)
        ctx.draw_model_host(col, dep, model, shader, u, dtex, ntex, assume_cleared=True, clear_rgba=(0, 0, 0, 1))
        path = a.out + ".tga"
        hana.tga_write(path, np.ascontiguousarray(col[::-1, :, 2::-1]), rle=a.rle)  # rows top-down, B,G,R (win32.cpp:348-370)
        written.append(path)
        for o in (model, dtex, ntex):
            o.close()
        ctx.close()
    else:
        ngpu = max(1, min(a.gpus, hana.device_count(), a.frames))
        per = (a.frames + ngpu - 1) // ngpu
        ctxs = [hana.Context(g) for g in range(ngpu)]
        jobs = []
        for g, ctx in enumerate(ctxs):  # contiguous block of the orbit per GPU; every GPU holds its own copy of mesh + textures
            first, count = g * per, max(0, min(per, a.frames - g * per))
            if count:
                jobs.append((ctx, scene.upload(ctx), ctx.sweep(W, H, min(a.batch, count)), first, count))
        turn = int(round(1.0 / a.orbit_step)) if a.orbit_step > 0 else 1024
        done = [0] * len(jobs)
        while any(d < j[4] for d, j in zip(done, jobs)):
            live = []
            for k, (ctx, objs, sw, first, count) in enumerate(jobs):  # queue one batch on every GPU, then collect
                n = min(sw.max_frames, count - done[k])
                if n > 0:
                    arr = hana.orbit_sweep_uniforms(W, H, first + done[k], n, frames_per_turn=turn, enable_shadow=a.shadow)
                    sw.render(objs[0], shader, arr, objs[1], objs[2])
                    live.append((k, n))
            for k, n in live:
                ctx, objs, sw, first, count = jobs[k]
                if a.rle:
                    files = sw.tga_files(0, n)  # encoded on the device
                    for i, data in enumerate(files):
                        path = "%s_%04d.tga" % (a.out, first + done[k] + i)
                        open(path, "wb").write(data)
                        written.append(path)
                else:
                    surf = sw.present(0, n, hana.PRESENT_BGR8)
                    for i in range(n):
                        path = "%s_%04d.tga" % (a.out, first + done[k] + i)
                        hana.tga_write(path, surf[i], rle=False)
                        written.append(path)
                done[k] += n
        for ctx, objs, sw, _, _ in jobs:
            for o in (sw,) + tuple(objs):
                o.close()
        for ctx in ctxs:
            ctx.close()
    dt = time.perf_counter() - t0
    print(json.dumps({"frames": a.frames, "size": [W, H], "model": scene.name, "faces": scene.nfaces, "shader": a.shader,
                      "shadow": a.shadow, "rle": a.rle, "files": len(written), "first_file": written[0] if written else None,
                      "seconds": dt, "bytes": sum(os.path.getsize(p) for p in written)}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
