#!/usr/bin/env python
"""Times BASELINE.json configs[3] (10 M tiny triangles @ 3840x2160, Blinn + shadow) and configs[4] (9 216 large
triangles, depth complexity ~8 @ 7680x4320, NormalMap + shadow) at full size on one GPU: single-frame sweeps,
CUDA-event time per kernel class. Informational (the bench line is configs[1]/[2]); parity of these configs is in
tests/test_configs.py.  python tools/time_configs.py [--reps 5]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def run(hana, ctx, name, sc, shader, W, Hh, reps):
    objs = sc.upload(ctx)
    u = hana.default_uniforms(W, Hh, True)
    sw = ctx.sweep(W, Hh, 1)
    for _ in range(2):
        sw.render(objs[0], shader, [u], objs[1], objs[2])
    ctx.sync()
    ctx.profile(True, reset=True)
    ctx.timer_start()
    for _ in range(reps):
        sw.render(objs[0], shader, [u], objs[1], objs[2])
    ms = ctx.timer_stop() / reps
    prof = {k: v[0] / reps for k, v in ctx.profile_get().items()}
    ctx.profile(False, reset=False)
    st = sw.stats(0)
    out = {"config": name, "width": W, "height": Hh, "faces": sc.nfaces, "ms_per_frame": ms,
           "mtri_per_s": 2 * sc.nfaces / ms / 1e3, "covered_mpix_per_s": st["pixels_covered"] / ms / 1e3,
           "tile_refs": st["tile_refs"], "tris_out": st["tris_out"], "pixels_covered": st["pixels_covered"], "kernel_ms": prof}
    print(json.dumps(out))
    for o in (sw,) + tuple(objs):
        o.close()


def main():
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 5
    hana = ge.load_package()
    ctx = hana.Context(0)
    a2v = hana.scene.synthetic_grid(2237, 2237, seed=1234)
    dif, nm = hana.scene.noise_textures(1234, 1024, flat_normal=True)
    run(hana, ctx, "configs[3]", hana.Scene("c4", a2v, dif, nm), hana.BLINN, 3840, 2160, reps)
    a2v = hana.scene.synthetic_layers(8, 32, 18, seed=99)
    dif, nm = hana.scene.noise_textures(99, 1024)
    run(hana, ctx, "configs[4]", hana.Scene("c5", a2v, dif, nm), hana.NORMALMAP, 7680, 4320, reps)
    ctx.close()


if __name__ == "__main__":
    main()
