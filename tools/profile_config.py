#!/usr/bin/env python
"""One profiled frame of BASELINE.json configs[3] or configs[4] for ncu (3 warm-up frames, then one between
cudaProfilerStart/Stop).   ncu --set full --profile-from-start off -o out python tools/profile_config.py c4"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c4"
    import torch
    hana = ge.load_package()
    ctx = hana.Context(0)
    if which == "c4":
        a2v = hana.scene.synthetic_grid(2237, 2237, seed=1234)
        dif, nm = hana.scene.noise_textures(1234, 1024, flat_normal=True)
        sc, shader, W, H = hana.Scene("c4", a2v, dif, nm), hana.BLINN, 3840, 2160
    else:
        a2v = hana.scene.synthetic_layers(8, 32, 18, seed=99)
        dif, nm = hana.scene.noise_textures(99, 1024)
        sc, shader, W, H = hana.Scene("c5", a2v, dif, nm), hana.NORMALMAP, 7680, 4320
    objs = sc.upload(ctx)
    u = hana.default_uniforms(W, H, True)
    sw = ctx.sweep(W, H, 1)
    for _ in range(3):
        sw.render(objs[0], shader, [u], objs[1], objs[2])
        ctx.sync()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    sw.render(objs[0], shader, [u], objs[1], objs[2])
    ctx.sync()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ctx.close()


if __name__ == "__main__":
    main()
