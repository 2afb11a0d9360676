#!/usr/bin/env python
"""Times the device-side TGA encoder (hana_sweep_encode_tga) and the fetch of its files, per batch size and frame size.
  python tools/time_tga.py"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

hana = ge.load_package()
from hana_softwarerenderer_b200.api import PinnedBuffer  # noqa: E402

ctx = hana.Context(0)
sc = hana.load_bundled("african_head", None, 3)
model, dtex, ntex = sc.upload(ctx)
CASES = ((1920, 1080, 1), (1920, 1080, 16), (960, 540, 16), (3840, 2160, 4), (1920, 1080, 128), (1920, 1080, 512))
if os.environ.get('TGA_CASES'):
    CASES = tuple(tuple(int(v) for v in c.split('x')) for c in os.environ['TGA_CASES'].split(','))
REPS = int(os.environ.get('TGA_REPS', '4'))
for (W, H, F) in CASES:
    sw = ctx.sweep(W, H, F)
    arr = hana.orbit_sweep_uniforms(W, H, 0, F, frames_per_turn=1024)
    sw.render(model, hana.BLINN, arr, dtex, ntex)
    ctx.sync()
    pin = PinnedBuffer(F * W * H * 3)
    offs = (C.c_uint64 * (F + 1))()
    szs = (C.c_uint64 * F)()
    best = 1e9
    for rep in range(REPS):
        ctx.timer_start()
        assert ctx.L.hana_sweep_encode_tga(sw.h, 0, F) == 0
        ms_enc = ctx.timer_stop()
        t0 = time.perf_counter()
        assert ctx.L.hana_sweep_fetch_tga(sw.h, C.c_void_p(pin.ptr), C.c_size_t(F * W * H * 3), offs, szs) == 0
        ctx.sync()
        ms_fetch = (time.perf_counter() - t0) * 1e3
        best = min(best, ms_enc)
    print("%dx%d F=%d: encode %.3f ms = %.2f us/frame; fetch %.3f ms, %.0f bytes/frame, %.1f GB/s" %
          (W, H, F, best, best * 1e3 / F, ms_fetch, offs[F] / F, offs[F] / ms_fetch / 1e6))
    pin.close()
    sw.close()
