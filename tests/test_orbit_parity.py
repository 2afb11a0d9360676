"""Parity of the frames bench.py TIMES (BASELINE.json configs[1]/[2]; SURVEY.md §8(d) C3: "parity-check a fixed subset,
e.g. every 64th frame"): the african_head 1080p orbit, cameras k = 0, 64, ..., 960 of the 1024-frame turn, rendered
through hana_sweep_render exactly as the bench does (one batch, ShadowShader pass + BlinnShader pass per frame), against
the CPU oracle on the same uniforms. Reference path: DrawModel::draw scene.h:53-99 driven by Camera::update_transform
camera.cpp:63-70. Plus the bundled-scene cases the round-1 review found missing: configs[0] (800x600, shadow OFF) and
diablo3_pose at 1080p with NormalMapShader.

Bars: coverage, depth bits and primitive ids at ZERO mismatches; colour <= 1/255 per channel (stated tolerance)."""
import numpy as np
import pytest

from conftest import compare_frames
from test_gpu_parity import check, oracle_two_pass

pytestmark = pytest.mark.gpu
FLT_MAX = np.float32(3.4028234663852886e38)
W, Hh, ORBIT = 1920, 1080, 1024


def test_bench_orbit_frames_every_64th(hana, horacle, port, ctx, african_head):
    ks = list(range(0, ORBIT, 64))
    arr = (hana.HanaUniforms * len(ks))()
    for i, k in enumerate(ks):
        arr[i] = hana.orbit_sweep_uniforms(W, Hh, k, 1, frames_per_turn=ORBIT)[0]
    model, dtex, ntex = african_head.upload(ctx)
    sw = ctx.sweep(W, Hh, len(ks))
    sw.render(model, hana.BLINN, arr, dtex, ntex)
    worst = 0
    for i, k in enumerate(ks):
        hu = horacle.HanaUniforms.from_bytes(arr[i].to_bytes())
        col, dep, pid, _, _ = oracle_two_pass(port, horacle, horacle.BLINN, hu, african_head, W, Hh, want_primid=(k % 256 == 0))
        gcol, gdep = sw.download(i)
        m = compare_frames(gcol, gdep, col, dep)
        check(m, W * Hh)
        worst = max(worst, m["colour_mismatch_px"])
        if k % 256 == 0:  # primitive ids through the RenderBuffer path for a quarter of them
            frame, shadow = ctx.renderbuffer(W, Hh), ctx.renderbuffer(W, Hh)
            for rb in (frame, shadow):
                rb.clear_color(0, 0, 0, 1)
                rb.clear_depth(FLT_MAX)
            ctx.draw(shadow, model, hana.SHADOW, arr[i])
            gpid = ctx.draw(frame, model, hana.BLINN, arr[i], dtex, ntex, shadow, want_primid=True)
            assert int((gpid != pid).sum()) == 0, "primitive ids differ at orbit frame %d" % k
            rcol, rdep = frame.download()
            assert np.array_equal(rdep.view(np.uint32), gdep.view(np.uint32)) and np.array_equal(rcol, gcol)
            frame.close()
            shadow.close()
    assert sw.overflow_count() == 0
    for o in (sw, model, dtex, ntex):
        o.close()


def test_c1_african_head_800x600_shadow_off(hana, horacle, port, ctx, african_head):
    """configs[0]: BlinnShader, 800x600, enable_shadow = false (scene.h:73 skips the ShadowShader pass)."""
    w, h = 800, 600
    u = hana.default_uniforms(w, h, False)
    hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
    col, dep, pid, _, _ = oracle_two_pass(port, horacle, horacle.BLINN, hu, african_head, w, h)
    model, dtex, ntex = african_head.upload(ctx)
    frame = ctx.renderbuffer(w, h)
    frame.clear_color(0, 0, 0, 1)
    frame.clear_depth(FLT_MAX)
    gpid = ctx.draw(frame, model, hana.BLINN, u, dtex, ntex, None, want_primid=True)
    gcol, gdep = frame.download()
    check(compare_frames(gcol, gdep, col, dep, gpid, pid), w * h)
    sw = ctx.sweep(w, h, 1)
    sw.render(model, hana.BLINN, [u], dtex, ntex)
    scol, sdep = sw.download(0)
    assert np.array_equal(sdep.view(np.uint32), dep.view(np.uint32))
    assert np.abs(scol[..., :3].astype(int) - col[..., :3].astype(int)).max() <= 1
    for o in (sw, frame, model, dtex, ntex):
        o.close()


@pytest.mark.parametrize("size", [(1920, 1080), (1000, 600)])
def test_diablo3_pose_normalmap(hana, horacle, port, ctx, diablo, size):
    """diablo3_pose with NormalMapShader + shadow: 1080p, and 1000x600 — the reference's only published workload
    (README.md:5, main.cpp:7-8, scene.cpp:86-103)."""
    w, h = size
    u = hana.default_uniforms(w, h, True)
    hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
    col, dep, pid, _, _ = oracle_two_pass(port, horacle, horacle.NORMALMAP, hu, diablo, w, h)
    model, dtex, ntex = diablo.upload(ctx)
    frame, shadow = ctx.renderbuffer(w, h), ctx.renderbuffer(w, h)
    for rb in (frame, shadow):
        rb.clear_color(0, 0, 0, 1)
        rb.clear_depth(FLT_MAX)
    ctx.draw(shadow, model, hana.SHADOW, u)
    gpid = ctx.draw(frame, model, hana.NORMALMAP, u, dtex, ntex, shadow, want_primid=True)
    gcol, gdep = frame.download()
    check(compare_frames(gcol, gdep, col, dep, gpid, pid), w * h)
    sw = ctx.sweep(w, h, 1)
    sw.render(model, hana.NORMALMAP, [u], dtex, ntex)
    scol, sdep = sw.download(0)
    assert np.array_equal(sdep.view(np.uint32), gdep.view(np.uint32)) and np.array_equal(scol, gcol)
    for o in (sw, frame, shadow, model, dtex, ntex):
        o.close()
