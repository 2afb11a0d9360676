"""The asynchronous sweep contract (include/hana_b200.h): a render returns before the GPU has run it and nothing is
read back inside it; if the batch ran out of internal scratch (triangle slots, tile-list records) the sweep's next
synchronising call renders it again with more. Checked by rendering, in ONE context sized by a tiny scene, a scene that
needs far more of both, and comparing with a context that never saw the tiny scene; and by interleaving two rings."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scenes(hana):
    small = hana.synthetic_scene("blob")
    a2v = hana.scene.synthetic_grid(300, 170, seed=5)          # 101 062 triangles, ~4 px each at 640x360
    dif, nm = hana.scene.noise_textures(5, 128, flat_normal=True)
    return small, hana.Scene("grid", a2v, dif, nm)


def test_overflowing_batch_is_rendered_again(hana):
    W, Hh, F = 640, 360, 3
    small, big = _scenes(hana)
    us = [hana.default_uniforms(W, Hh, True) for _ in range(F)]

    fresh = hana.Context(0)
    objs = big.upload(fresh)
    sw = fresh.sweep(W, Hh, F)
    sw.render(objs[0], hana.BLINN, us, objs[1], objs[2])
    want = sw.checksums(F)
    want_frame = sw.download(1)
    for o in (sw,) + tuple(objs):
        o.close()
    fresh.close()

    ctx = hana.Context(0)
    s_objs = small.upload(ctx)
    sw = ctx.sweep(W, Hh, F)
    sw.render(s_objs[0], hana.BLINN, us, s_objs[1], s_objs[2])   # sizes the scratch for ~1e3 triangles
    sw.checksums(F)
    b_objs = big.upload(ctx)
    sw.render(b_objs[0], hana.BLINN, us, b_objs[1], b_objs[2])   # needs 100x that: dropped, then re-rendered
    got = sw.checksums(F)
    assert np.array_equal(got, want)
    col, dep = sw.download(1)
    assert np.array_equal(col[..., :3], want_frame[0][..., :3]) and np.array_equal(dep.view(np.uint32), want_frame[1].view(np.uint32))
    st = sw.stats(1)
    assert st["faces_in"] == big.nfaces and st["tris_out"] > 1000
    # steady state: the same batch again must not need the second attempt and gives the same frames
    sw.render(b_objs[0], hana.BLINN, us, b_objs[1], b_objs[2])
    assert np.array_equal(sw.checksums(F), want)
    for o in (sw,) + tuple(s_objs) + tuple(b_objs):
        o.close()
    ctx.close()


def test_two_rings_with_overlapped_copies(hana, ctx):
    """Render into ring A, start its copy, render into ring B while the copy runs: both arrive intact."""
    W, Hh, F = 320, 240, 4
    scene = hana.synthetic_scene("blob")
    model, dtex, ntex = scene.upload(ctx)
    rings = [ctx.sweep(W, Hh, F), ctx.sweep(W, Hh, F)]
    pins = [hana.api.PinnedBuffer(W * Hh * 4 * F) for _ in range(2)]
    batches = [hana.orbit_sweep_uniforms(W, Hh, 16 * k, F, frames_per_turn=64) for k in range(4)]
    want = []
    for b in batches:
        rings[0].render(model, hana.BLINN, b, dtex, ntex)
        want.append(np.stack([rings[0].download(k)[0] for k in range(F)]))
    got = []
    for k, b in enumerate(batches):
        r = rings[k & 1]
        if k >= 2:                                           # the copy of batch k-2 used this pinned buffer
            ctx.sync()
            got.append(np.frombuffer(pins[k & 1].array, np.uint8).reshape(F, Hh, W, 4).copy())
        r.render(model, hana.BLINN, b, dtex, ntex)
        r.download_async(0, F, pins[k & 1].ptr, None)
    ctx.sync()
    got.append(np.frombuffer(pins[0].array, np.uint8).reshape(F, Hh, W, 4).copy())
    got.append(np.frombuffer(pins[1].array, np.uint8).reshape(F, Hh, W, 4).copy())
    for k in range(4):
        assert np.array_equal(got[k], want[k]), k
    for o in rings + [model, dtex, ntex] + pins:
        o.close()


def test_static_light_shadow_reuse_gives_identical_frames(hana, ctx):
    """hana_sweep_set_shadow_reuse (optional): an orbit batch (the light looks at the camera's fixed target:
    light_vp and model are the same in every frame, scene.h:69) rendered with ONE shadow map equals the batch rendered with
    a ShadowShader pass per frame, bit for bit; a batch whose light moves is rendered pass by pass whatever the setting."""
    W, Hh, F = 640, 360, 6
    scene = hana.synthetic_scene("blob")
    model, dtex, ntex = scene.upload(ctx)
    orbit = hana.orbit_sweep_uniforms(W, Hh, 3, F, frames_per_turn=32)
    moving = hana.orbit_sweep_uniforms(W, Hh, 3, F, frames_per_turn=32)
    moving[4].light_vp[3] += 0.05  # one frame's light elsewhere
    sw = ctx.sweep(W, Hh, F)
    for batch in (orbit, moving):
        sw.set_shadow_reuse(False)
        sw.render(model, hana.BLINN, batch, dtex, ntex)
        want = [sw.download(k) for k in range(F)]
        l0 = ctx.launches
        sw.set_shadow_reuse(True)
        sw.render(model, hana.BLINN, batch, dtex, ntex)
        for k in range(F):
            col, dep = sw.download(k)
            assert np.array_equal(col, want[k][0]) and np.array_equal(dep.view(np.uint32), want[k][1].view(np.uint32)), k
        assert ctx.launches > l0
    assert sw.overflow_count() == 0
    for o in (sw, model, dtex, ntex):
        o.close()


def test_pipelined_submissions_equal_serial_ones(hana):
    """Back-to-back submissions are pipelined (the next submission is binned on side streams, with the scratch
    sets and uniform block of the other parity, while the current one is rasterised). Six batches with different cameras
    and two scenes, queued without a synchronising call in between into one ring and into two alternating rings, must give
    the frames a context with HANA_NO_PIPELINE=1 gives, bit for bit."""
    import os
    W, Hh, F = 480, 270, 5
    blob = hana.synthetic_scene("blob")
    a2v = hana.scene.synthetic_grid(60, 40, seed=9)
    dif, nm = hana.scene.noise_textures(9, 64, flat_normal=True)
    grid = hana.Scene("grid", a2v, dif, nm)
    batches = [hana.orbit_sweep_uniforms(W, Hh, 11 * k, F, frames_per_turn=64) for k in range(6)]

    def run(pipelined):
        old = os.environ.pop("HANA_NO_PIPELINE", None)
        if not pipelined:
            os.environ["HANA_NO_PIPELINE"] = "1"
        try:
            ctx = hana.Context(0)
        finally:
            os.environ.pop("HANA_NO_PIPELINE", None)
            if old is not None:
                os.environ["HANA_NO_PIPELINE"] = old
        objs = [blob.upload(ctx), grid.upload(ctx)]
        rings = [ctx.sweep(W, Hh, F), ctx.sweep(W, Hh, F)]
        # six submissions queued back to back (nothing synchronises the host in between), alternating rings, scenes and
        # shaders: what the last two left in the rings
        for k, b in enumerate(batches):
            o = objs[k & 1]
            rings[k & 1].render(o[0], hana.BLINN if k % 3 else hana.NORMALMAP, b, o[1], o[2])
        sums = [rings[0].checksums(F), rings[1].checksums(F)]
        # two rings alternating, frames read after the NEXT submission has been queued
        frames = []
        for k, b in enumerate(batches):
            o = objs[(k >> 1) & 1]
            rings[k & 1].render(o[0], hana.BLINN, b, o[1], o[2])
            if k:
                frames.append(np.stack([rings[(k - 1) & 1].download(i)[0] for i in range(F)]))
        frames.append(np.stack([rings[(len(batches) - 1) & 1].download(i)[0] for i in range(F)]))
        assert rings[0].overflow_count() == 0 and rings[1].overflow_count() == 0
        for r in rings:
            r.close()
        for o in objs:
            for x in o:
                x.close()
        ctx.close()
        return sums, frames

    s_pipe, f_pipe = run(True)
    s_ser, f_ser = run(False)
    for k in range(2):
        assert np.array_equal(s_pipe[k], s_ser[k]), "ring %d after six queued submissions" % k
    for k in range(len(batches)):
        assert np.array_equal(f_pipe[k], f_ser[k]), "batch %d (two rings)" % k
