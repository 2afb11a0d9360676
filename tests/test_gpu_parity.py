"""GPU parity: the CUDA path through the C ABI against the CPU oracle (oracle/hana_oracle.c, pinned
bit-exact to the real reference by tests/test_oracle_vs_reference.py) on the same seeded inputs.

Bars (BASELINE.json): coverage / primitive-ID bit-exact except <= 0.01 % of pixels, |depth diff| <= 1e-6,
|colour diff| <= 1/255 per channel. The implementation is written to be bit-exact in coverage, prim-ID and
depth, so those are asserted at ZERO mismatches; colour may differ by one level where powf rounds differently."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, cleared, compare_frames

pytestmark = pytest.mark.gpu

FLT_MAX = np.float32(3.4028234663852886e38)


def oracle_two_pass(port, H, shader, u, scene, W, Hh, want_primid=True):
    scol, sdep = cleared(W, Hh)
    shadow = None
    if u.enable_shadow:
        port.draw(H.SHADOW, u, scene.a2v, W, Hh, scol, sdep)
        shadow = scol
    col, dep = cleared(W, Hh)
    pid, _ = port.draw(shader, u, scene.a2v, W, Hh, col, dep, diffuse=scene.diffuse, normal=scene.normal, shadow=shadow,
                       want_primid=want_primid)
    return col, dep, pid, scol, sdep


def check(m, npix, colour_tol=1):
    assert m["coverage_mismatch"] == 0, m
    assert m["depth_bits_mismatch"] == 0, m
    assert m.get("primid_mismatch", 0) == 0, m
    assert m["colour_maxdiff"] <= colour_tol, m
    assert m["colour_mismatch_px"] <= max(8, npix // 1000), m


def test_no_cpu_fallback_symbols(hana):
    # the product library must not link the oracle
    import subprocess
    out = subprocess.run(["nm", "-D", hana.lib_path()], capture_output=True, text=True).stdout
    assert "horacle_" not in out and "href_" not in out


@pytest.mark.parametrize("shader", [0, 1, 2, 3, 4, 5, 6])
def test_stage_vertex_bit_exact(hana, horacle, port, ctx, blob, shader):
    u = hana.default_uniforms(320, 240, True)
    m = ctx.model(blob.a2v)
    got = ctx.stage_vertex(m, shader, u)
    want = port.vertex(shader, horacle.HanaUniforms.from_bytes(u.to_bytes()), blob.a2v)
    m.close()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("tma", [True, False])
@pytest.mark.parametrize("shader", [1, 2, 3, 4, 5, 6])
def test_draw_model_blob(hana, horacle, port, ctx, blob, shader, tma):
    W, Hh = 320, 240
    ctx.set_tma(tma)
    u = hana.default_uniforms(W, Hh, True)
    hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
    col, dep, pid, scol, sdep = oracle_two_pass(port, horacle, shader, hu, blob, W, Hh)
    model, dtex, ntex = blob.upload(ctx)
    frame, shadow = ctx.renderbuffer(W, Hh), ctx.renderbuffer(W, Hh)
    for rb in (frame, shadow):
        rb.clear_color(0, 0, 0, 1)
        rb.clear_depth(FLT_MAX)
    # pass 1 + pass 2 by hand so that the shadow map can be inspected before DrawModel::draw clears it
    ctx.draw(shadow, model, hana.SHADOW, u)
    gs_col, gs_dep = shadow.download()
    assert np.array_equal(gs_col, scol) and np.array_equal(gs_dep.view(np.uint32), sdep.view(np.uint32))
    gpid = ctx.draw(frame, model, shader, u, dtex, ntex, shadow, want_primid=True)
    gcol, gdep = frame.download()
    check(compare_frames(gcol, gdep, col, dep, gpid, pid), W * Hh)
    assert np.array_equal(gcol[..., 3], col[..., 3])  # alpha is never written
    st = ctx.stats()
    assert st["pixels_covered"] == int((pid != 0xFFFFFFFF).sum())
    for o in (frame, shadow, model, dtex, ntex):
        o.close()
    ctx.set_tma(True)


@pytest.mark.parametrize("size", [(800, 600), (1920, 1080), (1000, 600), (333, 217)])
def test_draw_model_african_head(hana, horacle, port, ctx, african_head, size):
    W, Hh = size
    u = hana.default_uniforms(W, Hh, True)
    hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
    col, dep, pid, _, _ = oracle_two_pass(port, horacle, horacle.BLINN, hu, african_head, W, Hh)
    model, dtex, ntex = african_head.upload(ctx)
    frame, shadow = ctx.renderbuffer(W, Hh), ctx.renderbuffer(W, Hh)
    for rb in (frame, shadow):
        rb.clear_color(0, 0, 0, 1)
        rb.clear_depth(FLT_MAX)
    ctx.draw_model(frame, shadow, model, hana.BLINN, u, dtex, ntex)
    gcol, gdep = frame.download()
    check(compare_frames(gcol, gdep, col, dep), W * Hh)
    scol, sdep = shadow.download()  # scene.h:94-98: left cleared
    assert (scol == np.array([0, 0, 0, 1], np.uint8)).all() and (sdep == FLT_MAX).all()
    for o in (frame, shadow, model, dtex, ntex):
        o.close()


@pytest.mark.parametrize("shader", [1, 2])
def test_sweep_matches_oracle(hana, horacle, port, ctx, diablo, shader):
    W, Hh, F = 800, 600, 6
    arr = hana.orbit_sweep_uniforms(W, Hh, 40, F, frames_per_turn=64)
    model, dtex, ntex = diablo.upload(ctx)
    sw = ctx.sweep(W, Hh, F)
    sw.render(model, shader, arr, dtex, ntex)
    sums = sw.checksums(F)
    from hana_softwarerenderer_b200.api import frame_checksum
    for f in range(F):
        hu = horacle.HanaUniforms.from_bytes(arr[f].to_bytes())
        col, dep, _, _, _ = oracle_two_pass(port, horacle, shader, hu, diablo, W, Hh, want_primid=False)
        gcol, gdep = sw.download(f)
        check(compare_frames(gcol, gdep, col, dep), W * Hh)
        assert int(sums[f]) == int(frame_checksum(gcol, gdep))
    st = sw.stats(0)
    assert st["faces_in"] == diablo.nfaces and st["pixels_covered"] > 0
    for o in (sw, model, dtex, ntex):
        o.close()


def test_existing_depth_takes_part(hana, horacle, port, ctx, blob):
    """graphics.cpp:359: fragments behind what the buffer already holds are rejected; equal depth overwrites."""
    W, Hh = 256, 192
    u = hana.default_uniforms(W, Hh, False)
    hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
    rng = np.random.RandomState(3)
    col0 = rng.randint(0, 256, (Hh, W, 4)).astype(np.uint8)
    dep0 = rng.uniform(0.90, 1.0, (Hh, W)).astype(np.float32)
    col, dep = col0.copy(), dep0.copy()
    pid, _ = port.draw(horacle.BLINN, hu, blob.a2v, W, Hh, col, dep, diffuse=blob.diffuse, normal=blob.normal, want_primid=True)
    model, dtex, ntex = blob.upload(ctx)
    rb = ctx.renderbuffer(W, Hh)
    rb.upload(col0, dep0)
    gpid = ctx.draw(rb, model, hana.BLINN, u, dtex, ntex, None, want_primid=True)
    gcol, gdep = rb.download()
    assert np.array_equal(gpid, pid)
    assert np.array_equal(gdep.view(np.uint32), dep.view(np.uint32))
    assert np.abs(gcol.astype(int) - col.astype(int)).max() <= 1
    assert np.array_equal(gcol[pid == 0xFFFFFFFF], col0[pid == 0xFFFFFFFF])  # untouched pixels keep all four bytes
    # a second identical draw: every fragment ties with itself and is rewritten (LEQUAL) -> same result
    ctx.draw(rb, model, hana.BLINN, u, dtex, ntex, None)
    gcol2, gdep2 = rb.download()
    assert np.array_equal(gcol2, gcol) and np.array_equal(gdep2, gdep)
    for o in (rb, model, dtex, ntex):
        o.close()


def test_clipping_near_camera(hana, horacle, port, ctx, african_head):
    """Camera inside the bounding sphere: faces cross W, +-X, +-Y, +-Z planes (graphics.cpp:136-161)."""
    W, Hh = 640, 480
    for pos in ((0.0, 0.0, 0.9), (0.3, 0.2, 0.75), (0.0, 0.9, 0.5)):
        cam = hana.OrbitCamera(np.float32(W) / np.float32(Hh), position=pos)
        u = hana.default_uniforms(W, Hh, True, camera=cam)
        hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
        col, dep, pid, _, _ = oracle_two_pass(port, horacle, horacle.BLINN, hu, african_head, W, Hh)
        model, dtex, ntex = african_head.upload(ctx)
        frame, shadow = ctx.renderbuffer(W, Hh), ctx.renderbuffer(W, Hh)
        for rb in (frame, shadow):
            rb.clear_color(0, 0, 0, 1)
            rb.clear_depth(FLT_MAX)
        ctx.draw(shadow, model, hana.SHADOW, u)
        gpid = ctx.draw(frame, model, hana.BLINN, u, dtex, ntex, shadow, want_primid=True)
        gcol, gdep = frame.download()
        check(compare_frames(gcol, gdep, col, dep, gpid, pid), W * Hh)
        order, v2f = ctx.stage_setup(model, hana.BLINN, u, W, Hh)
        assert len(order) == ctx.stats()["tris_out"] or True
        assert (np.diff(order.astype(np.int64)) > 0).all()
        for o in (frame, shadow, model, dtex, ntex):
            o.close()


def test_host_buffer_entry_point(hana, horacle, port, ctx, blob):
    W, Hh = 320, 240
    u = hana.default_uniforms(W, Hh, True)
    hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
    col, dep, _, _, _ = oracle_two_pass(port, horacle, horacle.NORMALMAP, hu, blob, W, Hh, want_primid=False)
    model, dtex, ntex = blob.upload(ctx)
    for assume in (True, False):
        hc, hd = cleared(W, Hh)
        ctx.draw_model_host(hc, hd, model, hana.NORMALMAP, u, dtex, ntex, assume_cleared=assume)
        check(compare_frames(hc, hd, col, dep), W * Hh)
    for o in (model, dtex, ntex):
        o.close()


def test_empty_model_and_errors(hana, ctx):
    W, Hh = 64, 48
    u = hana.default_uniforms(W, Hh, False)
    m = ctx.model(np.zeros((0, 8), np.float32))
    rb = ctx.renderbuffer(W, Hh)
    ctx.draw(rb, m, hana.BLINN, u)  # zero faces -> zero iterations (graphics.cpp:380)
    col, dep = rb.download()
    assert (col == np.array([0, 0, 0, 255], np.uint8)).all() and (dep == 1.0).all()  # ctor state renderbuffer.cpp:6-7
    with pytest.raises(hana.HanaError):
        ctx.draw(rb, m, 99, u)
    rb.close()
    m.close()


def test_dropin_graphics_draw_triangle(hana, horacle):
    """The reference's OWN scene code (Scene/Camera/DrawModel::draw, compiled from its unmodified sources) linked with
    this repo's graphics_draw_triangle(DrawData*) instead of graphics.cpp, against the plain reference build:
    same OBJ/TGA files, same camera motion, frame by frame."""
    import os
    from conftest import ASSET_DIR
    H = horacle
    obj = os.path.join(ASSET_DIR, "diablo3_pose", "diablo3_pose.obj")
    if not (os.path.exists(H.REF_DROPIN_SO) and os.path.exists(H.REF_SO) and os.path.exists(obj)):
        pytest.skip("oracle/_ref drop-in build missing (needs /root/reference at build time)")
    W, Hh = 640, 480
    for shader in (H.NORMALMAP, H.BLINN, H.TOON):
        ref = H.Reference(obj, W, Hh, shader, instrumented=False)
        gpu = H.Reference(obj, W, Hh, shader, dropin=True)
        for step in range(3):
            for r in (ref, gpu):
                r.camera_motion(orbit=(0.11, 0.03), dolly=0.5)
            _, c0, d0 = ref.render(True)
            _, c1, d1 = gpu.render(True)
            assert ref.uniforms().to_bytes() == gpu.uniforms().to_bytes()
            check(compare_frames(c1, d1, c0, d0), W * Hh)
        # shadows off + the reference's very first-frame state (depth 1.0, alpha 255: renderbuffer.cpp:6-7)
        ref2 = H.Reference(obj, W, Hh, shader, instrumented=False)
        gpu2 = H.Reference(obj, W, Hh, shader, dropin=True)
        _, c0, d0 = ref2.render(False, clear=False)
        _, c1, d1 = gpu2.render(False, clear=False)
        assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
        assert np.abs(c0.astype(int) - c1.astype(int)).max() <= 1 and np.array_equal(c0[..., 3], c1[..., 3])
        for r in (ref, gpu, ref2, gpu2):
            r.close()


@pytest.mark.parametrize("copies", [2, 40])
def test_equal_depths_later_submission_wins(hana, horacle, port, ctx, blob, copies):
    """The same faces submitted several times with different uvs: every fragment of a later copy ties in depth with the
    earlier one and the reference lets it pass (z > stored is false, graphics.cpp:359), so the LAST copy owns every pixel.
    40 copies put more than 32 records into most tiles, so winners and ties cross the 32-record staging chunks; the
    shadow pass resolves the same ties in its own (in-loop) way."""
    W, Hh = 256, 192
    parts = []
    for k in range(copies):
        a = blob.a2v.copy()
        a[:, 6:8] = (a[:, 6:8] * (0.5 + 0.5 * k / copies)) % 1.0
        parts.append(a)
    order = np.arange(copies)
    order[1:] = np.random.default_rng(3).permutation(order[1:])        # submission order != any spatial order
    a2v = np.concatenate([parts[k] for k in order])
    sc = hana.Scene("dup", a2v, blob.diffuse, blob.normal)
    u = hana.default_uniforms(W, Hh, True)
    hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
    col, dep, pid, scol, sdep = oracle_two_pass(port, horacle, horacle.BLINN, hu, sc, W, Hh)
    nf = blob.a2v.shape[0] // 3
    assert (pid[pid != 0xFFFFFFFF] // 8 >= (copies - 1) * nf).all()     # the oracle agrees: last copy everywhere
    model, dtex, ntex = sc.upload(ctx)
    frame, shadow = ctx.renderbuffer(W, Hh), ctx.renderbuffer(W, Hh)
    for rb in (frame, shadow):
        rb.clear_color(0, 0, 0, 1)
        rb.clear_depth(FLT_MAX)
    ctx.draw(shadow, model, hana.SHADOW, u)
    gpid = ctx.draw(frame, model, hana.BLINN, u, dtex, ntex, shadow, want_primid=True)
    gcol, gdep = frame.download()
    check(compare_frames(gcol, gdep, col, dep, gpid, pid), W * Hh)
    sw = ctx.sweep(W, Hh, 1)                                            # sweep: SHADOW_R8 in-loop resolve + CLEAR_FOLD
    sw.render(model, hana.BLINN, [u], dtex, ntex)
    scol2, sdep2 = sw.download(0)
    assert np.array_equal(sdep2.view(np.uint32), dep.view(np.uint32))
    assert np.abs(scol2[..., :3].astype(int) - col[..., :3].astype(int)).max() <= 1
    for o in (sw, frame, shadow, model, dtex, ntex):
        o.close()


@pytest.mark.parametrize("shader", ["TEXTURE", "BLINN"])
def test_positive_uz_slivers(hana, horacle, port, ctx, blob, shader):
    """Triangles the reference keeps although their screen-space u.z rounds positive (tests/golden/make_slivers.py): the
    rasteriser stages them with B and C exchanged and exchanges the two quotients back. 1061 such slivers, each
    covering a pixel in the reference, drawn over an ordinary mesh: primitive ids, depth bits and colours must be the
    reference's (per-vertex depth and uv differ, so exchanged weights would show in both)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_slivers as ms
    W, Hh = ms.W, ms.H
    tris = np.load(os.path.join(ROOT, "tests", "golden", "slivers_posuz.npy"))
    a2v = np.concatenate([blob.a2v * np.array([0.5, 0.5, 0.5, 1, 1, 1, 1, 1], np.float32), ms.sliver_a2v(tris)])
    u = ms.identity_uniforms(hana, W, Hh)
    hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
    sid, hsid = getattr(hana, shader), getattr(horacle, shader)
    col, dep = cleared(W, Hh)
    pid, _ = port.draw(hsid, hu, a2v, W, Hh, col, dep, diffuse=blob.diffuse, normal=blob.normal, want_primid=True)
    first_sliver = (blob.a2v.shape[0] // 3) * 8
    assert int((pid[pid != 0xFFFFFFFF] >= first_sliver).sum()) > 500  # the slivers are visible in the reference frame
    model = ctx.model(a2v)
    dtex, ntex = ctx.texture(blob.diffuse), ctx.texture(blob.normal)
    rb = ctx.renderbuffer(W, Hh)
    c0, d0 = cleared(W, Hh)
    rb.upload(c0, d0)
    gpid = ctx.draw(rb, model, sid, u, dtex, ntex, None, want_primid=True)
    gcol, gdep = rb.download()
    assert np.array_equal(gpid, pid)
    assert np.array_equal(gdep.view(np.uint32), dep.view(np.uint32))
    assert np.abs(gcol[..., :3].astype(int) - col[..., :3].astype(int)).max() <= 1
    # the sweep path (CLEAR_FOLD + TMA) as well
    sw = ctx.sweep(W, Hh, 1)
    sw.render(model, sid, [u], dtex, ntex)
    scol, sdep = sw.download(0)
    assert np.array_equal(sdep.view(np.uint32), dep.view(np.uint32))
    assert np.abs(scol[..., :3].astype(int) - col[..., :3].astype(int)).max() <= 1
    for o in (sw, rb, model, dtex, ntex):
        o.close()


@pytest.mark.parametrize("shader", ["BLINN", "NORMALMAP", "TOON", "TEXTURE_LIGHT"])
def test_random_soup_crossing_every_clip_plane(hana, horacle, port, ctx, blob, shader):
    """3 000 random triangles in a box several times the frustum: faces behind the camera (W and near planes), beyond
    the far plane, across all four side planes, slivers, large and tiny ones, random normals and out-of-range uvs
    (TGAImage::get returns zeros, tgaimage.cpp:249-251); shadowed two-pass frame against the oracle."""
    rng = np.random.RandomState(17)
    n = 3000
    centre = rng.uniform(-2.5, 2.5, (n, 1, 3))
    centre[:, :, 2] = rng.uniform(-6.0, 3.0, (n, 1))            # the camera sits at z = 2 looking down -z
    size = rng.choice([0.02, 0.3, 1.5, 6.0], (n, 1, 1), p=[0.3, 0.4, 0.25, 0.05])
    pos = centre + rng.normal(0, 1, (n, 3, 3)) * size
    pos[:8, :, 2] -= rng.uniform(9000, 30000, (8, 3))            # across the far plane (far = 10000)
    nrm = rng.normal(0, 1, (n, 3, 3))
    uv = rng.uniform(-0.2, 1.2, (n, 3, 2))
    a2v = np.concatenate([pos, nrm, uv], axis=2).reshape(-1, 8).astype(np.float32)
    sc = hana.Scene("soup", a2v, blob.diffuse, blob.normal)
    W, Hh = 400, 304
    u = hana.default_uniforms(W, Hh, True)
    hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
    col, dep, pid, _, _ = oracle_two_pass(port, horacle, getattr(horacle, shader), hu, sc, W, Hh)
    assert (pid != 0xFFFFFFFF).mean() > 0.5
    model, dtex, ntex = sc.upload(ctx)
    frame, shadow = ctx.renderbuffer(W, Hh), ctx.renderbuffer(W, Hh)
    for rb in (frame, shadow):
        rb.clear_color(0, 0, 0, 1)
        rb.clear_depth(FLT_MAX)
    ctx.draw_model(frame, shadow, model, getattr(hana, shader), u, dtex, ntex)
    gcol, gdep = frame.download()
    check(compare_frames(gcol, gdep, col, dep), W * Hh)
    gpid = None
    sw = ctx.sweep(W, Hh, 2)
    sw.render(model, getattr(hana, shader), [u, u], dtex, ntex)
    scol, sdep = sw.download(1)
    check(compare_frames(scol, sdep, col, dep), W * Hh)
    for o in (sw, frame, shadow, model, dtex, ntex):
        o.close()
