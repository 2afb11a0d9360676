"""BASELINE.json configs[3] (dense tiny-triangle mesh, setup/binning-bound) and configs[4] (large triangles, depth
complexity ~8, NormalMap + shadow, fragment-bound): against the oracle at sizes it finishes in seconds, and at the
FULL sizes (10 M triangles @ 3840x2160; 9 216 triangles @ 7680x4320) through size-independent properties:
two independent device code paths agree bit for bit (batched sweep with folded clears and R8 shadow maps vs. the
RenderBuffer read-modify-write path, TMA vs plain stores), a draw is idempotent, and coverage statistics add up."""
import numpy as np
import pytest

from conftest import cleared, compare_frames
from test_gpu_parity import check, oracle_two_pass

pytestmark = pytest.mark.gpu
FLT_MAX = np.float32(3.4028234663852886e38)


def rmw_two_pass(ctx, hana, objs, shader, u, W, Hh):
    model, dtex, ntex = objs
    frame, shadow = ctx.renderbuffer(W, Hh), ctx.renderbuffer(W, Hh)
    for rb in (frame, shadow):
        rb.clear_color(0, 0, 0, 1)
        rb.clear_depth(FLT_MAX)
    ctx.draw_model(frame, shadow, model, shader, u, dtex, ntex)
    col, dep = frame.download()
    st = ctx.stats()
    frame.close()
    shadow.close()
    return col, dep, st


def test_c4_scaled_vs_oracle(hana, horacle, port, ctx):
    """Same generator as configs[3], 480 x 270 vertices (257 642 triangles) at 960x540: ~1.9 x 1.9 px quads."""
    W, Hh = 960, 540
    a2v = hana.scene.synthetic_grid(480, 270, seed=1234)
    dif, nm = hana.scene.noise_textures(1234, 256, flat_normal=True)
    sc = hana.Scene("c4_small", a2v, dif, nm)
    u = hana.default_uniforms(W, Hh, True)
    hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
    col, dep, pid, _, _ = oracle_two_pass(port, horacle, horacle.BLINN, hu, sc, W, Hh)
    objs = sc.upload(ctx)
    gcol, gdep, st = rmw_two_pass(ctx, hana, objs, hana.BLINN, u, W, Hh)
    check(compare_frames(gcol, gdep, col, dep), W * Hh)
    assert st["faces_in"] == sc.nfaces and st["pixels_covered"] == int((pid != 0xFFFFFFFF).sum())
    sw = ctx.sweep(W, Hh, 1)
    sw.render(objs[0], hana.BLINN, [u], objs[1], objs[2])
    scol, sdep = sw.download(0)
    assert np.array_equal(scol, gcol) and np.array_equal(sdep.view(np.uint32), gdep.view(np.uint32))
    for o in (sw,) + tuple(objs):
        o.close()


def test_c5_scaled_vs_oracle(hana, horacle, port, ctx):
    """configs[4] generator (8 layers of 32x18 quads, back to front) at 1280x720, NormalMap + shadow."""
    W, Hh = 1280, 720
    a2v = hana.scene.synthetic_layers(8, 32, 18, seed=99)
    dif, nm = hana.scene.noise_textures(99, 512)
    sc = hana.Scene("c5_small", a2v, dif, nm)
    u = hana.default_uniforms(W, Hh, True)
    hu = horacle.HanaUniforms.from_bytes(u.to_bytes())
    col, dep, pid, _, _ = oracle_two_pass(port, horacle, horacle.NORMALMAP, hu, sc, W, Hh)
    objs = sc.upload(ctx)
    gcol, gdep, st = rmw_two_pass(ctx, hana, objs, hana.NORMALMAP, u, W, Hh)
    check(compare_frames(gcol, gdep, col, dep), W * Hh)
    assert (pid != 0xFFFFFFFF).mean() > 0.9  # screen-filling layers
    for o in objs:
        o.close()


def full_size_properties(ctx, hana, sc, shader, W, Hh):
    u = hana.default_uniforms(W, Hh, True)
    objs = sc.upload(ctx)
    # path A: RenderBuffer read-modify-write, TMA tile load/store
    ctx.set_tma(True)
    colA, depA, stA = rmw_two_pass(ctx, hana, objs, shader, u, W, Hh)
    # path B: batched sweep, clears folded into the flush, R8 shadow map
    sw = ctx.sweep(W, Hh, 1)
    sw.render(objs[0], shader, [u], objs[1], objs[2])
    colB, depB = sw.download(0)
    from hana_softwarerenderer_b200.api import frame_checksum
    assert int(sw.checksums(1)[0]) == int(frame_checksum(colB, depB))
    assert sw.stats(0)["pixels_covered"] == stA["pixels_covered"]
    sw.close()
    assert np.array_equal(depA.view(np.uint32), depB.view(np.uint32))
    assert np.array_equal(colA, colB)
    # path C: plain global stores instead of TMA
    ctx.set_tma(False)
    colC, depC, stC = rmw_two_pass(ctx, hana, objs, shader, u, W, Hh)
    ctx.set_tma(True)
    assert np.array_equal(depA.view(np.uint32), depC.view(np.uint32)) and np.array_equal(colA, colC)
    assert stA["tile_refs"] == stC["tile_refs"] and stA["tris_out"] == stC["tris_out"]
    # idempotence: a second identical draw ties with itself everywhere (z == stored -> rewritten, graphics.cpp:359)
    rb = ctx.renderbuffer(W, Hh)
    rb.upload(colA, depA)
    u2 = u.copy()
    u2.enable_shadow = 0
    ctx.draw(rb, objs[0], hana.GROUND, u2)
    ctx.draw(rb, objs[0], hana.GROUND, u2)
    c1, d1 = rb.download()
    ctx.draw(rb, objs[0], hana.GROUND, u2)
    c2, d2 = rb.download()
    assert np.array_equal(c1, c2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
    assert np.array_equal(d1.view(np.uint32), depA.view(np.uint32))  # same geometry, same camera: depth unchanged
    rb.close()
    covered = depA != FLT_MAX
    assert int(covered.sum()) == stA["pixels_covered"]
    for o in objs:
        o.close()
    return stA, covered.mean()


def test_c4_full_size_properties(hana, ctx):
    """configs[3]: 2237 x 2237 vertices -> 9 999 392 triangles at 3840x2160, Blinn + shadow."""
    a2v = hana.scene.synthetic_grid(2237, 2237, seed=1234)
    assert a2v.shape[0] == 9999392 * 3
    dif, nm = hana.scene.noise_textures(1234, 1024, flat_normal=True)
    st, cov = full_size_properties(ctx, hana, hana.Scene("c4", a2v, dif, nm), hana.BLINN, 3840, 2160)
    assert st["faces_in"] == 9999392 and st["tris_out"] > 9_000_000 and cov > 0.9


def test_c5_full_size_properties(hana, ctx):
    """configs[4]: 8 layers x 32 x 18 quads x 2 = 9 216 triangles at 7680x4320, NormalMap + shadow."""
    a2v = hana.scene.synthetic_layers(8, 32, 18, seed=99)
    assert a2v.shape[0] == 9216 * 3
    dif, nm = hana.scene.noise_textures(99, 1024)
    st, cov = full_size_properties(ctx, hana, hana.Scene("c5", a2v, dif, nm), hana.NORMALMAP, 7680, 4320)
    assert st["tris_out"] == 9216 and cov > 0.9 and st["tile_refs"] > 1_000_000
