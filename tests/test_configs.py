"""BASELINE.json configs[3] (dense tiny-triangle mesh, setup/binning-bound) and configs[4] (large triangles, depth
complexity ~8, NormalMap + shadow, fragment-bound): against the oracle at sizes it finishes in seconds, and at the
FULL sizes (10 M triangles @ 3840x2160; 9 216 triangles @ 7680x4320) against the committed golden vectors the C port
of the reference produced at those sizes (tests/golden/make_configs_golden.py -> configs_full_golden.npz: depth and
primitive-id digests, row sums, 32x32 block colour sums over the whole frame, 256x256 colour crops), plus
size-independent properties: two independent device code paths agree bit for bit (batched sweep with folded clears and
R8 shadow maps vs. the RenderBuffer read-modify-write path, TMA vs plain stores), a draw is idempotent, and coverage
statistics add up."""
import hashlib
import os

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


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "configs_full_golden.npz")


def check_against_golden(tag, col, dep, pid):
    """Full-size frame vs the port's golden: depth and primitive ids bit-exact (digest; row sums localise a failure),
    colour within 1/255 per channel on the crops and within the same bound, summed, on every 32x32 block of the frame."""
    g = np.load(GOLDEN)
    Hh, W = dep.shape
    drow = dep.view(np.uint32).astype(np.uint64).sum(axis=1)
    bad = np.nonzero(drow != g[tag + "_depth_rowsum"])[0]
    assert bad.size == 0, "depth differs from the reference's in %d rows, first %d" % (bad.size, bad[0])
    assert hashlib.sha256(dep.tobytes()).digest() == g[tag + "_depth_sha256"].tobytes(), "depth plane digest"
    prow = pid.astype(np.uint64).sum(axis=1)
    bad = np.nonzero(prow != g[tag + "_primid_rowsum"])[0]
    assert bad.size == 0, "primitive ids differ from the reference's in %d rows, first %d" % (bad.size, bad[0])
    assert hashlib.sha256(pid.tobytes()).digest() == g[tag + "_primid_sha256"].tobytes(), "primitive-id plane digest"
    assert int((pid != 0xFFFFFFFF).sum()) == int(g[tag + "_covered"])
    for (x, y, w, h), ref in zip(g[tag + "_crops_xywh"], g[tag + "_crops"]):
        d = np.abs(col[y:y + h, x:x + w, :3].astype(int) - ref.astype(int))
        assert d.max() <= 1, "colour crop at (%d,%d) differs by %d levels" % (x, y, d.max())
        assert (d > 0).any(-1).sum() <= max(8, w * h // 200)
    B = 32
    hb, wb = (Hh + B - 1) // B, (W + B - 1) // B
    pad = np.zeros((hb * B, wb * B, 3), np.int64)
    pad[:Hh, :W] = col[..., :3]
    sums = pad.reshape(hb, B, wb, B, 3).sum(axis=(1, 3))
    dsum = np.abs(sums - g[tag + "_rgb_blocksum"])
    # every pixel may be one level off (1/255 contract) -> a block sum may move by at most B*B; in practice a handful
    assert dsum.max() <= B * B // 8, "block colour sums differ by up to %d" % dsum.max()
    assert (dsum > 0).mean() < 0.2
    return int(dsum.max()), float((dsum > 0).mean())


def full_size_properties(ctx, hana, sc, shader, W, Hh, golden_tag=None):
    u = hana.default_uniforms(W, Hh, True)
    objs = sc.upload(ctx)
    if golden_tag:
        # the two passes by hand through the RenderBuffer path, with the primitive-id plane, against the golden
        frame, shadow = ctx.renderbuffer(W, Hh), ctx.renderbuffer(W, Hh)
        for rb in (frame, shadow):
            rb.clear_color(0, 0, 0, 1)
            rb.clear_depth(FLT_MAX)
        ctx.draw(shadow, objs[0], hana.SHADOW, u)
        pid = ctx.draw(frame, objs[0], shader, u, objs[1], objs[2], shadow, want_primid=True)
        col, dep = frame.download()
        frame.close()
        shadow.close()
        check_against_golden(golden_tag, col, dep, pid)
        del col, dep, pid
    # path A: RenderBuffer read-modify-write, TMA tile load/store
    ctx.set_tma(True)
    colA, depA, stA = rmw_two_pass(ctx, hana, objs, shader, u, W, Hh)
    # path B: batched sweep, clears folded into the flush, R8 shadow map
    sw = ctx.sweep(W, Hh, 1)
    sw.render(objs[0], shader, [u], objs[1], objs[2])
    colB, depB = sw.download(0)
    if golden_tag: # the batched path's depth plane against the golden's digest as well
        g = np.load(GOLDEN)
        assert hashlib.sha256(depB.tobytes()).digest() == g[golden_tag + "_depth_sha256"].tobytes()
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


def test_c4_full_size_vs_golden_and_properties(hana, ctx):
    """configs[3]: 2237 x 2237 vertices -> 9 999 392 triangles at 3840x2160, Blinn + shadow."""
    a2v = hana.scene.synthetic_grid(2237, 2237, seed=1234)
    assert a2v.shape[0] == 9999392 * 3
    dif, nm = hana.scene.noise_textures(1234, 1024, flat_normal=True)
    st, cov = full_size_properties(ctx, hana, hana.Scene("c4", a2v, dif, nm), hana.BLINN, 3840, 2160, golden_tag="c4")
    assert st["faces_in"] == 9999392 and st["tris_out"] > 9_000_000 and cov > 0.9


def test_c5_full_size_vs_golden_and_properties(hana, ctx):
    """configs[4]: 8 layers x 32 x 18 quads x 2 = 9 216 triangles at 7680x4320, NormalMap + shadow."""
    a2v = hana.scene.synthetic_layers(8, 32, 18, seed=99)
    assert a2v.shape[0] == 9216 * 3
    dif, nm = hana.scene.noise_textures(99, 1024)
    st, cov = full_size_properties(ctx, hana, hana.Scene("c5", a2v, dif, nm), hana.NORMALMAP, 7680, 4320, golden_tag="c5")
    assert st["tris_out"] == 9216 and cov > 0.9 and st["tile_refs"] > 1_000_000
