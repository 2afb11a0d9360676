"""The CUDA kernels' arithmetic (csrc/hana_core.cuh) compiled for the host and walked with the kernels'
control flow (tests/emu/emu_render.cpp), against the CPU oracle — no GPU needed. Proves on CPU:
the division-free coverage test == the reference's barycentric() test, the order-free (min depth, max key)
resolve == the reference's in-order depth test, deferred shading of the winner == the reference's colours."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import cleared

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    return C.CDLL(os.path.join(HERE, "emu", "_build", "libhana_emu.so"))


def pack_tex(t):
    if t is None:
        return None
    out = np.zeros(t.shape[:2], np.uint32)
    for i in range(t.shape[2]):
        out |= t[..., i].astype(np.uint32) << np.uint32(8 * i)
    return np.ascontiguousarray(out)


def emu_draw(emu, shader, u, a2v, W, Hh, color, depth, dif=None, nm=None, shadow=None):
    a2v = np.ascontiguousarray(a2v, np.float32)
    primid = np.full((Hh, W), 0xFFFFFFFF, np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None  # noqa: E731
    sw = sh = sp = 0
    if shadow is not None:
        sh, sw = shadow.shape[:2]
        sp = sw * 4
    emu.emu_draw(shader, C.byref(u), p(a2v), a2v.shape[0], p(dif), dif.shape[1] if dif is not None else 0,
                 dif.shape[0] if dif is not None else 0, p(nm), nm.shape[1] if nm is not None else 0,
                 nm.shape[0] if nm is not None else 0, p(shadow), sw, sh, sp, 4, W, Hh, p(color), p(depth), p(primid), None)
    return primid


@pytest.mark.parametrize("shader", [1, 2, 3, 4, 5, 6])
def test_emulated_kernels_match_oracle(emu, port, horacle, hana, blob, shader):
    W, Hh = 200, 150
    for pos in ((0, 0, 2), (0.4, 0.3, 0.7)):  # the second camera sits inside the model's bounding sphere: clipping
        cam = hana.OrbitCamera(np.float32(W) / np.float32(Hh), position=pos)
        u = horacle.HanaUniforms.from_bytes(hana.default_uniforms(W, Hh, True, camera=cam).to_bytes())
        scol, sdep = cleared(W, Hh)
        port.draw(horacle.SHADOW, u, blob.a2v, W, Hh, scol, sdep)
        ecol, edep = cleared(W, Hh)
        emu_draw(emu, horacle.SHADOW, u, blob.a2v, W, Hh, ecol, edep)
        assert np.array_equal(scol, ecol) and np.array_equal(sdep.view(np.uint32), edep.view(np.uint32))
        col, dep = cleared(W, Hh)
        pid, _ = port.draw(shader, u, blob.a2v, W, Hh, col, dep, diffuse=blob.diffuse, normal=blob.normal, shadow=scol,
                           want_primid=True)
        col2, dep2 = cleared(W, Hh)
        pid2 = emu_draw(emu, shader, u, blob.a2v, W, Hh, col2, dep2, pack_tex(blob.diffuse), pack_tex(blob.normal), ecol)
        assert np.array_equal(pid, pid2)
        assert np.array_equal(dep.view(np.uint32), dep2.view(np.uint32))
        assert np.abs(col.astype(int) - col2.astype(int)).max() <= 1
        assert (col != col2).any(-1).mean() < 1e-3


def test_division_free_coverage_equals_barycentric(emu, port):
    """Adversarial edges: integer-aligned vertices (pixels exactly on edges), shared edges, slivers, both windings."""
    rng = np.random.RandomState(5)
    w3 = np.zeros(3, np.float32)
    n = 0
    for it in range(3000):
        abc = rng.uniform(0, 24, 6).astype(np.float32)
        mode = it % 4
        if mode == 0:
            abc = np.round(abc)
        elif mode == 1:
            abc = np.round(abc * 2) / 2
        elif mode == 2:
            abc[4:6] = abc[0:2] + (abc[2:4] - abc[0:2]) * np.float32(rng.rand()) + rng.uniform(-1e-3, 1e-3, 2).astype(np.float32)
        for (px, py) in rng.randint(0, 24, (12, 2)):
            ok, w = port.barycentric(abc, int(px), int(py))
            got = emu.emu_coverage(abc.ctypes.data_as(C.c_void_p), 24, 24, int(px), int(py), w3.ctypes.data_as(C.c_void_p))
            assert bool(got) == ok, (abc, px, py, w, w3)
            if ok:
                assert np.array_equal(w.view(np.uint32), w3.view(np.uint32))
                n += 1
    assert n > 1000


def test_uniform_prepare_matches_oracle_products(emu, port, hana, horacle):
    from hana_softwarerenderer_b200.api import HanaUniforms
    u = hana.default_uniforms(640, 480, True)
    dev = (C.c_float * 200)()
    emu.emu_prepare(C.byref(u), dev)
    mvp = port.mat4_mul(np.array(list(u.camera_vp), np.float32), np.array(list(u.model), np.float32))
    lmvp = port.mat4_mul(np.array(list(u.light_vp), np.float32), np.array(list(u.model), np.float32))
    got = np.array(list(dev), np.float32)
    assert np.array_equal(got[:16].view(np.uint32), mvp.view(np.uint32))
    assert np.array_equal(got[16:32].view(np.uint32), lmvp.view(np.uint32))
    assert HanaUniforms is not None


def test_binning_tile_test_never_drops_a_covered_tile(emu):
    """tile_may_touch (hana_core.cuh) decides which tile lists a triangle enters. Against exhaustive coverage of every
    tile a triangle's bounding box meets: it may keep an empty tile, it must never reject one with an inside pixel —
    random, integer-aligned (edges and vertices exactly on pixels / tile borders), sliver and huge triangles."""
    rng = np.random.default_rng(7)
    emu.emu_tile_touch.restype = C.c_int
    cases = []
    for _ in range(1500):
        c = rng.uniform(0, 256, 2)
        cases.append((c + rng.normal(0, rng.choice([2.0, 12.0, 60.0]), (3, 2))).astype(np.float32))
    for _ in range(500):  # integer / tile-aligned vertices
        cases.append(rng.integers(0, 17, (3, 2)).astype(np.float32) * rng.choice([1.0, 8.0, 16.0]))
    for _ in range(300):  # slivers
        a = rng.uniform(0, 256, 2)
        d = rng.normal(0, 80, 2)
        cases.append(np.stack([a, a + d, a + 0.5 * d + rng.normal(0, 0.05, 2)]).astype(np.float32))
    for _ in range(100):  # much larger than the window
        cases.append(rng.uniform(-4000, 4000, (3, 2)).astype(np.float32))
    kept = rejected = covered = 0
    for tri in cases:
        abc = np.ascontiguousarray(tri.reshape(-1))
        for order in (abc, abc[[0, 1, 4, 5, 2, 3]]):  # both windings: u.z < 0 is the tested path, u.z > 0 always keeps
            order = np.ascontiguousarray(order)
            x0, y0 = np.floor(order.reshape(3, 2).min(0) / 16).astype(int) * 16
            x1, y1 = np.floor(order.reshape(3, 2).max(0) / 16).astype(int) * 16
            xs = range(max(x0, -64), min(x1, 320) + 1, 16)
            ys = range(max(y0, -64), min(y1, 320) + 1, 16)
            for X0 in xs:
                for Y0 in ys:
                    r = emu.emu_tile_touch(order.ctypes.data_as(C.c_void_p), int(X0), int(Y0))
                    assert r != 2, (tri, X0, Y0)
                    kept += r & 1
                    rejected += 1 - (r & 1)
                    covered += r >> 1
    assert rejected > 0 and covered > 0 and kept >= covered
