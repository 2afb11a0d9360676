"""Pins the CPU oracle (oracle/hana_oracle.c). The reference has no tests or golden vectors of its own
(SURVEY.md §4), so the pin is the reference ITSELF: (1) committed fixtures produced by the real reference
(tests/golden/make_golden.py), checked everywhere; (2) where oracle/_ref/ was built from /root/reference,
live bit-exact comparisons of whole frames, primitive ids and every stage function."""
import hashlib
import os

import numpy as np
import pytest

from conftest import ASSET_DIR, cleared

HERE = os.path.dirname(os.path.abspath(__file__))
FLT_MAX = np.float32(3.4028234663852886e38)


def two_pass(port, H, shader, u, a2v, W, Hh, dif, nm):
    scol, sdep = cleared(W, Hh)
    port.draw(H.SHADOW, u, a2v, W, Hh, scol, sdep)
    col, dep = cleared(W, Hh)
    pid, _ = port.draw(shader, u, a2v, W, Hh, col, dep, diffuse=dif, normal=nm, shadow=scol, want_primid=True)
    return col, dep, pid, scol


@pytest.fixture(scope="module")
def blob_golden():
    return np.load(os.path.join(HERE, "golden", "blob_golden.npz"))


@pytest.fixture(scope="module")
def bundled_golden():
    return np.load(os.path.join(HERE, "golden", "bundled_golden.npz"))


@pytest.mark.parametrize("cam", [0, 1])
@pytest.mark.parametrize("shader", [1, 2, 3, 4, 5, 6])
def test_port_matches_reference_fixture_blob(port, horacle, blob_golden, shader, cam):
    g = blob_golden
    W, Hh = (int(x) for x in g["size"])
    k = "c%d_s%d" % (cam, shader)
    u = horacle.HanaUniforms.from_bytes(g[k + "_uniforms"].tobytes())
    col, dep, _, _ = two_pass(port, horacle, shader, u, g[k + "_a2v"], W, Hh, g["diffuse"], g["normal"])
    assert np.array_equal(dep.view(np.uint32), g[k + "_depth"].view(np.uint32))
    assert np.array_equal(col[..., :3], g[k + "_color"])


def test_port_primid_and_shadow_map_fixture(port, horacle, blob_golden):
    g = blob_golden
    W, Hh = (int(x) for x in g["size"])
    u = horacle.HanaUniforms.from_bytes(g["pp_uniforms"].tobytes())
    # pp_a2v was exported AFTER the passes; normals drift by ulps per access (App. A.9), which cannot move
    # positions, coverage, depth or ids, and moves colour by at most one level
    col, dep, pid, scol = two_pass(port, horacle, horacle.BLINN, u, g["pp_a2v"], W, Hh, g["diffuse"], g["normal"])
    assert np.array_equal(pid, g["pp_primid"])
    assert np.array_equal(dep.view(np.uint32), g["pp_depth"].view(np.uint32))
    assert np.array_equal(scol[..., 0], g["pp_shadow"])
    assert np.abs(col[..., :3].astype(int) - g["pp_color"].astype(int)).max() <= 1


@pytest.mark.parametrize("name", ["african_head", "diablo3_pose"])
def test_port_matches_reference_fixture_bundled(port, horacle, bundled_golden, name):
    g = bundled_golden
    W, Hh = 200, 150
    a2v = g[name + "_c0_s3_a2v"]
    for cam in range(3):
        for shader in (horacle.GROUND, horacle.TOON):
            k = "%s_c%d_s%d" % (name, cam, shader)
            u = horacle.HanaUniforms.from_bytes(g[k + "_uniforms"].tobytes())
            col, dep, _, _ = two_pass(port, horacle, shader, u, a2v, W, Hh, None, None)
            assert np.array_equal(dep.view(np.uint32), g[k + "_depth"].view(np.uint32)), k
            assert np.abs(col[..., :3].astype(int) - g[k + "_color"].astype(int)).max() <= 1, k
            assert (col[..., :3] != g[k + "_color"]).any(-1).mean() < 1e-3, k


# ---- live comparisons against the real reference (only where it was built) ----------------------------
needs_ref = pytest.mark.skipif(
    not (os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libhana_ref_inst.so"))
         and os.path.exists(os.path.join(ASSET_DIR, "african_head", "african_head.obj"))),
    reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("name,size", [("african_head", (800, 600)), ("diablo3_pose", (640, 360))])
def test_live_frames_bit_exact(port, horacle, bundled_golden, name, size):
    H = horacle
    W, Hh = size
    ref = H.Reference(os.path.join(ASSET_DIR, name, name + ".obj"), W, Hh, H.NORMALMAP)
    dif, nm = ref.texture(0), ref.texture(1)
    for shader in (H.NORMALMAP, H.BLINN, H.TEXTURE_LIGHT):
        ref.set_shader(shader)
        for step in range(2):
            ref.camera_motion(orbit=(0.13, 0.02), dolly=0.7)
            ref.warmup(True)
            rec = ref.record_a2v_next_pass()
            _, col, dep = ref.render(True)
            ref.stop_record()
            u = ref.uniforms()
            # the recorder holds the normals of the LAST pass (main); the shadow pass ignores normals
            pcol, pdep, _, _ = two_pass(port, H, shader, u, rec, W, Hh, dif, nm)
            assert np.array_equal(pdep.view(np.uint32), dep.view(np.uint32))
            assert np.array_equal(pcol[..., :3], col[..., :3])
    ref.close()


@needs_ref
def test_known_answer_md5(horacle, bundled_golden):
    """SURVEY.md §4 drift detectors: first frame, ctor state, NormalMap + shadow."""
    H = horacle
    want = {("african_head", 800, 600): "883c701334f4694fdc884e57f5ec48bf",
            ("diablo3_pose", 800, 600): "f838e6d9cb5e732f46e5be90115d0755"}
    for (name, W, Hh), md5 in want.items():
        assert bundled_golden["%s_md5_%dx%d" % (name, W, Hh)].tobytes().decode() == md5
        ref = H.Reference(os.path.join(ASSET_DIR, name, name + ".obj"), W, Hh, H.NORMALMAP)
        _, col, dep = ref.render(True, clear=False)
        ref.close()
        assert hashlib.md5(col.tobytes() + dep.tobytes()).hexdigest() == md5


@needs_ref
def test_live_stage_functions(port, horacle):
    H = horacle
    rng = np.random.RandomState(11)
    ref = H.Reference(os.path.join(ASSET_DIR, "african_head", "african_head.obj"), 320, 240, H.BLINN)
    ref.model_transform(pos=(0.1, -0.05, 0.2), rot_deg=(10, 25, -5), scale=(1.1, 0.9, 1.0))
    ref.render(True)
    u = ref.uniforms()
    dif, nm = ref.texture(0), ref.texture(1)
    a2v = ref.export_a2v()[:600]
    for shader in range(7):
        got = port.vertex(shader, u, a2v)
        want = ref.stage_vertex(shader, a2v)
        used = {0: [0, 1, 2, 3], 3: [0, 1, 2, 3, 12], 4: [0, 1, 2, 3, 12], 5: [0, 1, 2, 3, 10, 11],
                6: [0, 1, 2, 3, 7, 8, 9, 10, 11]}.get(shader, list(range(12)))
        assert np.array_equal(got[:, used].view(np.uint32), want[:, used].view(np.uint32)), shader
    # clip: random triangles straddling the planes
    for _ in range(400):
        tri = rng.uniform(-1.5, 1.5, (3, 13)).astype(np.float32)
        tri[:, 3] = rng.uniform(-0.5, 1.5, 3)
        a, b = port.clip(tri), ref.stage_clip(tri)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # barycentric incl. pixels exactly on edges and slivers
    for _ in range(400):
        abc = rng.uniform(0, 32, 6).astype(np.float32)
        if rng.rand() < 0.3:
            abc = np.round(abc)
        for (px, py) in rng.randint(0, 32, (16, 2)):
            ok, w = port.barycentric(abc, px, py)
            wr = ref.stage_barycentric(abc, px, py)
            assert np.array_equal(w.view(np.uint32), wr.view(np.uint32))
            assert ok == (not (wr < 0).any())
    # fragment shaders on random varyings, with a random shadow map
    shadow = rng.randint(0, 256, (240, 320, 4)).astype(np.uint8)
    for shader in range(7):
        for _ in range(200):
            v = rng.uniform(-1, 1, 13).astype(np.float32)
            v[10:12] = rng.uniform(-0.1, 1.1, 2)
            a = port.fragment(shader, u, v, dif, nm, shadow)
            b = ref.stage_fragment(shader, v, shadow)[:3]
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (shader, a, b)
    ref.close()
