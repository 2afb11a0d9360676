"""the data formats either side of the path. OBJ / TGA ingestion and TGA output are host code
(no GPU needed) checked byte for byte against the reference's own codecs (oracle/_ref, built from the unmodified
model.cpp / tgaimage.cpp) and against the packs made with them; the present conversion is a device kernel checked
against a numpy restatement of window_draw_buffer (win32.cpp:348-370)."""
import os

import numpy as np
import pytest

from conftest import ASSET_DIR

NAMES = ("african_head", "diablo3_pose")


def _obj(name, suffix=".obj"):
    return os.path.join(ASSET_DIR, name, name + suffix)


@pytest.fixture(scope="module")
def codecs(horacle):
    if not os.path.exists(horacle.REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return horacle.RefCodecs()


@pytest.mark.parametrize("name", NAMES)
def test_obj_load_matches_reference_model(hana, codecs, name):
    if not os.path.exists(_obj(name)):
        pytest.skip("bundled assets not packed")
    for normal_pass in (1, 3):
        mine = hana.obj_load(_obj(name), normal_pass)
        ref = codecs.obj_a2v(_obj(name), normal_pass)
        assert mine.shape == ref.shape and mine.shape[0] % 3 == 0 and mine.shape[0] > 0
        assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32)), "a2v stream differs from Model's"


@pytest.mark.parametrize("name", NAMES)
def test_obj_load_matches_the_pack_the_gpu_tests_use(hana, name):
    pack = os.path.join(ASSET_DIR, name + ".npz")
    if not (os.path.exists(pack) and os.path.exists(_obj(name))):
        pytest.skip("bundled assets not packed")
    z = np.load(pack)
    # pack_assets.py: one warm-up frame (two passes over the model) and then the export walk
    mine = hana.obj_load(_obj(name), 3)
    assert np.array_equal(mine.view(np.uint32), z["a2v"].astype(np.float32).view(np.uint32))
    for key, suffix in (("diffuse", "_diffuse.tga"), ("normal", "_nm_tangent.tga")):
        tex = hana.tga_load(_obj(name, suffix), model_flip=True)
        assert np.array_equal(tex.reshape(-1), np.asarray(z[key]).reshape(-1)), key


@pytest.mark.parametrize("name", NAMES)
def test_tga_load_matches_reference(hana, codecs, name):
    for suffix in ("_diffuse.tga", "_nm_tangent.tga", "_spec.tga"):
        path = _obj(name, suffix)
        if not os.path.exists(path):
            pytest.skip("bundled assets not packed")
        for flip in (False, True):
            assert np.array_equal(hana.tga_load(path, flip), codecs.tga_read(path, flip)), (suffix, flip)


def _images(rng):
    yield rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)                       # noise: raw packets only
    img = np.zeros((64, 300, 3), np.uint8)                                         # long runs (> 128) and a frame-like mix
    img[10:40, 20:250] = (10, 200, 30)
    img[30:50, 100:120] = rng.integers(0, 256, (20, 20, 3), dtype=np.uint8)
    yield img
    yield rng.integers(0, 3, (40, 129, 1), dtype=np.uint8)                         # grayscale, short runs, odd width
    yield np.repeat(rng.integers(0, 256, (16, 16, 4), dtype=np.uint8), 9, axis=1)  # RGBA, runs of exactly 9
    yield np.full((5, 128, 3), 7, np.uint8)                                        # runs ending exactly at 128
    yield np.full((1, 1, 3), 9, np.uint8)


def test_tga_write_is_byte_identical_and_round_trips(hana, codecs, tmp_path):
    rng = np.random.default_rng(7)
    for k, img in enumerate(_images(rng)):
        for rle in (False, True):
            a, b = str(tmp_path / ("mine_%d_%d.tga" % (k, rle))), str(tmp_path / ("ref_%d_%d.tga" % (k, rle)))
            hana.tga_write(a, img, rle)
            codecs.tga_write(b, img, rle)
            assert open(a, "rb").read() == open(b, "rb").read(), (k, rle)
            assert np.array_equal(hana.tga_load(a, model_flip=False), img)          # top-left origin: no flip on read
            assert np.array_equal(codecs.tga_read(a, False), img)


def test_loaders_report_errors(hana, tmp_path):
    with pytest.raises(hana.HanaError):
        hana.obj_load(str(tmp_path / "missing.obj"))
    bad = tmp_path / "bad.tga"
    bad.write_bytes(b"\x00" * 10)
    with pytest.raises(hana.HanaError):
        hana.tga_load(str(bad))
    quad = tmp_path / "two_corner_face.obj"
    quad.write_text("v 0 0 0\nv 1 0 0\nvt 0 0\nvn 0 0 1\nf 1/1/1 2/1/1\n")
    with pytest.raises(hana.HanaError):
        hana.obj_load(str(quad))
    ok = tmp_path / "quad.obj"   # a 4-corner face: the draw path reads its first three corners (graphics.cpp:381)
    ok.write_text("# c\nv 0 0 0\r\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0.5 0.25\nvn 0 0 2\nf 1/1/1 2/1/1 3/1/1 4/1/1\n")
    a = hana.obj_load(str(ok))
    assert a.shape == (3, 8) and np.allclose(a[2, :3], (1, 1, 0)) and np.allclose(a[0, 3:6], (0, 0, 1)) and np.allclose(a[1, 6:], (0.5, 0.25))


@pytest.mark.gpu
def test_present_matches_window_draw_buffer(hana, ctx):
    for (W, Hh) in ((64, 48), (70, 33)):                                          # W % 4 != 0 exercises the scalar tail
        scene = hana.synthetic_scene("blob")
        model, dtex, ntex = scene.upload(ctx)
        sweep = ctx.sweep(W, Hh, 3)
        us = [hana.default_uniforms(W, Hh, True) for _ in range(3)]
        sweep.render(model, hana.BLINN, us, dtex, ntex, clear_rgba=(9, 8, 7, 1))
        frames = [sweep.download(k)[0] for k in range(3)]
        for fmt, nch in ((hana.PRESENT_BGRA8, 4), (hana.PRESENT_BGR8, 3)):
            got = sweep.present(1, 2, fmt)
            for k in range(2):
                src = frames[1 + k]
                want = np.empty((Hh, W, nch), np.uint8)
                want[..., 0], want[..., 1], want[..., 2] = src[::-1, :, 2], src[::-1, :, 1], src[::-1, :, 0]
                if nch == 4:
                    want[..., 3] = 255
                assert np.array_equal(got[k], want), (W, Hh, fmt, k)
        for o in (sweep, model, dtex, ntex):
            o.close()
