"""tools/render.py — the headless render-to-TGA driver (replacement of main.cpp:56-161 + tgaimage.cpp:145-246 for batch
output) — against the REAL reference on BASELINE.json configs[0]: the bundled african_head, BlinnShader, 800x600, shadow
off, one frame, the reference's own frame written by its own TGAImage::write_tga_file. File structure identical (header,
footer, size of the decoded image), RGB within 1/255. And the orbit/batch mode: files of a 6-frame orbit, encoded on the
device, byte-identical to the host writer's."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
TOOL = os.path.join(ROOT, "tools", "render.py")
ASSETS = os.path.join(ROOT, "assets")


def run_tool(args):
    r = subprocess.run([sys.executable, TOOL] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_c1_tga_matches_reference_written_tga(hana, horacle, tmp_path):
    obj = os.path.join(ASSETS, "african_head", "african_head.obj")
    if not (os.path.exists(obj) and os.path.exists(horacle.REF_SO)):
        pytest.skip("bundled assets / oracle/_ref missing (needs /root/reference at build time)")
    W, Hh = 800, 600
    out = str(tmp_path / "c1")
    info = run_tool(["--model", "african_head", "--size", "%dx%d" % (W, Hh), "--shader", "blinn", "--no-shadow", "--out", out])
    assert info["files"] == 1 and info["faces"] == 2492
    ref = horacle.Reference(obj, W, Hh, horacle.BLINN, instrumented=False)
    _, col, _ = ref.render(False)  # the first frame the reference would draw: default camera, first walk over the normals
    ref.close()
    ref_path = str(tmp_path / "ref.tga")
    horacle.RefCodecs().tga_write(ref_path, np.ascontiguousarray(col[::-1, :, 2::-1]), True)
    a, b = open(out + ".tga", "rb").read(), open(ref_path, "rb").read()
    assert a[:18] == b[:18] and a[-26:] == b[-26:]
    mine, theirs = hana.tga_load(out + ".tga", model_flip=False), hana.tga_load(ref_path, model_flip=False)
    assert mine.shape == theirs.shape == (Hh, W, 3)
    d = np.abs(mine.astype(int) - theirs.astype(int))
    assert d.max() <= 1, "render.py's TGA differs from the reference's by %d levels" % d.max()
    assert (theirs.reshape(-1, 3).max(1) > 0).mean() > 0.2  # the head is there


def test_orbit_files_identical_to_host_writer(hana, ctx, tmp_path):
    if not os.path.exists(os.path.join(ASSETS, "diablo3_pose", "diablo3_pose.obj")):
        pytest.skip("bundled assets missing")
    W, Hh, F = 640, 360, 6
    out = str(tmp_path / "orbit")
    info = run_tool(["--model", "diablo3_pose", "--size", "%dx%d" % (W, Hh), "--shader", "normalmap", "--shadow", "--frames", str(F),
                     "--orbit-step", str(1.0 / 64), "--batch", "4", "--out", out])
    assert info["files"] == F
    sc = hana.load_bundled("diablo3_pose", ASSETS, 1)
    objs = sc.upload(ctx)
    sw = ctx.sweep(W, Hh, F)
    sw.render(objs[0], hana.NORMALMAP, hana.orbit_sweep_uniforms(W, Hh, 0, F, frames_per_turn=64), objs[1], objs[2])
    for f in range(F):
        col, _ = sw.download(f)
        p = str(tmp_path / ("h%d.tga" % f))
        hana.tga_write(p, np.ascontiguousarray(col[::-1, :, 2::-1]), rle=True)
        assert open(p, "rb").read() == open("%s_%04d.tga" % (out, f), "rb").read(), "frame %d" % f
    for o in (sw,) + tuple(objs):
        o.close()
