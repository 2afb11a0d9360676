#!/usr/bin/env python
"""Generates the committed golden fixtures by running the REAL reference (oracle/_ref/libhana_ref_inst.so,
built by oracle/build_ref.sh from /root/reference). Run here only; the GPU box just reads the .npz files.

  blob_golden.npz      a small procedural scene written as OBJ + TGA, loaded by the reference's own Model /
                       TGAImage, rendered by the reference's DrawModel::draw with each of its 7 shaders
                       (shadow on) at 160x120: inputs as the reference holds them (a2v, textures, uniforms)
                       + outputs (colour, depth, primitive ids).
  bundled_golden.npz   african_head + diablo3_pose: a2v streams, uniforms and the reference's frames for the
                       texture-free shaders (Shadow/Ground/Toon) at 200x150 for three cameras, plus md5 of
                       the full NormalMap+shadow frames at 800x600 / 1920x1080 (SURVEY.md §4 drift detectors).
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package  # noqa: E402
from oracle import horacle as H  # noqa: E402


def write_tga(path, bgr):
    h, w, c = bgr.shape
    hdr = bytearray(18)
    hdr[2] = 2  # uncompressed true-colour
    hdr[12:14] = int(w).to_bytes(2, "little")
    hdr[14:16] = int(h).to_bytes(2, "little")
    hdr[16] = 8 * c
    hdr[17] = 0  # bottom-left origin, like the bundled textures
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(np.ascontiguousarray(bgr).tobytes())


def write_obj(path, a2v):
    with open(path, "w") as f:
        for r in a2v:
            f.write("v %.9g %.9g %.9g\n" % tuple(r[0:3]))
        for r in a2v:
            f.write("vt %.9g %.9g 0\n" % tuple(r[6:8]))
        for r in a2v:
            f.write("vn %.9g %.9g %.9g\n" % tuple(r[3:6]))
        for i in range(0, len(a2v), 3):
            f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % tuple(k for j in range(3) for k in (i + j + 1,) * 3))


def render_all(ref, W, Hh, shaders, cams):
    out = {}
    for ci, cam in enumerate(cams):
        ref.camera_set(cam)
        for sh in shaders:
            ref.set_shader(sh)
            ref.warmup(True)
            rec = ref.record_a2v_next_pass()
            _, col, dep = ref.render(True)
            ref.stop_record()
            key = "c%d_s%d" % (ci, sh)
            out[key + "_color"] = col[..., :3].copy()
            out[key + "_depth"] = dep
            out[key + "_uniforms"] = np.frombuffer(ref.uniforms().to_bytes(), np.uint8).copy()
            out[key + "_a2v"] = rec.copy()  # as that frame's passes saw it (App. A.9)
    return out


def main():
    hana = load_package()
    tmp = tempfile.mkdtemp(prefix="hana_golden_")
    # ---- blob through the reference's own loaders
    sc = hana.synthetic_scene("blob", tex=64)
    d = os.path.join(tmp, "blob")
    os.makedirs(d)
    write_obj(os.path.join(d, "blob.obj"), sc.a2v)
    write_tga(os.path.join(d, "blob_diffuse.tga"), sc.diffuse)
    write_tga(os.path.join(d, "blob_nm_tangent.tga"), sc.normal)
    write_tga(os.path.join(d, "blob_spec.tga"), sc.diffuse)
    W, Hh = 160, 120
    ref = H.Reference(os.path.join(d, "blob.obj"), W, Hh, H.BLINN)
    out = render_all(ref, W, Hh, range(1, 7), [(0, 0, 2), (1.2, 0.8, 1.0)])
    out["diffuse"] = ref.texture(0)
    out["normal"] = ref.texture(1)
    out["size"] = np.array([W, Hh])
    # primitive ids + shadow map of one configuration, via the single-pass entry
    ref.camera_set((0, 0, 2))
    ref.set_shader(H.BLINN)
    ref.render(True)
    a2v = ref.export_a2v()
    scol = np.zeros((Hh, W, 4), np.uint8); scol[..., 3] = 1
    sdep = np.full((Hh, W), H.FLT_MAX, np.float32)
    ref.draw_pass(H.SHADOW, scol, sdep, None, want_primid=False)
    col = np.zeros((Hh, W, 4), np.uint8); col[..., 3] = 1
    dep = np.full((Hh, W), H.FLT_MAX, np.float32)
    pid = ref.draw_pass(H.BLINN, col, dep, scol)
    out.update(pp_a2v=ref.export_a2v(), pp_shadow=scol[..., 0].copy(), pp_color=col[..., :3].copy(), pp_depth=dep, pp_primid=pid,
               pp_uniforms=np.frombuffer(ref.uniforms().to_bytes(), np.uint8).copy())
    ref.close()
    np.savez_compressed(os.path.join(HERE, "blob_golden.npz"), **out)
    print("blob_golden.npz", os.path.getsize(os.path.join(HERE, "blob_golden.npz")))

    # ---- bundled scenes
    out = {}
    assets = os.path.join(ROOT, "oracle", "_ref", "assets")
    for name in ("african_head", "diablo3_pose"):
        obj = os.path.join(assets, name, name + ".obj")
        W, Hh = 200, 150
        ref = H.Reference(obj, W, Hh, H.GROUND)
        r = render_all(ref, W, Hh, (H.GROUND, H.TOON), [(0, 0, 2), (-1.5, 0.5, 1.2), (0.2, 0.1, 0.8)])
        ref.close()
        for k, v in r.items():
            if k.endswith("_a2v") and not k.startswith("c0_s3"):
                continue  # one a2v stream per model is enough for the texture-free shaders (normals differ by ulps only in colour)
            out[name + "_" + k] = v
        for (W2, H2) in ((800, 600), (1920, 1080)):
            ref = H.Reference(obj, W2, H2, H.NORMALMAP)
            _, col, dep = ref.render(True, clear=False)  # first frame, ctor state, no update_camera: SURVEY.md §4
            ref.close()
            md5 = hashlib.md5(col.tobytes() + dep.tobytes()).hexdigest()
            out["%s_md5_%dx%d" % (name, W2, H2)] = np.frombuffer(md5.encode(), np.uint8).copy()
            print(name, W2, H2, md5)
    np.savez_compressed(os.path.join(HERE, "bundled_golden.npz"), **out)
    print("bundled_golden.npz", os.path.getsize(os.path.join(HERE, "bundled_golden.npz")))


if __name__ == "__main__":
    main()
