#!/usr/bin/env python
"""Packs the reference's bundled scenes for use where /root/reference does not exist (the GPU box).

TEST / BENCH INPUT DATA, produced with the REAL reference (oracle/_ref/libhana_ref_inst.so):
  oracle/_ref/assets/<name>.npz        a2v stream exactly as graphics.cpp:380-386 gathers it (after the
                                       in-place normal re-normalisation of one warm-up frame)
                                       + the textures as TGAImage holds them after Model's load (B,G,R rows bottom-up)
  oracle/_ref/assets/<name>/...        the OBJ/TGA files themselves, so that the prebuilt reference library can
                                       load the scene on the GPU box
Nothing is written outside oracle/_ref/ (git-ignored; travels with gpurun).
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import horacle as H  # noqa: E402

SRC = os.environ.get("HANA_REFERENCE", "/root/reference/Hana-SoftwareRenderer") + "/assets"


def main():
    out = os.path.join(HERE, "_ref", "assets")
    os.makedirs(out, exist_ok=True)
    for name in ("african_head", "diablo3_pose"):
        d = os.path.join(out, name)
        os.makedirs(d, exist_ok=True)
        for suffix in (".obj", "_diffuse.tga", "_nm_tangent.tga", "_spec.tga"):
            src = os.path.join(SRC, name, name + suffix)
            dst = os.path.join(d, name + suffix)
            if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src):
                shutil.copyfile(src, dst)
        ref = H.Reference(os.path.join(d, name + ".obj"), 64, 48, H.BLINN)
        ref.warmup(True)
        a2v = ref.export_a2v()
        np.savez_compressed(os.path.join(out, name + ".npz"), a2v=a2v, diffuse=ref.texture(0), normal=ref.texture(1))
        ref.close()
        print("packed", name, a2v.shape)


if __name__ == "__main__":
    main()
