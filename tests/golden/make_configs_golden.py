#!/usr/bin/env python
"""Golden vectors of BASELINE.json configs[3] and configs[4] AT FULL SIZE, produced on the CPU box by the C port of
the reference (oracle/hana_oracle.c, pinned bit-exact to the real reference by tests/test_oracle_vs_reference.py;
where oracle/_ref exists the 4K frame is additionally rendered by the unmodified reference from an OBJ file when
--check-reference is given).

  configs[3]: synthetic_grid(2237, 2237, seed 1234) = 9 999 392 triangles, Blinn + shadow, 3840x2160
  configs[4]: synthetic_layers(8, 32, 18, seed 99) = 9 216 triangles, NormalMap + shadow, 7680x4320

Written to tests/golden/configs_full_golden.npz (small: digests, counts, block sums, a few crops):
  <c>_depth_sha256      sha256 of the float32 depth plane (bit-exact contract)
  <c>_primid_sha256     sha256 of the uint32 primitive-id plane (face*8 + fan index, 0xFFFFFFFF = nothing drawn)
  <c>_covered           pixels drawn
  <c>_zpass             [2] fragments the reference shades (depth test passed) in the shadow pass and in the main pass:
                        N_frag of SURVEY.md §8(d)'s algorithmic-byte formula (bench.py reads it)
  <c>_depth_rowsum      per-row sum of depth bits as uint64 (localises a mismatch to rows)
  <c>_primid_rowsum     per-row sum of primitive ids as uint64
  <c>_rgb_blocksum      [H/32, W/32, 3] int64 sums of R, G, B over 32x32 pixel blocks (whole-frame colour, +-1/255 per pixel)
  <c>_crops_xywh, <c>_crops   three 256x256 RGB crops (per-pixel colour check, <= 1/255)

  python tests/golden/make_configs_golden.py            (about 2-4 minutes of one core)
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import horacle as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "configs_full_golden.npz")
BLOCK = 32


def summarise(tag, r, W, Hh, crops_xy):
    depth, primid, color = r["depth"], r["primid"], r["color"]
    out = {}
    out[tag + "_depth_sha256"] = np.frombuffer(hashlib.sha256(depth.tobytes()).digest(), np.uint8)
    out[tag + "_primid_sha256"] = np.frombuffer(hashlib.sha256(primid.tobytes()).digest(), np.uint8)
    out[tag + "_covered"] = np.int64((primid != 0xFFFFFFFF).sum())
    out[tag + "_zpass"] = np.array([r["counters"][0]["zpass"], r["counters"][1]["zpass"]], np.int64)
    out[tag + "_depth_rowsum"] = depth.view(np.uint32).astype(np.uint64).sum(axis=1)
    out[tag + "_primid_rowsum"] = primid.astype(np.uint64).sum(axis=1)
    hb, wb = (Hh + BLOCK - 1) // BLOCK, (W + BLOCK - 1) // BLOCK
    pad = np.zeros((hb * BLOCK, wb * BLOCK, 3), np.int64)
    pad[:Hh, :W] = color[..., :3]
    out[tag + "_rgb_blocksum"] = pad.reshape(hb, BLOCK, wb, BLOCK, 3).sum(axis=(1, 3))
    out[tag + "_crops_xywh"] = np.array([(x, y, 256, 256) for x, y in crops_xy], np.int32)
    out[tag + "_crops"] = np.stack([color[y:y + 256, x:x + 256, :3] for x, y in crops_xy])
    return out


def main():
    ge.build()
    hana = ge.load_package()
    port = H.Port()
    out = {}
    # configs[3]
    W, Hh = 3840, 2160
    t0 = time.time()
    a2v = hana.scene.synthetic_grid(2237, 2237, seed=1234)
    dif, nm = hana.scene.noise_textures(1234, 1024, flat_normal=True)
    u = H.HanaUniforms.from_bytes(hana.default_uniforms(W, Hh, True).to_bytes())
    r = port.draw_model(H.BLINN, u, a2v, W, Hh, diffuse=dif, normal=nm, want_primid=True, want_counters=True)
    print("configs[3] rendered by the port in %.1f s, covered %d" % (time.time() - t0, (r["primid"] != 0xFFFFFFFF).sum()))
    out.update(summarise("c4", r, W, Hh, [(0, 0), (1792, 952), (3584, 1904)]))
    del a2v, r
    # configs[4]
    W, Hh = 7680, 4320
    t0 = time.time()
    a2v = hana.scene.synthetic_layers(8, 32, 18, seed=99)
    dif, nm = hana.scene.noise_textures(99, 1024)
    u = H.HanaUniforms.from_bytes(hana.default_uniforms(W, Hh, True).to_bytes())
    r = port.draw_model(H.NORMALMAP, u, a2v, W, Hh, diffuse=dif, normal=nm, want_primid=True, want_counters=True)
    print("configs[4] rendered by the port in %.1f s, covered %d" % (time.time() - t0, (r["primid"] != 0xFFFFFFFF).sum()))
    out.update(summarise("c5", r, W, Hh, [(64, 64), (3712, 2032), (7360, 4000)]))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
