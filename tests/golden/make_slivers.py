#!/usr/bin/env python
"""Generates tests/golden/slivers_posuz.npy: 2-D triangles (NDC x,y of vertices 0,1,2; six floats each) that the
reference KEEPS (NDC signed area > 0, graphics.cpp:172-180) although their screen-space u.z (graphics.cpp:222-229) rounds
to > +0.01 at 1920x1080 — slivers a few ulps wide whose two area computations disagree in sign — and that cover at least
one pixel in the reference (checked with the CPU oracle). Every other kept triangle has u.z < 0; these exercise the
rasteriser's "B and C exchanged" staging (hana_kernels.cuh). Brute force: ~1e8 random slivers, a few minutes.
  python tests/golden/make_slivers.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
f = np.float32
W, H = 1920, 1080


def candidates(want=30000, seed=11):
    rng = np.random.default_rng(seed)
    out, N = [], 1000000
    while len(out) < want:
        a = rng.uniform(-0.95, 0.95, (N, 2)).astype(f)
        ang = rng.uniform(0, 2 * np.pi, N)
        d = np.stack([np.cos(ang), np.sin(ang)], 1).astype(f)
        L = rng.uniform(0.3, 1.5, (N, 1)).astype(f)
        b = (a + d * L).astype(f)
        t = rng.uniform(0.1, 0.9, (N, 1)).astype(f)
        nrm = np.stack([-d[:, 1], d[:, 0]], 1).astype(f)
        c = (a + d * L * t + nrm * rng.uniform(-6e-8, 6e-8, (N, 1)).astype(f)).astype(f)
        inside = (np.abs(b) < 0.999).all(1) & (np.abs(c) < 0.999).all(1)
        n0, n1, n2 = a, b, c
        area = (n0[:, 0] * n1[:, 1] - n0[:, 1] * n1[:, 0]).astype(f)  # is_back_facing, float32 step by step
        area = (area + n1[:, 0] * n2[:, 1]).astype(f)
        area = (area - n1[:, 1] * n2[:, 0]).astype(f)
        area = (area + n2[:, 0] * n0[:, 1]).astype(f)
        area = (area - n2[:, 1] * n0[:, 0]).astype(f)
        sx = [((n[:, 0] + f(1)) * f(0.5) * f(W)).astype(f) for n in (n0, n1, n2)]
        sy = [((n[:, 1] + f(1)) * f(0.5) * f(H)).astype(f) for n in (n0, n1, n2)]
        s0x, s0y = (sx[2] - sx[0]).astype(f), (sx[1] - sx[0]).astype(f)
        s1x, s1y = (sy[2] - sy[0]).astype(f), (sy[1] - sy[0]).astype(f)
        uz = ((s0x * s1y).astype(f) - (s0y * s1x).astype(f)).astype(f)
        for i in np.nonzero((area > 0) & (uz > 0.01) & inside)[0]:
            out.append(np.concatenate([n0[i], n1[i], n2[i]]))
    return np.array(out, f)


def sliver_a2v(tris):
    """a2v stream: identity matrices make object x,y the NDC x,y; per-vertex depth, uv and a +z normal."""
    n = len(tris)
    a2v = np.zeros((n * 3, 8), f)
    z = (np.arange(n, dtype=f) % f(97)) / f(97) - f(0.5)
    uv = np.array([[0.1, 0.2], [0.8, 0.3], [0.4, 0.9]], f)
    for k in range(3):
        a2v[k::3, 0] = tris[:, 2 * k]
        a2v[k::3, 1] = tris[:, 2 * k + 1]
        a2v[k::3, 2] = z + f(0.01) * k
        a2v[k::3, 6:8] = uv[k]
    a2v[:, 5] = 1
    return a2v


def identity_uniforms(hana, w, h):
    u = hana.default_uniforms(w, h, False)
    eye = np.eye(4, dtype=f).reshape(-1)
    for name in ("model", "model_I", "camera_vp", "light_vp"):
        arr = getattr(u, name)
        for k in range(16):
            arr[k] = float(eye[k])
    return u


def main():
    from conftest import load_package
    from oracle import horacle as Hh
    hana = load_package()
    cand = candidates()
    hu = Hh.HanaUniforms.from_bytes(identity_uniforms(hana, W, H).to_bytes())
    col = np.zeros((H, W, 4), np.uint8)
    dep = np.full((H, W), np.float32(3.4028234663852886e38), f)
    pid, cnt = Hh.Port().draw(Hh.GROUND, hu, sliver_a2v(cand), W, H, col, dep, want_primid=True, want_counters=True)
    faces = np.unique(pid[pid != 0xFFFFFFFF] // 8)
    print("%d candidates, %d cover a pixel (%d pixels)" % (len(cand), len(faces), int((pid != 0xFFFFFFFF).sum())))
    np.save(os.path.join(ROOT, "tests", "golden", "slivers_posuz.npy"), cand[faces])


if __name__ == "__main__":
    main()
