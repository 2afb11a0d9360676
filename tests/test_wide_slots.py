"""The sweep's shadow pass packs {ShadowShader byte << 24 | triangle slot} into one state word per pixel, which addresses
16 777 214 triangles per frame; run_pass routes a pass with a larger per-frame capacity to the wide-slot variant of the
rasteriser (byte beside a full 32-bit slot). HANA_R8_SLOT_LIMIT lowers the threshold so that this test takes that route
with a small scene: the frames must be bit-identical to the packed variant's, including equal-depth ties in the shadow
pass (graphics.cpp:359: the later submission wins), which are resolved through the slot."""
import os

import numpy as np
import pytest

from conftest import compare_frames
from test_gpu_parity import check, oracle_two_pass

pytestmark = pytest.mark.gpu


def test_wide_slot_shadow_pass_matches_packed_and_oracle(hana, horacle, port, ctx, blob):
    W, Hh = 320, 240
    a2v = np.concatenate([blob.a2v, blob.a2v])  # every fragment ties with its copy
    sc = hana.Scene("blob_x2", a2v, blob.diffuse, blob.normal)
    arr = hana.orbit_sweep_uniforms(W, Hh, 3, 4, frames_per_turn=64)

    def render(c):
        objs = sc.upload(c)
        sw = c.sweep(W, Hh, 4)
        sw.render(objs[0], hana.BLINN, arr, objs[1], objs[2])
        frames = [sw.download(f) for f in range(4)]
        for o in (sw,) + tuple(objs):
            o.close()
        return frames

    base = render(ctx)
    n0 = ctx.wide_r8_launches
    os.environ["HANA_R8_SLOT_LIMIT"] = "100"
    try:
        wide_ctx = hana.Context(0)
    finally:
        del os.environ["HANA_R8_SLOT_LIMIT"]
    wide = render(wide_ctx)
    assert wide_ctx.wide_r8_launches > 0 and ctx.wide_r8_launches == n0
    wide_ctx.close()
    for (c0, d0), (c1, d1) in zip(base, wide):
        assert np.array_equal(c0, c1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
    hu = horacle.HanaUniforms.from_bytes(arr[0].to_bytes())
    col, dep, _, _, _ = oracle_two_pass(port, horacle, horacle.BLINN, hu, sc, W, Hh, want_primid=False)
    check(compare_frames(wide[0][0], wide[0][1], col, dep), W * Hh)
