"""Host-side logic and the C-ABI surface, without a GPU: every symbol include/hana_b200.h declares is exported,
the library fails loudly (no CPU fallback) when no CUDA device exists, and the C++ host mirror of
Camera / DrawModel::draw's uniform block is bit-identical to the reference's."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ASSET_DIR, ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hana_b200.h")).read()
    return sorted(set(re.findall(r"HANA_API\s+[\w\s\*]*?\b(hana_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(hana):
    out = subprocess.run(["nm", "-D", "--defined-only", hana.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (hana_\w+)", out))
    decl = declared_symbols()
    assert len(decl) >= 50
    missing = [s for s in decl if s not in exported]
    assert not missing, missing
    # and nothing from the checker is linked in
    assert "horacle" not in out and "href_" not in out


def test_sass_is_sm100a_with_tma(hana):
    out = subprocess.run(["cuobjdump", "-lelf", hana.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run("cuobjdump -sass %s | grep -c UTMASTG" % hana.lib_path(), shell=True, capture_output=True, text=True).stdout
    assert int(sass.strip() or 0) > 0  # cp.async.bulk.tensor stores are really there


def test_fails_loudly_without_gpu(hana):
    if hana.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(hana.HanaError) as e:
        hana.Context(0)
    assert e.value.code == -3  # HANA_E_NODEVICE: there is no CPU fallback


def test_uniform_struct_layout(hana, horacle):
    assert C.sizeof(hana.HanaUniforms) == C.sizeof(horacle.HanaUniforms) == 368


needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhana_ref_inst.so"))
                               or not os.path.exists(os.path.join(ASSET_DIR, "diablo3_pose", "diablo3_pose.obj")),
                               reason="oracle/_ref not built")


@needs_ref
def test_host_mirror_uniforms_bit_identical_to_reference(hana, horacle):
    H = horacle
    W, Hh = 1920, 1080
    ref = H.Reference(os.path.join(ASSET_DIR, "diablo3_pose", "diablo3_pose.obj"), W, Hh, H.BLINN)
    cam = hana.OrbitCamera(np.float32(W) / np.float32(Hh))
    arr = hana.orbit_sweep_uniforms(W, Hh, 0, 24, frames_per_turn=1024)
    ref.warmup(True)
    for k in range(24):
        ref.render_time(True)
        want = ref.uniforms().to_bytes()
        assert hana.default_uniforms(W, Hh, True, camera=cam).to_bytes() == want
        assert arr[k].to_bytes() == want
        ref.camera_motion(orbit=(1 / 1024, 0))
        cam.update(orbit=(1 / 1024, 0))
    # general motion + model transform + light
    ref.model_transform(pos=(0.1, -0.2, 0.05), rot_deg=(10, 20, 30), scale=(1.1, 0.9, 1.2))
    ref.light_set((1, 3, 2))
    for motion in (dict(orbit=(0.1, 0.05), pan=(0.01, 0.02), dolly=1.5), dict(orbit=(-0.3, 0.2), pan=(-0.05, 0.0), dolly=-2.0)):
        ref.camera_motion(**motion)
        cam.update(**motion)
        ref.render_time(False)
        d = hana.scene_desc(model_pos=(0.1, -0.2, 0.05), model_rot_deg=(10, 20, 30), model_scale=(1.1, 0.9, 1.2), light_pos=(1, 3, 2))
        assert hana.default_uniforms(W, Hh, False, camera=cam, desc=d).to_bytes() == ref.uniforms().to_bytes()
    ref.close()


def test_orbit_sweep_slices_are_consistent(hana):
    """Sharding by frames: any rank's slice equals the corresponding part of the whole sweep."""
    full = hana.orbit_sweep_uniforms(320, 240, 0, 32, frames_per_turn=32)
    part = hana.orbit_sweep_uniforms(320, 240, 8, 8, frames_per_turn=32)
    for k in range(8):
        assert part[k].to_bytes() == full[8 + k].to_bytes()


def test_frame_checksum_is_order_free_and_sensitive(hana):
    from hana_softwarerenderer_b200.api import frame_checksum
    rng = np.random.RandomState(0)
    col = rng.randint(0, 256, (12, 16, 4)).astype(np.uint8)
    dep = rng.rand(12, 16).astype(np.float32)
    a = frame_checksum(col, dep)
    col2 = col.copy()
    col2[..., 3] = 7  # alpha is excluded (the reference's alpha byte is undefined, App. D5)
    assert frame_checksum(col2, dep) == a
    col2[3, 4, 1] ^= 1
    assert frame_checksum(col2, dep) != a
