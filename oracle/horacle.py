"""ctypes bindings for the CPU oracle. TEST INFRASTRUCTURE ONLY.

Two checkers live behind this module:

* ``Port``      — oracle/_build/libhana_oracle.so, the C restatement (hana_oracle.c).
* ``Reference`` — oracle/_ref/libhana_ref[_inst].so, the REAL reference compiled from
                  /root/reference by oracle/build_ref.sh (only where it was built; the
                  built files travel to the GPU box, the sources do not).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs may import this. The product path (the hana-softwarerenderer_b200 package) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libhana_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libhana_ref.so")
REF_INST_SO = os.path.join(HERE, "_ref", "libhana_ref_inst.so")
# the reference's scene code linked with THIS repo's graphics_draw_triangle shim instead of graphics.cpp
# (hana-softwarerenderer_b200/host/graphics_dropin.cpp): the object under test of the drop-in parity test
REF_DROPIN_SO = os.path.join(HERE, "_ref", "libhana_ref_dropin.so")
ASSET_DIR = os.path.join(HERE, "_ref", "assets")

SHADOW, BLINN, NORMALMAP, GROUND, TOON, TEXTURE, TEXTURE_LIGHT = range(7)
FLT_MAX = np.float32(3.4028234663852886e38)


class HanaUniforms(C.Structure):
    """include/hana_b200.h: HanaUniforms (ShaderData + Material, IShader.h:7-32)."""

    _fields_ = [
        ("model", C.c_float * 16),
        ("model_I", C.c_float * 16),
        ("camera_vp", C.c_float * 16),
        ("light_vp", C.c_float * 16),
        ("view_pos", C.c_float * 3),
        ("gloss", C.c_float),
        ("light_dir", C.c_float * 3),
        ("bump_scale", C.c_float),
        ("light_color", C.c_float * 4),
        ("ambient", C.c_float * 4),
        ("mat_color", C.c_float * 4),
        ("mat_specular", C.c_float * 4),
        ("enable_shadow", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]

    def copy(self):
        u = HanaUniforms()
        C.memmove(C.byref(u), C.byref(self), C.sizeof(HanaUniforms))
        return u

    def to_bytes(self):
        return bytes(C.string_at(C.byref(self), C.sizeof(HanaUniforms)))

    @staticmethod
    def from_bytes(b):
        u = HanaUniforms()
        C.memmove(C.byref(u), b, C.sizeof(HanaUniforms))
        return u


assert C.sizeof(HanaUniforms) == 368


class _Tex(C.Structure):
    _fields_ = [("data", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("bpp", C.c_int32)]


class _Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("corners", "tris_raster", "bbox_pixels", "inside", "zpass")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def _fptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _tex(arr):
    """arr: uint8 [h, w, bpp] in TGAImage::data layout (B,G,R[,A]) or None."""
    if arr is None:
        return None, None
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    t = _Tex(arr.ctypes.data, arr.shape[1], arr.shape[0], arr.shape[2])
    return t, arr


def build_port():
    subprocess.check_call(["make", "-s", "-C", HERE])


class Port:
    """The C restatement (hana_oracle.c)."""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build_port()
        self.lib = C.CDLL(PORT_SO)
        L = self.lib
        L.horacle_barycentric.restype = C.c_int

    def mat4_mul(self, a, b):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        out = np.zeros(16, np.float32)
        self.lib.horacle_mat4_mul(_fptr(a), _fptr(b), _fptr(out))
        return out

    def vertex(self, shader, u, a2v):
        a2v = np.ascontiguousarray(a2v, np.float32).reshape(-1, 8)
        out = np.zeros((a2v.shape[0], 13), np.float32)
        for i in range(a2v.shape[0]):
            self.lib.horacle_vertex(C.c_int(shader), C.byref(u), _fptr(a2v[i]), _fptr(out[i]))
        return out

    def clip(self, tri39):
        tri39 = np.ascontiguousarray(tri39, np.float32).reshape(39)
        out = np.zeros((10, 13), np.float32)
        n = self.lib.horacle_clip(_fptr(tri39), _fptr(out))
        return out[:n].copy()

    def barycentric(self, abc6, px, py):
        abc6 = np.ascontiguousarray(abc6, np.float32).reshape(6)
        w = np.zeros(3, np.float32)
        ok = self.lib.horacle_barycentric(_fptr(abc6[0:2]), _fptr(abc6[2:4]), _fptr(abc6[4:6]), int(px), int(py), _fptr(w))
        return bool(ok), w

    def fragment(self, shader, u, v2f13, diffuse=None, normal=None, shadow=None):
        v = np.ascontiguousarray(v2f13, np.float32).reshape(13)
        td, _kd = _tex(diffuse)
        tn, _kn = _tex(normal)
        rgb = np.zeros(3, np.float32)
        sp, sw, sh = None, 0, 0
        if shadow is not None:
            shadow = np.ascontiguousarray(shadow, np.uint8)
            sp, sh, sw = _fptr(shadow), shadow.shape[0], shadow.shape[1]
        self.lib.horacle_fragment(C.c_int(shader), C.byref(u), _fptr(v), C.byref(td) if td else None,
                                  C.byref(tn) if tn else None, sp, sw, sh, _fptr(rgb))
        return rgb

    def draw(self, shader, u, a2v, W, H, color, depth, diffuse=None, normal=None, shadow=None, want_primid=False,
             want_counters=False):
        """graphics_draw_triangle over flat arrays; color [H,W,4] u8 and depth [H,W] f32 are updated in place."""
        a2v = np.ascontiguousarray(a2v, np.float32).reshape(-1, 8)
        assert color.dtype == np.uint8 and color.shape == (H, W, 4) and color.flags.c_contiguous
        assert depth.dtype == np.float32 and depth.shape == (H, W) and depth.flags.c_contiguous
        td, _kd = _tex(diffuse)
        tn, _kn = _tex(normal)
        sp, sw, sh = None, 0, 0
        if shadow is not None:
            shadow = np.ascontiguousarray(shadow, np.uint8)
            sp, sh, sw = _fptr(shadow), shadow.shape[0], shadow.shape[1]
        primid = np.full((H, W), 0xFFFFFFFF, np.uint32) if want_primid else None
        cnt = _Counters() if want_counters else None
        self.lib.horacle_draw(C.c_int(shader), C.byref(u), _fptr(a2v), C.c_int(a2v.shape[0]),
                              C.byref(td) if td else None, C.byref(tn) if tn else None, sp, sw, sh, W, H,
                              _fptr(color), _fptr(depth), _fptr(primid) if want_primid else None,
                              C.byref(cnt) if cnt else None)
        return primid, (cnt.as_dict() if cnt else None)

    def draw_model(self, shader, u, a2v, W, H, diffuse=None, normal=None, clear_rgba=(0, 0, 0, 1), clear_depth=FLT_MAX,
                   want_primid=False, want_counters=False):
        """Clear + DrawModel::draw (scene.h:53-99). Returns dict(color, depth, primid, counters)."""
        a2v = np.ascontiguousarray(a2v, np.float32).reshape(-1, 8)
        color = np.empty((H, W, 4), np.uint8)
        color[:] = np.array(clear_rgba, np.uint8)
        depth = np.full((H, W), clear_depth, np.float32)
        scol = np.empty((H, W, 4), np.uint8)
        scol[:] = np.array((0, 0, 0, 1), np.uint8)
        sdep = np.full((H, W), FLT_MAX, np.float32)
        td, _kd = _tex(diffuse)
        tn, _kn = _tex(normal)
        primid = np.full((H, W), 0xFFFFFFFF, np.uint32) if want_primid else None
        cnt = (_Counters * 2)() if want_counters else None
        self.lib.horacle_draw_model(C.c_int(shader), C.byref(u), _fptr(a2v), C.c_int(a2v.shape[0]),
                                    C.byref(td) if td else None, C.byref(tn) if tn else None, W, H, _fptr(color),
                                    _fptr(depth), _fptr(scol), _fptr(sdep), _fptr(primid) if want_primid else None,
                                    cnt)
        return dict(color=color, depth=depth, primid=primid,
                    counters=[cnt[0].as_dict(), cnt[1].as_dict()] if cnt else None)


def reference_available(instrumented=True):
    return os.path.exists(REF_INST_SO if instrumented else REF_SO)


class RefCodecs:
    """The reference's own OBJ / TGA codecs (model.cpp, tgaimage.cpp) through oracle/ref_driver.cpp."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (run oracle/build_ref.sh where /root/reference exists)")
        L = self.lib = C.CDLL(REF_SO)
        L.href_tga_write.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.href_tga_read.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.href_obj_a2v.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int]

    def tga_write(self, path, data, rle):
        a = np.ascontiguousarray(data, np.uint8)
        if a.ndim == 2:
            a = a[:, :, None]
        r = self.lib.href_tga_write(os.fsencode(path), a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], a.shape[2], int(rle))
        if r != 0:
            raise RuntimeError("reference write_tga_file failed")

    def tga_read(self, path, model_flip):
        w, h, b = C.c_int(), C.c_int(), C.c_int()
        if self.lib.href_tga_read(os.fsencode(path), int(model_flip), None, C.byref(w), C.byref(h), C.byref(b)) != 0:
            raise RuntimeError("reference read_tga_file failed")
        out = np.empty((h.value, w.value, b.value), np.uint8)
        self.lib.href_tga_read(os.fsencode(path), int(model_flip), out.ctypes.data_as(C.c_void_p), C.byref(w), C.byref(h), C.byref(b))
        return out

    def obj_a2v(self, path, normal_pass):
        n = self.lib.href_obj_a2v(os.fsencode(path), normal_pass, None, 0)
        out = np.empty((n, 8), np.float32)
        if self.lib.href_obj_a2v(os.fsencode(path), normal_pass, out.ctypes.data_as(C.c_void_p), n) != n:
            raise RuntimeError("reference Model export failed")
        return out


class Reference:
    """The real reference behind oracle/ref_driver.cpp (one scene per instance)."""

    def __init__(self, obj_path, W, H, shader, instrumented=True, dropin=False):
        if dropin:
            instrumented = False
        so = REF_DROPIN_SO if dropin else (REF_INST_SO if instrumented else REF_SO)
        if not os.path.exists(so):
            raise FileNotFoundError(so + " (run oracle/build_ref.sh where /root/reference exists)")
        self.inst = instrumented
        L = self.lib = C.CDLL(so)
        L.href_scene_create.restype = C.c_void_p
        L.href_scene_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.href_render.restype = C.c_double
        L.href_render.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in ("href_scene_destroy", "href_scene_nfaces"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.href_scene_set_shader.argtypes = [C.c_void_p, C.c_int]
        L.href_camera_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.href_camera_motion.argtypes = [C.c_void_p] + [C.c_float] * 5
        L.href_camera_get.argtypes = [C.c_void_p, C.c_void_p]
        L.href_light_set.argtypes = [C.c_void_p, C.c_void_p]
        L.href_material_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.href_model_transform.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.href_get_frame.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.href_get_uniforms.argtypes = [C.c_void_p, C.c_void_p]
        L.href_model_export_a2v.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.href_texture_info.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.href_warmup.argtypes = [C.c_void_p, C.c_int]
        if instrumented:
            L.href_counters_get.argtypes = [C.c_void_p]
            L.href_record_a2v.argtypes = [C.c_void_p]
            L.href_draw_pass.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int, C.c_void_p]
            L.href_set_uniforms.argtypes = [C.c_void_p, C.c_void_p]
            L.href_stage_vertex.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            L.href_stage_fragment.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            L.href_stage_clip.argtypes = [C.c_void_p, C.c_void_p]
            L.href_stage_barycentric.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            L.href_stage_backface.argtypes = [C.c_void_p]
            L.href_stage_interp.argtypes = [C.c_void_p] * 4
            L.href_stage_depth.argtypes = [C.c_void_p] * 2
            L.href_stage_depth.restype = C.c_float
        self.W, self.H = W, H
        self.s = L.href_scene_create(obj_path.encode(), W, H, shader)
        if not self.s:
            raise RuntimeError("href_scene_create failed")
        self.nfaces = L.href_scene_nfaces(self.s)

    def close(self):
        if self.s:
            self.lib.href_scene_destroy(self.s)
            self.s = None

    def set_shader(self, shader):
        self.lib.href_scene_set_shader(self.s, shader)

    def camera_set(self, pos, target=(0, 0, 0)):
        p = np.array(pos, np.float32)
        t = np.array(target, np.float32)
        self.lib.href_camera_set(self.s, _fptr(p), _fptr(t))

    def camera_motion(self, orbit=(0, 0), pan=(0, 0), dolly=0.0):
        self.lib.href_camera_motion(self.s, orbit[0], orbit[1], pan[0], pan[1], dolly)

    def camera_get(self):
        p = np.zeros(3, np.float32)
        self.lib.href_camera_get(self.s, _fptr(p))
        return p

    def light_set(self, pos):
        p = np.array(pos, np.float32)
        self.lib.href_light_set(self.s, _fptr(p))

    def material_set(self, color=(1, 1, 1, 255), specular=(1, 1, 1, 255), gloss=50.0, bump=1.0):
        c = np.array(color, np.float32)
        s = np.array(specular, np.float32)
        self.lib.href_material_set(self.s, _fptr(c), _fptr(s), gloss, bump)

    def model_transform(self, pos=(0, 0, 0), rot_deg=(0, 0, 0), scale=(1, 1, 1)):
        p, r, s = (np.array(x, np.float32) for x in (pos, rot_deg, scale))
        self.lib.href_model_transform(self.s, _fptr(p), _fptr(r), _fptr(s))

    def warmup(self, enable_shadow=True):
        self.lib.href_warmup(self.s, int(enable_shadow))

    def render(self, enable_shadow=True, clear=True):
        """clear + DrawModel::draw. Returns (seconds, color[H,W,4] copy, depth[H,W] copy)."""
        t = self.lib.href_render(self.s, int(enable_shadow), int(clear))
        c, d, w, h = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
        self.lib.href_get_frame(self.s, C.byref(c), C.byref(d), C.byref(w), C.byref(h))
        n = w.value * h.value
        color = np.frombuffer(C.string_at(c.value, n * 4), np.uint8).reshape(h.value, w.value, 4).copy()
        depth = np.frombuffer(C.string_at(d.value, n * 4), np.float32).reshape(h.value, w.value).copy()
        return t, color, depth

    def render_time(self, enable_shadow=True):
        return self.lib.href_render(self.s, int(enable_shadow), 1)

    def uniforms(self):
        u = HanaUniforms()
        self.lib.href_get_uniforms(self.s, C.byref(u))
        return u

    def export_a2v(self):
        n = self.nfaces * 3
        out = np.zeros((n, 8), np.float32)
        r = self.lib.href_model_export_a2v(self.s, _fptr(out), n)
        assert r == n
        return out

    def texture(self, which):
        w, h, bpp, data = C.c_int(), C.c_int(), C.c_int(), C.c_void_p()
        if self.lib.href_texture_info(self.s, which, C.byref(w), C.byref(h), C.byref(bpp), C.byref(data)) != 0:
            return None
        if not data.value or w.value <= 0:
            return None
        n = w.value * h.value * bpp.value
        return np.frombuffer(C.string_at(data.value, n), np.uint8).reshape(h.value, w.value, bpp.value).copy()

    # ---- instrumented-only ------------------------------------------------------
    def counters(self, reset=False):
        out = np.zeros(5, np.uint64)
        self.lib.href_counters_get(_fptr(out))
        if reset:
            self.lib.href_counters_reset()
        return dict(zip(("corners", "tris_raster", "bbox_pixels", "inside", "zpass"), (int(x) for x in out)))

    def record_a2v_next_pass(self):
        self._rec = np.zeros((self.nfaces * 3, 8), np.float32)
        self.lib.href_record_a2v(_fptr(self._rec))
        return self._rec

    def stop_record(self):
        self.lib.href_record_a2v(None)

    def set_uniforms(self, u):
        self.lib.href_set_uniforms(self.s, C.byref(u))

    def draw_pass(self, shader, color, depth, shadow=None, want_primid=True):
        """One graphics_draw_triangle pass into caller buffers (in/out), with the current ShaderData."""
        H, W = depth.shape
        primid = np.full((H, W), 0xFFFFFFFF, np.uint32) if want_primid else None
        sp, sw, sh = None, 0, 0
        if shadow is not None:
            shadow = np.ascontiguousarray(shadow, np.uint8)
            sp, sh, sw = _fptr(shadow), shadow.shape[0], shadow.shape[1]
        self.lib.href_draw_pass(self.s, shader, W, H, _fptr(color), _fptr(depth), sp, sw, sh,
                                _fptr(primid) if want_primid else None)
        return primid

    def stage_vertex(self, shader, a2v):
        a2v = np.ascontiguousarray(a2v, np.float32).reshape(-1, 8)
        out = np.zeros((a2v.shape[0], 13), np.float32)
        for i in range(a2v.shape[0]):
            self.lib.href_stage_vertex(self.s, shader, _fptr(a2v[i]), _fptr(out[i]))
        return out

    def stage_fragment(self, shader, v2f13, shadow=None):
        v = np.ascontiguousarray(v2f13, np.float32).reshape(13)
        rgba = np.zeros(4, np.float32)
        sp, sw, sh = None, 0, 0
        if shadow is not None:
            shadow = np.ascontiguousarray(shadow, np.uint8)
            sp, sh, sw = _fptr(shadow), shadow.shape[0], shadow.shape[1]
        self.lib.href_stage_fragment(self.s, shader, _fptr(v), _fptr(rgba), sp, sw, sh)
        return rgba

    def stage_clip(self, tri39):
        tri39 = np.ascontiguousarray(tri39, np.float32).reshape(39)
        out = np.zeros((10, 13), np.float32)
        n = self.lib.href_stage_clip(_fptr(tri39), _fptr(out))
        return out[:n].copy()

    def stage_barycentric(self, abc6, px, py):
        abc6 = np.ascontiguousarray(abc6, np.float32).reshape(6)
        w = np.zeros(3, np.float32)
        self.lib.href_stage_barycentric(_fptr(abc6), int(px), int(py), _fptr(w))
        return w

    def stage_interp(self, v39, w3, rw3):
        v39 = np.ascontiguousarray(v39, np.float32).reshape(39)
        w3 = np.ascontiguousarray(w3, np.float32)
        rw3 = np.ascontiguousarray(rw3, np.float32)
        out = np.zeros(13, np.float32)
        self.lib.href_stage_interp(_fptr(v39), _fptr(w3), _fptr(rw3), _fptr(out))
        return out
