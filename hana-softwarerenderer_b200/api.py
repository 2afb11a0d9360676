"""ctypes bindings over the C ABI in include/hana_b200.h (libhana_b200.so)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SHADOW, BLINN, NORMALMAP, GROUND, TOON, TEXTURE, TEXTURE_LIGHT = range(7)
FLT_MAX = np.float32(3.4028234663852886e38)
PROF_NAMES = ("begin", "setup", "scan", "fill", "raster_shadow", "raster_main", "other")


class HanaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("hana_b200 error %d: %s" % (code, msg))
        self.code = code


class HanaUniforms(C.Structure):
    """HanaUniforms of include/hana_b200.h (ShaderData + Material, IShader.h:7-32)."""

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
        C.memmove(C.byref(u), bytes(b), C.sizeof(HanaUniforms))
        return u


assert C.sizeof(HanaUniforms) == 368


class HanaStats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("faces_in", "tris_clipped", "tris_out", "tile_refs", "tiles_touched",
                                          "pixels_covered", "overflow", "reserved")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def lib_path():
    return os.path.join(HERE, "libhana_b200.so")


def build(force=False):
    """Compile csrc/ for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(HERE, "csrc")
    if force and os.path.exists(lib_path()):
        os.remove(lib_path())
    subprocess.check_call(["make", "-s", "-C", src])
    return lib_path()


def load():
    """Load libhana_b200.so. Raises if it has not been built: there is no fallback path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise HanaError(-3, "libhana_b200.so is missing (%s): build it with __graft_entry__.build(); "
                            "there is no CPU fallback" % p)
    L = C.CDLL(p)
    L.hana_last_error.restype = C.c_char_p
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    sig = {
        "hana_ctx_create": [i, C.POINTER(vp)],
        "hana_ctx_destroy": [vp],
        "hana_ctx_set_stream": [vp, vp],
        "hana_sync": [vp],
        "hana_ctx_launch_count": [vp, C.POINTER(C.c_uint64)],
        "hana_ctx_uses_tma": [vp],
        "hana_ctx_set_tma": [vp, i],
        "hana_ctx_set_pipeline": [vp, i],
        "hana_ctx_sm_count": [vp],
        "hana_ctx_wide_r8_launches": [vp, C.POINTER(C.c_uint64)],
        "hana_timer_start": [vp],
        "hana_timer_stop": [vp, C.POINTER(f)],
        "hana_ctx_profile": [vp, i],
        "hana_ctx_profile_reset": [vp],
        "hana_ctx_profile_get": [vp, i, C.POINTER(C.c_double), C.POINTER(C.c_uint64)],
        "hana_model_upload": [vp, vp, i, C.POINTER(vp)],
        "hana_model_destroy": [vp],
        "hana_model_ncorners": [vp],
        "hana_texture_upload": [vp, vp, i, i, i, C.POINTER(vp)],
        "hana_texture_destroy": [vp],
        "hana_rb_create": [vp, i, i, C.POINTER(vp)],
        "hana_rb_destroy": [vp],
        "hana_rb_size": [vp, C.POINTER(i), C.POINTER(i)],
        "hana_rb_clear_color": [vp, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint8],
        "hana_rb_clear_depth": [vp, f],
        "hana_rb_upload": [vp, vp, vp],
        "hana_rb_download": [vp, vp, vp],
        "hana_rb_device_ptrs": [vp, C.POINTER(vp), C.POINTER(vp)],
        "hana_draw": [vp, vp, vp, i, vp, vp, vp, vp],
        "hana_draw_model": [vp, vp, vp, vp, i, vp, vp, vp],
        "hana_draw_model_host": [vp, i, i, vp, vp, vp, i, vp, vp, vp, i, vp, f],
        "hana_last_stats": [vp, vp],
        "hana_sweep_create": [vp, i, i, i, C.POINTER(vp)],
        "hana_sweep_destroy": [vp],
        "hana_sweep_render": [vp, vp, i, vp, i, vp, vp, vp, f],
        "hana_sweep_render_dev": [vp, vp, i, vp, i, i, vp, vp, vp, f],
        "hana_sweep_uniforms_dev": [vp, C.POINTER(vp)],
        "hana_sweep_download": [vp, i, vp, vp],
        "hana_sweep_download_async": [vp, i, i, vp, vp],
        "hana_sweep_device_ptrs": [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)],
        "hana_sweep_checksums": [vp, i, vp],
        "hana_sweep_stats": [vp, i, vp],
        "hana_sweep_overflow_count": [vp, C.POINTER(C.c_uint64)],
        "hana_sweep_present": [vp, i, i, i, vp, C.POINTER(vp)],
        "hana_sweep_encode_tga": [vp, i, i],
        "hana_sweep_fetch_tga": [vp, vp, C.c_size_t, vp, vp],
        "hana_sweep_set_bands": [vp, i, i, i, i],
        "hana_sweep_set_shadow_reuse": [vp, i],
        "hana_sweep_render_pass": [vp, i, vp, i, vp, i, vp, vp, vp, f],
        "hana_sweep_shadow_ptrs": [vp, C.POINTER(vp), C.POINTER(i), C.POINTER(C.c_size_t)],
        "hana_sweep_render_pass_async": [vp, i, vp, i, vp, i, vp, vp, vp, f],
        "hana_sweep_passes_ok": [vp, C.POINTER(i)],
        "hana_tga_write": [C.c_char_p, vp, i, i, i, i],
        "hana_obj_load": [C.c_char_p, i, C.POINTER(vp), C.POINTER(i)],
        "hana_tga_load": [C.c_char_p, i, C.POINTER(vp), C.POINTER(i), C.POINTER(i), C.POINTER(i)],
        "hana_host_alloc": [C.c_size_t, C.POINTER(vp)],
        "hana_host_free": [vp],
        "hana_stage_vertex": [vp, vp, i, vp, vp],
        "hana_stage_setup": [vp, vp, i, vp, i, i, i, vp, vp, C.POINTER(i)],
        "hana_draw_primid": [vp, vp, vp, i, vp, vp, vp, vp, vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.hana_free.argtypes = [vp]
    L.hana_free.restype = None
    _LIB = L
    return L


def _ck(code):
    if code != 0:
        raise HanaError(code, load().hana_last_error().decode(errors="replace"))


def device_count():
    return load().hana_device_count()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _h(obj):
    return obj.h if obj is not None else None


class Context:
    """One per GPU. Calls on one context must be serialised by the caller."""

    def __init__(self, device=0):
        self.L = load()
        h = C.c_void_p()
        _ck(self.L.hana_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.L.hana_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        _ck(self.L.hana_sync(self.h))

    def set_stream(self, cuda_stream_ptr):
        _ck(self.L.hana_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    @property
    def launches(self):
        n = C.c_uint64()
        _ck(self.L.hana_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    @property
    def uses_tma(self):
        return bool(self.L.hana_ctx_uses_tma(self.h))

    def set_tma(self, enable):
        _ck(self.L.hana_ctx_set_tma(self.h, int(enable)))

    def set_pipeline(self, enable):
        """Pipelined sweep submissions on / off (on by default); waits for the work queued so far."""
        _ck(self.L.hana_ctx_set_pipeline(self.h, int(bool(enable))))

    @property
    def sm_count(self):
        return self.L.hana_ctx_sm_count(self.h)

    @property
    def wide_r8_launches(self):
        n = C.c_uint64()
        _ck(self.L.hana_ctx_wide_r8_launches(self.h, C.byref(n)))
        return int(n.value)

    def timer_start(self):
        _ck(self.L.hana_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        _ck(self.L.hana_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile(self, enable=True, reset=True):
        _ck(self.L.hana_ctx_profile(self.h, int(enable)))
        if reset:
            _ck(self.L.hana_ctx_profile_reset(self.h))

    def profile_get(self):
        out = {}
        for k, name in enumerate(PROF_NAMES):
            ms, n = C.c_double(), C.c_uint64()
            _ck(self.L.hana_ctx_profile_get(self.h, k, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    # --- inputs
    def model(self, a2v):
        return Model(self, a2v)

    def texture(self, tga_data):
        return Texture(self, tga_data) if tga_data is not None else None

    def renderbuffer(self, w, h):
        return RenderBuffer(self, w, h)

    def sweep(self, w, h, max_frames):
        return Sweep(self, w, h, max_frames)

    # --- draws
    def draw(self, rb, model, shader, uniforms, diffuse=None, normal=None, shadow_map=None, want_primid=False):
        """graphics_draw_triangle(DrawData*) (graphics.cpp:378-407): one pass into `rb`."""
        if want_primid:
            pid = np.empty((rb.height, rb.width), np.uint32)
            _ck(self.L.hana_draw_primid(self.h, rb.h, model.h, shader, C.byref(uniforms), _h(diffuse), _h(normal),
                                        _h(shadow_map), _ptr(pid)))
            return pid
        _ck(self.L.hana_draw(self.h, rb.h, model.h, shader, C.byref(uniforms), _h(diffuse), _h(normal), _h(shadow_map)))
        return None

    def draw_model(self, frame, shadow_map, model, shader, uniforms, diffuse=None, normal=None):
        """DrawModel::draw (scene.h:53-99)."""
        _ck(self.L.hana_draw_model(self.h, frame.h, _h(shadow_map), model.h, shader, C.byref(uniforms), _h(diffuse),
                                   _h(normal)))

    def draw_model_host(self, color, depth, model, shader, uniforms, diffuse=None, normal=None, assume_cleared=False,
                        clear_rgba=(0, 0, 0, 1), clear_depth=FLT_MAX):
        """Same with HOST buffers (numpy, updated in place): the drop-in shim's call."""
        Hh, W = depth.shape
        assert color.shape == (Hh, W, 4) and color.dtype == np.uint8 and depth.dtype == np.float32
        assert color.flags.c_contiguous and depth.flags.c_contiguous
        clr = (C.c_uint8 * 4)(*clear_rgba)
        _ck(self.L.hana_draw_model_host(self.h, W, Hh, _ptr(color), _ptr(depth), model.h, shader, C.byref(uniforms),
                                        _h(diffuse), _h(normal), int(assume_cleared), clr, float(clear_depth)))

    def stats(self):
        s = HanaStats()
        _ck(self.L.hana_last_stats(self.h, C.byref(s)))
        return s.as_dict()

    # --- stage-level
    def stage_vertex(self, model, shader, uniforms):
        out = np.zeros((model.ncorners, 13), np.float32)
        _ck(self.L.hana_stage_vertex(self.h, model.h, shader, C.byref(uniforms), _ptr(out)))
        return out

    def stage_setup(self, model, shader, uniforms, W, Hh, capacity=None):
        cap = capacity or (model.ncorners // 3) * 7 + 16
        order = np.zeros(cap, np.uint32)
        v2f = np.zeros((cap, 3, 13), np.float32)
        n = C.c_int()
        _ck(self.L.hana_stage_setup(self.h, model.h, shader, C.byref(uniforms), W, Hh, cap, _ptr(order), _ptr(v2f),
                                    C.byref(n)))
        return order[:n.value].copy(), v2f[:n.value].copy()


class Model:
    """The a2v stream graphics.cpp:380-386 gathers: ncorners x {obj_pos, obj_normal, uv}."""

    def __init__(self, ctx, a2v):
        a2v = np.ascontiguousarray(a2v, np.float32).reshape(-1, 8)
        self.ctx = ctx
        self.ncorners = a2v.shape[0]
        h = C.c_void_p()
        _ck(ctx.L.hana_model_upload(ctx.h, _ptr(a2v), self.ncorners, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.ctx.L.hana_model_destroy(self.h)
            self.h = None


class Texture:
    """TGAImage storage (tgaimage.cpp:248-253): uint8 [h, w, bytespp] in B,G,R[,A] order."""

    def __init__(self, ctx, data):
        data = np.ascontiguousarray(data, np.uint8)
        assert data.ndim == 3
        self.ctx = ctx
        h = C.c_void_p()
        _ck(ctx.L.hana_texture_upload(ctx.h, _ptr(data), data.shape[1], data.shape[0], data.shape[2], C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.ctx.L.hana_texture_destroy(self.h)
            self.h = None


class RenderBuffer:
    """Device-resident RenderBuffer (renderbuffer.h:5-22): RGBA8 colour + f32 depth, y up."""

    def __init__(self, ctx, w, h):
        self.ctx = ctx
        self.width, self.height = w, h
        hh = C.c_void_p()
        _ck(ctx.L.hana_rb_create(ctx.h, w, h, C.byref(hh)))
        self.h = hh

    def close(self):
        if self.h:
            self.ctx.L.hana_rb_destroy(self.h)
            self.h = None

    def clear_color(self, r=0, g=0, b=0, a=1):
        _ck(self.ctx.L.hana_rb_clear_color(self.h, r, g, b, a))

    def clear_depth(self, d=FLT_MAX):
        _ck(self.ctx.L.hana_rb_clear_depth(self.h, float(d)))

    def upload(self, color=None, depth=None):
        if color is not None:
            color = np.ascontiguousarray(color, np.uint8)
        if depth is not None:
            depth = np.ascontiguousarray(depth, np.float32)
        _ck(self.ctx.L.hana_rb_upload(self.h, _ptr(color), _ptr(depth)))

    def download(self):
        color = np.empty((self.height, self.width, 4), np.uint8)
        depth = np.empty((self.height, self.width), np.float32)
        _ck(self.ctx.L.hana_rb_download(self.h, _ptr(color), _ptr(depth)))
        return color, depth


class Sweep:
    """Batched frames: per frame clear + DrawModel::draw, the frame index a grid dimension."""

    def __init__(self, ctx, w, h, max_frames):
        self.ctx = ctx
        self.width, self.height, self.max_frames = w, h, max_frames
        hh = C.c_void_p()
        _ck(ctx.L.hana_sweep_create(ctx.h, w, h, max_frames, C.byref(hh)))
        self.h = hh

    def close(self):
        if self.h:
            self.ctx.L.hana_sweep_destroy(self.h)
            self.h = None

    @staticmethod
    def pack_uniforms(uniforms):
        arr = (HanaUniforms * len(uniforms))()
        for i, u in enumerate(uniforms):
            C.memmove(C.byref(arr, i * C.sizeof(HanaUniforms)), C.byref(u), C.sizeof(HanaUniforms))
        return arr

    def render(self, model, shader, uniforms, diffuse=None, normal=None, clear_rgba=(0, 0, 0, 1), clear_depth=FLT_MAX,
               n_frames=None):
        """uniforms: list of HanaUniforms or a packed ctypes array (host memory; copied H2D inside)."""
        arr = uniforms if isinstance(uniforms, C.Array) else self.pack_uniforms(uniforms)
        n = n_frames if n_frames is not None else len(arr)
        clr = (C.c_uint8 * 4)(*clear_rgba)
        _ck(self.ctx.L.hana_sweep_render(self.h, model.h, shader, arr, n, _h(diffuse), _h(normal), clr, float(clear_depth)))
        return n

    def render_resident(self, model, shader, n_frames, diffuse=None, normal=None, clear_rgba=(0, 0, 0, 1),
                        clear_depth=FLT_MAX, enable_shadow=True):
        """Uniforms already in the sweep's device buffer (from an earlier render())."""
        dev = C.c_void_p()
        _ck(self.ctx.L.hana_sweep_uniforms_dev(self.h, C.byref(dev)))
        clr = (C.c_uint8 * 4)(*clear_rgba)
        _ck(self.ctx.L.hana_sweep_render_dev(self.h, model.h, shader, dev, int(bool(enable_shadow)), n_frames, _h(diffuse),
                                             _h(normal), clr, float(clear_depth)))

    def set_shadow_reuse(self, enable=True):
        """Optional (off by default): a batch whose frames all carry the same light_vp and model matrices renders ONE
        shadow map for all of them (scene.h:73-88 with a static light); frames are byte-identical either way."""
        _ck(self.ctx.L.hana_sweep_set_shadow_reuse(self.h, int(bool(enable))))

    # -- one frame split by screen tiles over several GPUs --
    def set_bands(self, shadow=(0, 0), main=(0, 0)):
        """Tile rows (first, count) of the shadow / main pass this sweep renders; count 0 = all rows."""
        _ck(self.ctx.L.hana_sweep_set_bands(self.h, int(shadow[0]), int(shadow[1]), int(main[0]), int(main[1])))

    def render_pass(self, which, model, shader, uniforms, diffuse=None, normal=None, clear_rgba=(0, 0, 0, 1),
                    clear_depth=FLT_MAX):
        """ONE pass of DrawModel::draw (PASS_SHADOW: scene.h:86, PASS_MAIN: scene.h:91) over this sweep's band."""
        arr = uniforms if isinstance(uniforms, C.Array) else self.pack_uniforms(uniforms)
        clr = (C.c_uint8 * 4)(*clear_rgba)
        _ck(self.ctx.L.hana_sweep_render_pass(self.h, int(which), model.h, shader, arr, len(arr), _h(diffuse), _h(normal), clr,
                                              float(clear_depth)))

    def render_pass_async(self, which, model, shader, uniforms, diffuse=None, normal=None, clear_rgba=(0, 0, 0, 1),
                          clear_depth=FLT_MAX):
        """render_pass queued without any host synchronisation; passes_ok() afterwards."""
        arr = uniforms if isinstance(uniforms, C.Array) else self.pack_uniforms(uniforms)
        clr = (C.c_uint8 * 4)(*clear_rgba)
        _ck(self.ctx.L.hana_sweep_render_pass_async(self.h, int(which), model.h, shader, arr, len(arr), _h(diffuse), _h(normal),
                                                    clr, float(clear_depth)))

    def passes_ok(self):
        """Waits for the queued passes; False if scratch ran out (it has been grown: queue the frame again)."""
        ok = C.c_int()
        _ck(self.ctx.L.hana_sweep_passes_ok(self.h, C.byref(ok)))
        return bool(ok.value)

    def device_planes(self):
        """(colour ptr, depth ptr, frame stride in pixels) of the frame ring; synchronises."""
        c, d, st = C.c_void_p(), C.c_void_p(), C.c_size_t()
        _ck(self.ctx.L.hana_sweep_device_ptrs(self.h, C.byref(c), C.byref(d), C.byref(st)))
        return c.value, d.value, st.value

    def shadow_plane(self):
        """(ptr, pitch in bytes, frame stride in bytes) of the 1-byte shadow maps; synchronises."""
        p, pitch, st = C.c_void_p(), C.c_int(), C.c_size_t()
        _ck(self.ctx.L.hana_sweep_shadow_ptrs(self.h, C.byref(p), C.byref(pitch), C.byref(st)))
        return p.value, pitch.value, st.value

    def download(self, frame):
        color = np.empty((self.height, self.width, 4), np.uint8)
        depth = np.empty((self.height, self.width), np.float32)
        _ck(self.ctx.L.hana_sweep_download(self.h, frame, _ptr(color), _ptr(depth)))
        return color, depth

    def download_async(self, first, count, color_ptr, depth_ptr):
        _ck(self.ctx.L.hana_sweep_download_async(self.h, first, count, C.c_void_p(color_ptr),
                                                 C.c_void_p(depth_ptr) if depth_ptr else None))

    def present(self, first, count, fmt=0):
        """window_draw_buffer's surface (win32.cpp:348-370) of frames [first, first+count): top-down B,G,R,255 (fmt 0)
        or B,G,R (fmt 1), converted on the device."""
        out = np.empty((count, self.height, self.width, 4 if fmt == 0 else 3), np.uint8)
        _ck(self.ctx.L.hana_sweep_present(self.h, first, count, fmt, _ptr(out), None))
        self.ctx.sync()
        return out

    def tga_files(self, first, count):
        """Frames [first, first+count) as RLE TGA files (bytes objects), encoded on the device: what
        TGAImage::write_tga_file(rle=True) would write (tgaimage.cpp:145-246)."""
        _ck(self.ctx.L.hana_sweep_encode_tga(self.h, first, count))
        cap = count * (self.width * self.height * 3 + self.width * self.height // 64 + 2048)
        buf = np.empty(cap, np.uint8)
        offs = (C.c_uint64 * (count + 1))()
        sizes = (C.c_uint64 * count)()
        _ck(self.ctx.L.hana_sweep_fetch_tga(self.h, _ptr(buf), cap, offs, sizes))
        self.ctx.sync()
        return [bytes(buf[int(offs[f]):int(offs[f]) + int(sizes[f])]) for f in range(count)]

    def checksums(self, n_frames):
        out = np.zeros(n_frames, np.uint64)
        _ck(self.ctx.L.hana_sweep_checksums(self.h, n_frames, _ptr(out)))
        return out

    def overflow_count(self):
        """Batches that ran out of scratch and were overwritten before they could be rendered again (0 = none)."""
        n = C.c_uint64()
        _ck(self.ctx.L.hana_sweep_overflow_count(self.h, C.byref(n)))
        return int(n.value)

    def stats(self, frame):
        s = HanaStats()
        _ck(self.ctx.L.hana_sweep_stats(self.h, frame, C.byref(s)))
        return s.as_dict()


def frame_checksum(color, depth):
    """numpy twin of checksum_kernel (csrc/hana_kernels.cuh): order-free 64-bit sum of mixed pixels."""
    M1, M2 = np.uint64(0xFF51AFD7ED558CCD), np.uint64(0xC4CEB9FE1A85EC53)

    def mix(x):
        x = x ^ (x >> np.uint64(33))
        x = x * M1
        x = x ^ (x >> np.uint64(33))
        x = x * M2
        x = x ^ (x >> np.uint64(33))
        return x

    with np.errstate(over="ignore"):
        c = np.ascontiguousarray(color).reshape(-1, 4).astype(np.uint64)
        rgb = c[:, 0] | (c[:, 1] << np.uint64(8)) | (c[:, 2] << np.uint64(16))
        d = np.ascontiguousarray(depth, np.float32).reshape(-1).view(np.uint32).astype(np.uint64)
        v = (rgb << np.uint64(32)) | d
        idx = np.arange(v.size, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        return np.uint64(np.sum(mix(v ^ mix(idx)), dtype=np.uint64))


PRESENT_BGRA8, PRESENT_BGR8 = 0, 1
PASS_SHADOW, PASS_MAIN = 1, 2


def obj_load(path, normal_pass=1):
    """OBJ -> (ncorners, 8) float32 a2v stream; no GPU needed."""
    L = load()
    p, n = C.c_void_p(), C.c_int()
    _ck(L.hana_obj_load(os.fsencode(path), int(normal_pass), C.byref(p), C.byref(n)))
    try:
        a = np.ctypeslib.as_array((C.c_float * (n.value * 8)).from_address(p.value)).reshape(n.value, 8).copy() if n.value else \
            np.zeros((0, 8), np.float32)
    finally:
        L.hana_free(p)
    return a


def tga_load(path, model_flip=True):
    """TGA -> (h, w, bytespp) uint8 as Model holds it (model_flip) or as TGAImage::read_tga_file leaves it."""
    L = load()
    p, w, h, b = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
    _ck(L.hana_tga_load(os.fsencode(path), int(bool(model_flip)), C.byref(p), C.byref(w), C.byref(h), C.byref(b)))
    try:
        n = w.value * h.value * b.value
        a = np.ctypeslib.as_array((C.c_uint8 * n).from_address(p.value)).reshape(h.value, w.value, b.value).copy()
    finally:
        L.hana_free(p)
    return a


def tga_write(path, data, rle=False):
    """(h, w, bytespp) uint8, rows in file order -> TGA file (tgaimage.cpp:145-246 byte for byte)."""
    a = np.ascontiguousarray(data, np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    _ck(load().hana_tga_write(os.fsencode(path), _ptr(a), a.shape[1], a.shape[0], a.shape[2], int(bool(rle))))


class PinnedBuffer:
    """cudaMallocHost memory viewed as numpy (the e2e path's host side)."""

    def __init__(self, nbytes):
        self.L = load()
        p = C.c_void_p()
        _ck(self.L.hana_host_alloc(nbytes, C.byref(p)))
        self.ptr = p.value
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def close(self):
        if self.ptr:
            self.array = None
            self.L.hana_host_free(C.c_void_p(self.ptr))
            self.ptr = None
