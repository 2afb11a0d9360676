"""Scene inputs for the path: packed scenes (a2v stream + decoded textures), the synthetic
workloads of BASELINE.json configs[3..4], and the Python view of the C++ host mirror
(csrc/../host/hana_host.cpp: Camera, DrawModel::draw's uniform block, the orbit sweep)."""
import ctypes as C
import os

import numpy as np

from . import api
from .api import HanaUniforms


class HanaCamera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("target", C.c_float * 3), ("aspect", C.c_float)]


class HanaSceneDesc(C.Structure):
    _fields_ = [
        ("light_pos", C.c_float * 3),
        ("model_pos", C.c_float * 3),
        ("model_rot_deg", C.c_float * 3),
        ("model_scale", C.c_float * 3),
        ("light_color", C.c_float * 4),
        ("ambient", C.c_float * 4),
        ("mat_color", C.c_float * 4),
        ("mat_specular", C.c_float * 4),
        ("gloss", C.c_float),
        ("bump_scale", C.c_float),
    ]


def _lib():
    L = api.load()
    if not getattr(L, "_host_sigs", False):
        L.hana_camera_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        L.hana_camera_update.argtypes = [C.c_void_p] + [C.c_float] * 5
        L.hana_scene_defaults.argtypes = [C.c_void_p]
        L.hana_scene_uniforms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.hana_orbit_sweep_uniforms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L._host_sigs = True
    return L


def scene_desc(**kw):
    d = HanaSceneDesc()
    _lib().hana_scene_defaults(C.byref(d))
    for k, v in kw.items():
        cur = getattr(d, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(d, k, v)
    return d


class OrbitCamera:
    """Camera (camera.h:13-33) with Camera::update_transform (camera.cpp:63-70)."""

    def __init__(self, aspect, position=(0.0, 0.0, 2.0), target=(0.0, 0.0, 0.0)):
        self.c = HanaCamera()
        p = (C.c_float * 3)(*position)
        t = (C.c_float * 3)(*target)
        _lib().hana_camera_init(C.byref(self.c), p, t, aspect)

    def update(self, orbit=(0.0, 0.0), pan=(0.0, 0.0), dolly=0.0):
        _lib().hana_camera_update(C.byref(self.c), orbit[0], orbit[1], pan[0], pan[1], dolly)

    @property
    def position(self):
        return np.array(list(self.c.position), np.float32)


def default_uniforms(width, height, enable_shadow=True, camera=None, desc=None):
    """The ShaderData DrawModel::draw builds (scene.h:55-71) for `camera` (default: CAMERA_POSITION -> origin)."""
    cam = camera or OrbitCamera(np.float32(width) / np.float32(height))
    d = desc or scene_desc()
    u = HanaUniforms()
    r = _lib().hana_scene_uniforms(C.byref(cam.c), C.byref(d), width, height, int(enable_shadow), C.byref(u))
    if r != 0:
        raise api.HanaError(r, "hana_scene_uniforms")
    return u


def orbit_sweep_uniforms(width, height, first, count, frames_per_turn=1024, enable_shadow=True, desc=None):
    """Packed HanaUniforms array for frames [first, first+count) of the orbit sweep (BASELINE configs[2])."""
    d = desc or scene_desc()
    arr = (HanaUniforms * max(count, 1))()
    r = _lib().hana_orbit_sweep_uniforms(C.byref(d), width, height, int(enable_shadow), first, count, frames_per_turn, arr)
    if r != 0:
        raise api.HanaError(r, "hana_orbit_sweep_uniforms")
    return arr


class Scene:
    """Host-side scene inputs: a2v [ncorners, 8] f32, diffuse / normal textures in TGAImage layout (or None)."""

    def __init__(self, name, a2v, diffuse=None, normal=None):
        self.name = name
        self.a2v = np.ascontiguousarray(a2v, np.float32).reshape(-1, 8)
        self.diffuse = diffuse
        self.normal = normal

    @property
    def nfaces(self):
        return self.a2v.shape[0] // 3

    def upload(self, ctx):
        return ctx.model(self.a2v), ctx.texture(self.diffuse), ctx.texture(self.normal)


ASSET_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")


def load_bundled(name, root=None, normal_pass=1):
    """A scene of the reference's assets/ layout (<root>/<name>/<name>.obj, <name>_diffuse.tga, <name>_nm_tangent.tga:
    model.cpp:6-48), read by THIS library's own OBJ / TGA readers (hana_obj_load, hana_tga_load). Raises if the files
    are not there: nothing is substituted."""
    from .api import obj_load, tga_load
    d = os.path.join(root or ASSET_ROOT, name)
    obj = os.path.join(d, name + ".obj")
    if not os.path.exists(obj):
        raise FileNotFoundError("bundled scene %s not found under %s (__graft_entry__.build() copies the reference's "
                                "assets/ there)" % (name, root or ASSET_ROOT))
    tex = []
    for suffix in ("_diffuse.tga", "_nm_tangent.tga"):
        p = os.path.join(d, name + suffix)
        tex.append(tga_load(p, model_flip=True) if os.path.exists(p) else None)
    return Scene(name, obj_load(obj, normal_pass), tex[0], tex[1])


def load_hscene(path):
    """A packed scene: npz with a2v, diffuse, normal (written by oracle/pack_assets.py from the bundled assets)."""
    z = np.load(path)
    return Scene(os.path.splitext(os.path.basename(path))[0], z["a2v"], z["diffuse"] if "diffuse" in z else None,
                 z["normal"] if "normal" in z else None)


def _value_noise(n, seed, lattice=64):
    rng = np.random.RandomState(seed)
    g = rng.rand(lattice + 1, lattice + 1).astype(np.float32)
    t = np.linspace(0, lattice, n, endpoint=False, dtype=np.float32)
    i = np.floor(t).astype(np.int32)
    f = t - i
    f = f * f * (3 - 2 * f)
    a = g[i][:, i]
    b = g[i][:, i + 1]
    c = g[i + 1][:, i]
    d = g[i + 1][:, i + 1]
    fx = f[None, :]
    fy = f[:, None]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def noise_textures(seed, size=1024, flat_normal=False):
    rng = np.random.RandomState(seed)
    diffuse = rng.randint(0, 256, (size, size, 3)).astype(np.uint8)
    if flat_normal:
        normal = np.empty((size, size, 3), np.uint8)
        normal[..., 0] = 255  # B = z
        normal[..., 1] = 128
        normal[..., 2] = 128
    else:
        v = rng.randn(size, size, 3).astype(np.float32) * 0.25 + np.array([0, 0, 1], np.float32)
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        rgb = np.clip((v * 0.5 + 0.5) * 255, 0, 255).astype(np.uint8)
        normal = np.ascontiguousarray(rgb[..., ::-1])  # B,G,R
    return diffuse, normal


def synthetic_grid(nx, ny, seed=1234, x_half=2.0, y_half=1.1, z_amp=0.05):
    """Height-field grid of (nx-1)*(ny-1)*2 triangles facing +z (SURVEY.md §8d, config C4 shape)."""
    n = max(nx, ny)
    h = _value_noise(n, seed)[:ny, :nx] * np.float32(z_amp)
    xs = np.linspace(-x_half, x_half, nx, dtype=np.float32)
    ys = np.linspace(-y_half, y_half, ny, dtype=np.float32)
    X, Y = np.meshgrid(xs, ys)
    P = np.stack([X, Y, h], -1)
    dzdx = np.gradient(h, axis=1) / np.float32(xs[1] - xs[0])
    dzdy = np.gradient(h, axis=0) / np.float32(ys[1] - ys[0])
    N = np.stack([-dzdx, -dzdy, np.ones_like(h)], -1)
    N /= np.linalg.norm(N, axis=-1, keepdims=True)
    U, V = np.meshgrid(np.linspace(0.001, 0.999, nx, dtype=np.float32), np.linspace(0.001, 0.999, ny, dtype=np.float32))
    vert = np.concatenate([P, N, U[..., None], V[..., None]], -1).astype(np.float32)  # [ny,nx,8]
    v00 = vert[:-1, :-1]
    v10 = vert[:-1, 1:]
    v01 = vert[1:, :-1]
    v11 = vert[1:, 1:]
    t1 = np.stack([v00, v10, v11], 2)  # CCW seen from +z
    t2 = np.stack([v00, v11, v01], 2)
    tris = np.stack([t1, t2], 2).reshape(-1, 3, 8)
    return np.ascontiguousarray(tris.reshape(-1, 8))


def synthetic_layers(layers=8, qx=32, qy=18, seed=99, aspect=16.0 / 9.0, fill=0.98):
    """`layers` screen-filling rectangles of qx*qy quads, submitted back to front (config C5 shape)."""
    rng = np.random.RandomState(seed)
    out = []
    tan_half = np.tan(np.radians(60.0) / 2)
    for k in range(layers):
        z = -0.35 + 0.1 * k
        dist = 2.0 - z
        hy = dist * tan_half * fill
        hx = hy * aspect
        jx, jy = rng.uniform(-0.005, 0.005, 2)
        xs = np.linspace(-hx, hx, qx + 1, dtype=np.float32) + np.float32(jx)
        ys = np.linspace(-hy, hy, qy + 1, dtype=np.float32) + np.float32(jy)
        X, Y = np.meshgrid(xs, ys)
        U, V = np.meshgrid(np.linspace(0.001, 0.999, qx + 1, dtype=np.float32),
                           np.linspace(0.001, 0.999, qy + 1, dtype=np.float32))
        Z = np.full_like(X, z)
        nrm = np.zeros(X.shape + (3,), np.float32)
        nrm[..., 2] = 1
        nrm[..., 0] = 0.05 * np.sin(X * 3 + k)
        nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
        vert = np.concatenate([X[..., None], Y[..., None], Z[..., None], nrm, U[..., None], V[..., None]], -1).astype(np.float32)
        v00, v10, v01, v11 = vert[:-1, :-1], vert[:-1, 1:], vert[1:, :-1], vert[1:, 1:]
        t1 = np.stack([v00, v10, v11], 2)
        t2 = np.stack([v00, v11, v01], 2)
        out.append(np.stack([t1, t2], 2).reshape(-1, 8))
    return np.ascontiguousarray(np.concatenate(out, 0))


def synthetic_scene(kind="blob", seed=7, tex=256):
    """Small procedural scenes for tests and for runs without the bundled assets."""
    if kind == "grid":
        a2v = synthetic_grid(64, 36, seed)
    elif kind == "layers":
        a2v = synthetic_layers(4, 8, 5, seed)
    else:  # a bumpy sphere: front/back faces, silhouettes, all shader inputs varying
        nu, nv = 48, 24
        th = np.linspace(0, 2 * np.pi, nu + 1, dtype=np.float32)
        ph = np.linspace(0.05, np.pi - 0.05, nv + 1, dtype=np.float32)
        T, Pp = np.meshgrid(th, ph)
        rng = np.random.RandomState(seed)
        r = 0.8 + 0.08 * np.sin(3 * T) * np.sin(4 * Pp) + 0.01 * rng.rand(*T.shape).astype(np.float32)
        r[:, -1] = r[:, 0]
        X, Y, Z = r * np.sin(Pp) * np.sin(T), r * np.cos(Pp), r * np.sin(Pp) * np.cos(T)
        P = np.stack([X, Y, Z], -1).astype(np.float32)
        N = P / np.linalg.norm(P, axis=-1, keepdims=True)
        N = N * (1.0 + 0.0005 * rng.randn(*T.shape)[..., None]).astype(np.float32)  # not exactly unit, as OBJ normals
        U = (T / (2 * np.pi)) * 0.998 + 0.001
        V = (Pp / np.pi) * 0.998 + 0.001
        vert = np.concatenate([P, N, U[..., None], V[..., None]], -1).astype(np.float32)
        v00, v10, v01, v11 = vert[:-1, :-1], vert[:-1, 1:], vert[1:, :-1], vert[1:, 1:]
        t1 = np.stack([v00, v11, v10], 2)
        t2 = np.stack([v00, v01, v11], 2)
        a2v = np.stack([t1, t2], 2).reshape(-1, 8)
    diffuse, normal = noise_textures(seed, tex)
    return Scene("synthetic_" + kind, a2v, diffuse, normal)
