"""hana-softwarerenderer_b200 — B200-native rasterisation path behind Hana-SoftwareRenderer's draw API.

The product is ``libhana_b200.so`` (CUDA sm_100a kernels + the C ABI of ``include/hana_b200.h``).
This package is the thin Python view of that C ABI (ctypes; plain pointers and sizes), mirroring
the reference's names for the path: ``DrawData``-style draws (graphics.h:9-15), ``RenderBuffer``
(renderbuffer.h:5-22), ``DrawModel.draw`` (scene.h:53-99) and the batched frame sweep.

There is no CPU fallback: importing works anywhere, but every compute entry point raises
``HanaError`` without a CUDA device, and ``load()`` raises if the library has not been built.
Nothing here imports, links or executes anything under ``oracle/``.
"""
from .api import (  # noqa: F401
    BLINN,
    GROUND,
    NORMALMAP,
    SHADOW,
    TEXTURE,
    TEXTURE_LIGHT,
    TOON,
    FLT_MAX,
    Context,
    HanaError,
    HanaStats,
    HanaUniforms,
    Model,
    RenderBuffer,
    Sweep,
    Texture,
    build,
    device_count,
    lib_path,
    load,
    obj_load,
    tga_load,
    tga_write,
    PRESENT_BGRA8,
    PRESENT_BGR8,
    PASS_SHADOW,
    PASS_MAIN,
)
from . import scene, sharding  # noqa: F401
from .scene import (  # noqa: F401
    OrbitCamera,
    Scene,
    default_uniforms,
    load_bundled,
    load_hscene,
    orbit_sweep_uniforms,
    scene_desc,
    synthetic_scene,
)
