import sys, os, time, ctypes as C
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
import numpy as np, torch
hana = ge.load_package()
from hana_softwarerenderer_b200.api import PinnedBuffer, HanaUniforms
ctx = hana.Context(0)
sc = hana.load_bundled("african_head", None, 3)
model, dtex, ntex = sc.upload(ctx)
W,H=1920,1080
F=512
rings=(ctx.sweep(W,H,F), ctx.sweep(W,H,F))
arr = hana.orbit_sweep_uniforms(W,H,0,F,frames_per_turn=1024)
usz=C.sizeof(HanaUniforms)
pu=PinnedBuffer(usz*F); C.memmove(pu.ptr, arr, usz*F)
clr=(C.c_uint8*4)(0,0,0,1)
def render(r): assert ctx.L.hana_sweep_render(rings[r].h, model.h, hana.BLINN, C.c_void_p(pu.ptr), F, dtex.h, ntex.h, clr, float(hana.FLT_MAX))==0
def encode(r): assert ctx.L.hana_sweep_encode_tga(rings[r].h,0,F)==0
dev=torch.empty(500_000_000, dtype=torch.uint8, device="cuda"); host=torch.empty(500_000_000, dtype=torch.uint8).pin_memory()
side=torch.cuda.Stream()
N=8
for mode in ("kernels only","copy only","kernels + independent torch copy","render only + copy"):
    ctx.sync(); torch.cuda.synchronize(); t0=time.perf_counter()
    for s in range(N):
        if mode!="copy only":
            render(s&1)
            if mode!="render only + copy": encode(s&1)
        if mode!="kernels only":
            with torch.cuda.stream(side): host.copy_(dev, non_blocking=True)
    ctx.sync(); torch.cuda.synchronize()
    print("%-36s %.2f ms per iteration"%(mode,(time.perf_counter()-t0)*1e3/N))
