import sys, os, time, ctypes as C
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
import numpy as np
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
cap=F*(1<<21)
pin=(PinnedBuffer(cap),PinnedBuffer(cap))
offs=((C.c_uint64*(F+1))(),(C.c_uint64*(F+1))()); szs=((C.c_uint64*F)(),(C.c_uint64*F)())
def render(r): assert ctx.L.hana_sweep_render(rings[r].h, model.h, hana.BLINN, C.c_void_p(pu.ptr), F, dtex.h, ntex.h, clr, float(hana.FLT_MAX))==0
def encode(r): assert ctx.L.hana_sweep_encode_tga(rings[r].h,0,F)==0
def fetch(r): assert ctx.L.hana_sweep_fetch_tga(rings[r].h, C.c_void_p(pin[r].ptr), C.c_size_t(cap), offs[r], szs[r])==0
N=10
for mode in ("render","render+encode","render+encode+fetch(sync each)","pipelined","fetch-first"):
    for rep in range(2):
        ctx.sync(); ctx.timer_start(); t0=time.perf_counter(); host=[]
        pend=None
        for s in range(N):
            r=s&1
            h0=time.perf_counter()
            if mode=="fetch-first" and pend is not None: fetch(pend)
            render(r)
            if mode!="render": encode(r)
            if mode.startswith("render+encode+fetch"): fetch(r)
            if mode=="pipelined":
                if pend is not None: fetch(pend)
                pend=r
            if mode=="fetch-first": pend=r
            host.append((time.perf_counter()-h0)*1e3)
        if mode in ("pipelined","fetch-first") and pend is not None: fetch(pend)
        ms=ctx.timer_stop()
    print("%-34s %.2f ms per %d-frame batch = %.0f frames/s; host ms per iteration: %s"%(mode, ms/N, F, F*N/ms*1e3, " ".join("%.1f"%x for x in host[:6])))
