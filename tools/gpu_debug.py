"""Scratch diagnostics for the GPU box: where does the CUDA path differ from the oracle?"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package, cleared
hana = load_package()
from oracle import horacle as H
port = H.Port()
W, Hh = 320, 240
scene = hana.synthetic_scene("blob")
u = hana.default_uniforms(W, Hh, True)
hu = H.HanaUniforms.from_bytes(u.to_bytes())
ctx = hana.Context(0)
print("tma", ctx.uses_tma, "sms", ctx.sm_count)
model, dtex, ntex = scene.upload(ctx)
scol, sdep = cleared(W, Hh)
pid_o, cnt = port.draw(H.SHADOW, hu, scene.a2v, W, Hh, scol, sdep, want_primid=True, want_counters=True)
print("oracle counters", cnt)
order, v2f = ctx.stage_setup(model, hana.SHADOW, u, W, Hh)
print("gpu setup tris", len(order), order[:8])
for tma in (False, True):
    ctx.set_tma(tma)
    rb = ctx.renderbuffer(W, Hh)
    rb.clear_color(0, 0, 0, 1); rb.clear_depth(3.4028234663852886e38)
    pid = ctx.draw(rb, model, hana.SHADOW, u, want_primid=True)
    col, dep = rb.download()
    print("tma", tma, "stats", ctx.stats())
    dm = dep.view(np.uint32) != sdep.view(np.uint32)
    cm = (col != scol).any(-1)
    pm = pid != pid_o
    print("  depth mismatches", dm.sum(), "colour", cm.sum(), "primid", pm.sum(), "covered gpu", (pid != 0xFFFFFFFF).sum(), "oracle", (pid_o != 0xFFFFFFFF).sum())
    if dm.any():
        ys, xs = np.nonzero(dm)
        print("  bbox of depth diffs x", xs.min(), xs.max(), "y", ys.min(), ys.max())
        for i in range(min(6, len(xs))):
            y, x = ys[i], xs[i]
            print("   (%d,%d) gpu z=%r pid=%d col=%s | oracle z=%r pid=%d col=%s" % (x, y, dep[y, x], pid[y, x], col[y, x], sdep[y, x], pid_o[y, x], scol[y, x]))
        # tile pattern
        tm = np.zeros(((Hh + 15) // 16, (W + 15) // 16), int)
        for y, x in zip(ys, xs): tm[y // 16, x // 16] += 1
        print(tm)
    rb.close()
# ---- deeper: which primitives are missing?
ctx.set_tma(False)
rb = ctx.renderbuffer(W, Hh)
rb.clear_color(0, 0, 0, 1); rb.clear_depth(3.4028234663852886e38)
pid = ctx.draw(rb, model, hana.SHADOW, u, want_primid=True)
want = set(np.unique(pid_o).tolist()) - {0xFFFFFFFF}
got = set(np.unique(pid).tolist()) - {0xFFFFFFFF}
oset = set(order.tolist())
print("prims oracle", len(want), "gpu", len(got), "setup list", len(oset), "missing from setup list", len(want - oset), "in list but never drawn", len((want & oset) - got))
miss = sorted((want & oset) - got)[:5]
print("examples never drawn", miss)
for k in miss[:3]:
    i = int(np.nonzero(order == k)[0][0])
    print(" key", k, "clip", v2f[i][:, :4])
    ys, xs = np.nonzero(pid_o == k)
    print("   oracle pixels x", xs.min(), xs.max(), "y", ys.min(), ys.max())
drawn = np.array(sorted(got)); notdrawn = np.array(sorted((want & oset) - got))
print("drawn keys min/max", drawn.min(), drawn.max(), "not drawn min/max", notdrawn.min(), notdrawn.max())
print("order index of drawn: min", np.searchsorted(order, drawn.min()), "of notdrawn max", np.searchsorted(order, notdrawn.max()))
for nf in (64, 256, 600, 1200):
    sub = scene.a2v[:nf * 3]
    m2 = ctx.model(sub)
    rb2 = ctx.renderbuffer(W, Hh); rb2.clear_color(0, 0, 0, 1); rb2.clear_depth(3.4028234663852886e38)
    p2 = ctx.draw(rb2, m2, hana.SHADOW, u, want_primid=True)
    c2, d2 = cleared(W, Hh)
    po, _ = port.draw(H.SHADOW, hu, sub, W, Hh, c2, d2, want_primid=True)
    print("faces", nf, "stats", ctx.stats(), "mismatch", (p2 != po).sum(), "oracle covered", (po != 0xFFFFFFFF).sum())
    rb2.close(); m2.close()
