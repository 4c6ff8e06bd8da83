"""CPU proxy for fast-tree tuning (not collected by pytest): search work (8-wide steps, leaf steps, triangle tests) of the
speculative traversal on a C3-like ray set - primary rays of the reference camera (every 2nd pixel of 1920x1080), plus shadow and
reflection rays leaving the primary hit points. usage: python tests/tree_work.py   (env: CGRT_FAST_TREE, CGRT_SAH_LEAF, ...)"""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import __graft_entry__ as ge
import test_spec_certificate_cpu as T
from oracle import bindings as ob

capi = ge.load_package().capi
so = os.path.join(T.HERE, "libspec_harness.so")
lib = C.CDLL(so)
d = capi.dragon_standin()
flat = ob.FlatScene(d.vcount, d.tcount, d.vertices, d.triangles, d.materials, d.spheres)
W, H = 960, 540
prim = ob.OracleLib().generate_rays(ob.default_camera(W, H), W, H)
ex, _, _, _ = T.run(lib, flat, prim, sah=True)
hit = ex[:, 0] >= 0
t = ex[hit, 1].view(np.float32)
P = prim["o"][hit] + prim["d"][hit] * t[:, None]
# geometric normals of the hit triangles
V, Tr = flat.vertices[:, :3], flat.triangles.astype(np.int64)
a, b, c = V[Tr[ex[hit, 0], 0]], V[Tr[ex[hit, 0], 1]], V[Tr[ex[hit, 0], 2]]
n = np.cross(b - a, c - a); n /= np.linalg.norm(n, axis=1, keepdims=True)
D = prim["d"][hit]
n = np.where((np.sum(n * -D, axis=1) > 0)[:, None], n, -n)
R = D - 2 * np.sum(D * n, axis=1, keepdims=True) * n
R /= np.linalg.norm(R, axis=1, keepdims=True)
L = np.array([-1, 1, -1], np.float32) - P
Ld = L / np.linalg.norm(L, axis=1, keepdims=True)
sets = {"primary (entering)": prim[np.linalg.norm(prim["d"], axis=1) > 0],
        "shadow": T.mk_rays((P + 0.001 * Ld).astype(np.float32), Ld.astype(np.float32), np.full(len(P), T.FLT_MAX)),
        "bounce": T.mk_rays((P + 0.001 * R).astype(np.float32), R.astype(np.float32), np.full(len(P), np.float32(1.0)))}
for flavour in (os.environ.get("FLAVOURS", "ref,sah").split(",")):
    tot = np.zeros(3)
    for name, rays in sets.items():
        r = np.ascontiguousarray(rays).view(np.float32).reshape(len(rays), 8)
        out = np.zeros(4, np.int64)
        dd = flat.desc()
        lib.spec_work(C.byref(dd), C.c_int(12), C.c_int(1 if flavour == "sah" else 0), r.ctypes.data_as(C.c_void_p), C.c_int64(len(rays)),
                      out.ctypes.data_as(C.c_void_p))
        tot += out[:3]
        print(f"{flavour:4s} {name:20s} rays {len(rays):7d}: wide {out[0] / len(rays):6.2f} leaf {out[1] / len(rays):5.2f} tris {out[2] / len(rays):6.2f} per ray, max steps {out[3]}")
    print(f"{flavour:4s} TOTAL wide {tot[0]:.0f} leaf {tot[1]:.0f} tris {tot[2]:.0f}  cost proxy (300*wide + 100*tris) {300 * tot[0] + 100 * tot[2]:.3g}")
