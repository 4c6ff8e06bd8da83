"""C4: 1 M-triangle soup, 16 777 216 incoherent rays, device-resident batch queries: closest hit, any-hit with
range = inf, any-hit with range ~ U(0,1). Prints one JSON line; CGRT_BATCH_ONE_KERNEL=1 selects the one-kernel form for A/B,
C4_EXACT=1 builds the scene without the fast tree (exact reference-order traversal only)."""
import ctypes as C
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import bindings as ob  # noqa: E402

capi = ge.load_package().capi
lib = capi.load_library()
n = int(os.environ.get("C4_RAYS", 16 * 1024 * 1024))
flat = ob.random_soup(1_000_000, seed=1234, scale=0.01, smooth_normals=False)
s = capi.Scene(flat, exact_only=bool(os.environ.get("C4_EXACT")))
rays = ob.random_rays(n, seed=5678)
dR, dH, dM, dO = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
for p, b in ((dR, n * 32), (dH, n * 32), (dM, n * 4), (dO, n)):
    capi.check(lib.cgrt_device_malloc(0, b, C.byref(p)))
capi.check(lib.cgrt_memcpy_h2d(0, dR, C.c_void_p(rays.ctypes.data), n * 32))


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        lib.cgrt_device_synchronize(0)
        t0 = time.perf_counter()
        fn()
        lib.cgrt_device_synchronize(0)
        best = min(best, time.perf_counter() - t0)
    return best


out = {"config": "C4 soup 1M tris / %d incoherent rays" % n, "one_kernel": bool(os.environ.get("CGRT_BATCH_ONE_KERNEL")),
       "exact_only": bool(os.environ.get("C4_EXACT"))}
tc = timed(lambda: capi.check(lib.cgrt_intersect_closest_device(s.h, dR, n, dH, None, None)))
hits = np.zeros(n, capi.HIT_DTYPE)
capi.check(lib.cgrt_memcpy_d2h(0, C.c_void_p(hits.ctypes.data), dH, n * 32))
out.update(closest_ms=round(tc * 1e3, 2), closest_Mrays_s=round(n / tc / 1e6, 1), hit_frac=round(float((hits["tri"] >= 0).mean()), 4),
           hits_sha=hashlib.sha1(hits.tobytes()).hexdigest()[:16])
for label, md in (("any_inf", np.full(n, np.inf, np.float32)), ("any_u01", np.random.default_rng(2).uniform(0, 1, n).astype(np.float32))):
    capi.check(lib.cgrt_memcpy_h2d(0, dM, C.c_void_p(md.ctypes.data), n * 4))
    ta = timed(lambda: capi.check(lib.cgrt_intersect_any_device(s.h, dR, dM, C.c_float(0.001), n, dO, None)))
    occ = np.zeros(n, np.uint8)
    capi.check(lib.cgrt_memcpy_d2h(0, C.c_void_p(occ.ctypes.data), dO, n))
    out.update({label + "_ms": round(ta * 1e3, 2), label + "_Mrays_s": round(n / ta / 1e6, 1), label + "_occluded": round(float(occ.mean()), 4),
                label + "_sha": hashlib.sha1(occ.tobytes()).hexdigest()[:16]})
print(json.dumps(out))
