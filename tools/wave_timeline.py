"""Timeline of one k_wave frame (rays in flight, queue counters per 4 us): CGRT_WAVE_TRACE=1 python tools/wave_timeline.py [scene W H L]"""
import ctypes as C
import os
import sys

import numpy as np

os.environ.setdefault("CGRT_WAVE_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from conftest import load_golden  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dragon"
W, H, L = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (1920, 1080, 5)
rank, world = (int(v) for v in sys.argv[5:7]) if len(sys.argv) > 6 else (0, 1)
capi = ge.load_package().capi
flat, lights = ob.dragon_standin_fixture() if name == "dragon" else (load_golden(name).flat, load_golden(name).lights)
s = capi.Scene(flat, lights=lights, device=0)
cam = capi.make_camera(W, H)
out = np.zeros((H, W, 3), np.float32)
for _ in range(3):
    _, st = s.render(cam, W, H, trace_limit=L, rank=rank, world=world, out=out)
lib = capi.load_library()
lib.cgrt_debug_wave_timeline.restype = C.c_int
lib.cgrt_debug_wave_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32]
buf = np.zeros((1024, 8), np.int32)
n = lib.cgrt_debug_wave_timeline(s.h, buf.ctypes.data_as(C.POINTER(C.c_int32)), 1024)
print(f"{name} {W}x{H} L{L} rank {rank}/{world}: device ms {st['device_ms']:.4f}; samples {int(buf[:, 0].sum())}")
print("   t_us  in_flight  q1_backlog  q1_tail  fin_backlog  fin_tail  q2_backlog  q2_tail")
for k in range(n):
    v, p, h1, t1, fh, ft, h2, t2 = [int(x) for x in buf[k]]
    if v:
        t1c = t1 & 0x3fffffff
        print(f"{k * 4.096:7.1f} {p:10d} {t1c - h1:11d} {t1c:8d}{'*' if t1 & 0x40000000 else ' '} {ft - fh:11d} {ft:9d} {t2 - h2:11d} {t2:8d}")
