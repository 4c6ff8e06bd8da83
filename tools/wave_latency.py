"""Where a k_wave frame's time goes per ray (instrumented build: tools/build_lat.sh, then
    CGRT_LIB=$PWD/build_variants/lib_lat.so CGRT_WAVE_LAT=1 python tools/wave_latency.py [scene W H L rank world])
Per ray ticket the kernel records: emitted, search started, search ended, finish loaded, finish done (ns since the frame's
origin), steps (GROUP form), form, helpers. Printed: the stages' mean / median / p95 for the rays emitted in each tenth of the
frame - the last rows are the frame's critical path."""
import ctypes as C
import os
import sys

import numpy as np

os.environ.setdefault("CGRT_WAVE_LAT", "1")
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
lib.cgrt_debug_wave_latency.restype = C.c_int
lib.cgrt_debug_wave_latency.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_int32]
cap = 8 * 1024 * 1024
buf = np.zeros((cap, 16), np.uint32)
n = lib.cgrt_debug_wave_latency(s.h, buf.ctypes.data_as(C.POINTER(C.c_uint32)), cap)
r = buf[:n].astype(np.int64)
r = r[(r[:, 0] > 0) & (r[:, 5] > 0)]
print(f"{name} {W}x{H} L{L} rank {rank}/{world}: device ms {st['device_ms']:.4f}; rays with a complete record {len(r)}")
end = r[:, 5].max()
print(f"frame's last finish at {end / 1000:.1f} us")
hdr = "emitted in us      rays   wait->start      search   steps  ->finisher      finish   resumed%  helped%   (us: mean / median / p95)"
print(hdr)


def q(v):
    return f"{v.mean() / 1000:5.1f}/{np.median(v) / 1000:5.1f}/{np.percentile(v, 95) / 1000:5.1f}" if len(v) else "     -"


edges = np.linspace(0, end, 11)
for a, b in zip(edges[:-1], edges[1:]):
    m = r[(r[:, 0] >= a) & (r[:, 0] < b)]
    if len(m) == 0:
        continue
    stp = m[m[:, 2] > 0][:, 2]
    print(f"{a / 1000:6.0f}-{b / 1000:6.0f} {len(m):9d}  {q(m[:, 1] - m[:, 0])} {q(m[:, 3] - m[:, 1])} {stp.mean() if len(stp) else 0:6.1f} "
          f"{q(m[:, 4] - m[:, 3])} {q(m[:, 5] - m[:, 4])}  {100.0 * (m[:, 6] >= 4).mean():6.1f} {100.0 * (m[:, 7] > 0).mean():7.1f}")
# the rays whose finish came in the last 5 % of the frame: the end of the critical path
m = r[r[:, 5] > 0.95 * end]
print(f"rays finished in the last 5 % of the frame: {len(m)}")
print(f"  wait->start {q(m[:, 1] - m[:, 0])}  search {q(m[:, 3] - m[:, 1])}  ->finisher {q(m[:, 4] - m[:, 3])}  finish {q(m[:, 5] - m[:, 4])}"
      f"  steps {m[m[:, 2] > 0][:, 2].mean() if (m[:, 2] > 0).any() else 0:.1f}")
f = r[(r[:, 8] > 0) & (r[:, 9] > 0) & (r[:, 11] > 0)]
h = f[f[:, 10] > 0]
print(f"inside the finish batch (all rays): ray record {q(f[:, 8] - f[:, 4])}  certificate+epilogue {q(f[:, 9] - f[:, 8])}  "
      f"-> tickets reserved {q(f[:, 11] - f[:, 9])}  emission {q(f[:, 5] - f[:, 11])}")
print(f"   hits only ({len(h)}): certificate+epilogue {q(h[:, 9] - h[:, 8])}  normal/material/record {q(h[:, 10] - h[:, 9])}")
late = f[f[:, 5] > 0.7 * end]
print(f"   finished in the last 30 % ({len(late)}): ray record {q(late[:, 8] - late[:, 4])}  certificate+epilogue {q(late[:, 9] - late[:, 8])}  "
      f"-> tickets reserved {q(late[:, 11] - late[:, 9])}  emission {q(late[:, 5] - late[:, 11])}")
long = r[(r[:, 3] - r[:, 1]) > 20000]
print(f"searches longer than 20 us: {len(long)}; of them in GROUP form {int((long[:, 6] & 2).astype(bool).sum())}, "
      f"mean steps {long[long[:, 2] > 0][:, 2].mean() if (long[:, 2] > 0).any() else 0:.1f}, "
      f"ns per step {((long[:, 3] - long[:, 1])[long[:, 2] > 0] / long[long[:, 2] > 0][:, 2]).mean() if (long[:, 2] > 0).any() else 0:.0f}")

print("GROUP-form searches by number of steps: count, search us mean / median / p95")
gr = r[(r[:, 6] & 2).astype(bool)]
for lo_, hi_ in ((0, 0), (1, 1), (2, 2), (3, 4), (5, 8), (9, 16), (17, 32), (33, 64), (65, 10000)):
    m = gr[(gr[:, 2] >= lo_) & (gr[:, 2] <= hi_)]
    if len(m):
        late = m[m[:, 1] > 0.6 * end]
        print(f"  steps {lo_:3d}-{hi_:5d}: {len(m):8d}  {q(m[:, 3] - m[:, 1])}    started in the last 40 %: {len(late):6d} {q(late[:, 3] - late[:, 1])}")
