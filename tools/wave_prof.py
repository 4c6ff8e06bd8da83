"""One configuration rendered a few times through the default pipeline (ncu target for k_wave):
    python tools/wave_prof.py [dragon|dodge|cornell|monkey] W H L [frames]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from conftest import load_golden  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dragon"
W, H, L = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (1920, 1080, 5)
frames = int(sys.argv[5]) if len(sys.argv) > 5 else 4
rank, world = (int(v) for v in sys.argv[6:8]) if len(sys.argv) > 7 else (0, 1)
capi = ge.load_package().capi
if name == "dragon":
    flat, lights = ob.dragon_standin_fixture()
else:
    g = load_golden(name)
    flat, lights = g.flat, g.lights
s = capi.Scene(flat, lights=lights, device=0)
cam = capi.make_camera(W, H)
ms = []
out = np.zeros((H, W, 3), np.float32)
for _ in range(frames):
    _, st = s.render(cam, W, H, trace_limit=L, rank=rank, world=world, out=out)
    ms.append(round(st["device_ms"], 4))
print(name, W, H, L, "rank", rank, "of", world, "pipeline", st["pipeline"], "device ms", ms, "rays", st["primary"] + st["shadow"] + st["bounce"])
import ctypes as C  # noqa: E402
buf = (C.c_float * 16)()
if capi.load_library().cgrt_debug_wave_tuner(s.h, buf, 16):
    print("  finisher share: 1/%d of the SMs after %d frames; measured ms per setting %s" % (
        int(buf[0]), int(buf[1]), {f: round(buf[f], 4) for f in range(2, 15) if buf[f] > 0}))
