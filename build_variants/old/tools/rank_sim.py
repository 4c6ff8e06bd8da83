"""One rank's share of the C3 frame on one GPU (world = N tiles partition, rank 0): per-class device times of the round
pipeline in the regime an N-GPU run puts every GPU in. usage: python tools/rank_sim.py <world> [frames]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
capi = ge.load_package().capi
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 10
d = capi.dragon_standin()
s = capi.Scene(d, lights=d.lights)
W, H, L = 1920, 1080, 5
cam = capi.make_camera(W, H)
p = capi.render_params(W, H, L, 0, world)
buf = torch.empty(capi.tile_buffer_floats(p), dtype=torch.float32, device="cuda:0")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for _ in range(3):
    s.render_device(cam, p, buf.data_ptr(), st)
torch.cuda.synchronize()
tot = np.zeros(4); ms = 0.0
for _ in range(frames):
    flush.zero_()
    p.flags = capi.RENDER_PROFILE_ALL
    s.render_device(cam, p, buf.data_ptr(), st)
    r = s.collect_stats()
    tot += np.array(r["class_ms"]); ms += r["device_ms"]
print(f"world {world} rank 0: {ms / frames:.3f} ms/frame device; classes", dict(zip(capi.class_names(r), (tot / frames).round(4))),
      "rays", r["primary"], r["shadow"], r["bounce"], "launches", r["class_launches"])
