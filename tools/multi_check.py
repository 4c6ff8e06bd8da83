"""Multi-GPU parity check (run under torchrun, one rank per GPU): the frame assembled on rank 0 from all ranks' tiles must be
bit-identical to the frame one GPU renders alone, for both exchange modes (NVLink peer stores + flags, NCCL gather).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    pkg = ge.load_package()
    capi = pkg.capi
    from importlib import import_module
    D = import_module("cg_raytracer_b200.distributed")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    d = capi.dragon_standin()
    scene = capi.Scene(d, lights=d.lights, device=local)
    ok = True
    rounds = int(os.environ.get("MULTI_CHECK_ROUNDS", "1"))  # > 1: soak (renderers opened and closed again, shapes alternating)
    for (W, H, L) in ((1920, 1080, 5), (333, 201, 2), (64, 64, 0), (1280, 720, 3)) * rounds:
        cam = capi.make_camera(W, H)
        ref = None
        if rank == 0:
            single = D.TiledRenderer(scene, W, H, L, 0, 1, local)
            ref = single.render_device(cam).clone()
            torch.cuda.synchronize(dev)
        for mode in ("p2p", "nccl"):
            R = D.TiledRenderer(scene, W, H, L, rank, world, local, mode=mode)
            for k in range(4):
                f = R.render_device(cam)
                if rank == 0:
                    same = bool(torch.equal(f, ref))
                    ok &= same
                    if k == 0 or not same:
                        print(f"[multi_check] {W}x{H} limit {L} mode {R.mode} (asked {mode}, fallback {R.fallback_reason}) frame {k}: "
                              f"{'bit-identical' if same else 'MISMATCH max|d|=%g' % float((f - ref).abs().max())}", flush=True)
                    f.zero_()  # the next frame must rewrite every pixel
            hs = R.stream_to_host(cam, 5)  # two frames in flight (p2p) / synchronous frames (nccl)
            if rank == 0:
                ok &= bool(np.array_equal(hs.reshape(-1), ref.cpu().numpy()))
            h = R.render_to_host(cam)
            if rank == 0:
                same = np.array_equal(h.reshape(-1), ref.cpu().numpy()) and ok
                ok &= bool(same)
                print(f"[multi_check] {W}x{H} mode {R.mode} host frame: {'bit-identical' if same else 'MISMATCH'}; timeouts {R.timeouts()}", flush=True)
                ok &= R.timeouts() == 0
            else:
                R.timeouts()
            dist.barrier(device_ids=[local])
            R.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        raise SystemExit(1)
    if rank == 0:
        print("[multi_check] OK")


if __name__ == "__main__":
    main()
