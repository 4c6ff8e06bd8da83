"""All five BASELINE.json configurations on N GPUs of one box: frames through the tile partition with the
NVLink peer-store exchange (C5 also through the NCCL gather), C4 split in contiguous ray chunks (no exchange).
    python tools/configs_multi.py                                                        # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/configs_multi.py
Rank 0 prints one JSON line per configuration: device-timed Mrays/s (CUDA events per frame, max over ranks, exchange included),
frame == the frame one GPU renders alone, and at N = 1 the CPU reference (oracle/_ref, all host threads) beside it."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from conftest import load_golden  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from importlib import import_module
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    ge.build()
    capi = ge.load_package().capi
    D = import_module("cg_raytracer_b200.distributed")
    lib = capi.load_library()
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    def allsum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nthreads = max(1, len(os.sched_getaffinity(0)))
    steps = int(os.environ.get("CONFIG_STEPS", 20))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    only = [t for t in os.environ.get("CONFIG_ONLY", "").split(",") if t]  # e.g. CONFIG_ONLY=C1,C3

    def frame_config(name, flat, lights, W, H, L, modes):
        if only and name.split()[0] not in only:
            return
        scene = capi.Scene(flat, lights=lights, device=local)
        cam = capi.make_camera(W, H)
        for mode in modes if world > 1 else [None]:
            R = D.TiledRenderer(scene, W, H, L, rank, world, local, mode=mode)
            for _ in range(3):
                R.render_device(cam)
            barrier()
            st = scene.collect_stats()
            rays = allsum([st["primary"] + st["shadow"] + st["bounce"], st["primary"], st["shadow"], st["bounce"]])
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            barrier()
            for k in range(steps):
                flush.zero_()
                ev[k][0].record()
                R.render_device(cam)
                ev[k][1].record()
            barrier()
            ms = allmax(sum(a.elapsed_time(b) for a, b in ev)) / steps
            got = R.render_to_host(cam)
            got = got.copy() if rank == 0 else None
            same = None
            barrier()
            if rank == 0 and world > 1:
                single, _ = scene.render(cam, W, H, trace_limit=L)
                same = bool(np.array_equal(single.view(np.uint32), got.view(np.uint32)))
            barrier()
            line = dict(config=name, n_gpus=world, exchange=R.mode, W=W, H=H, trace_limit=L, rays=rays[0], primary=rays[1], shadow=rays[2],
                        bounce=rays[3], ms_per_frame=round(ms, 4), Mrays_s=round(rays[0] / ms / 1e3, 1), frame_equals_single_gpu=same,
                        pipeline=st["pipeline"], handoff_timeouts=R.timeouts())
            if rank == 0 and world == 1:
                b = ob.RefLib().scene(flat, lights).bvh(mode=1)
                ocam = ob.default_camera(W, H)
                b.render(ocam, min(W, 480), min(H, 270), trace_limit=L, nthreads=nthreads)  # thread pool warm-up
                t0 = time.perf_counter()
                ref, cnt = b.render(ocam, W, H, trace_limit=L, duplicate_shading=True, nthreads=nthreads)
                dt = time.perf_counter() - t0
                crays = cnt["primary"] + cnt["shadow"] + cnt["bounce"]
                line.update(cpu_reference_Mrays_s=round(crays / dt / 1e6, 3), cpu_threads=nthreads, cpu_rays_equal=bool(crays == rays[0]),
                            max_abs_vs_cpu=float(np.abs(ref - got).max()),
                            bit_equal_frac_vs_cpu=float((ref.view(np.uint32) == got.view(np.uint32)).all(axis=2).mean()))
            if rank == 0:
                print(json.dumps(line), flush=True)
            barrier()
            R.close()
        scene.close()

    g = load_golden("cornell")
    frame_config("C1 CornellBox-Mirror-Rotated 512x512, 1 light, limit 2", g.flat, g.lights, 512, 512, 2, ["p2p"])
    g = load_golden("monkey")
    frame_config("C2 monkey-rotated 1920x1080, 2 lights, limit 1", g.flat, g.lights, 1920, 1080, 1, ["p2p"])
    flat, lights = ob.dragon_standin_fixture()
    frame_config("C3 dragon stand-in 1920x1080, 1 light, limit 5", flat, lights, 1920, 1080, 5, ["p2p"])
    g = load_golden("dodge")
    frame_config("C5 dodgeColorTest 3840x2160, 3 lights, limit 2", g.flat, g.lights, 3840, 2160, 2, ["p2p", "nccl"])

    # ---- C4: contiguous 1/world chunks of the 16 M rays, full scene replica per rank, no exchange
    if only and "C4" not in only:
        if world > 1:
            dist.destroy_process_group()
        return
    n = 16 * 1024 * 1024
    per = n // world
    flat = ob.random_soup(1_000_000, seed=1234, scale=0.01, smooth_normals=False)
    s = capi.Scene(flat, device=local)
    rays = ob.random_rays(n, seed=5678)[rank * per:(rank + 1) * per]
    md = np.random.default_rng(2).uniform(0, 1, n).astype(np.float32)[rank * per:(rank + 1) * per]
    inf = np.full(per, np.inf, np.float32)
    dR, dH, dM, dO = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    for p, b in ((dR, per * 32), (dH, per * 32), (dM, per * 4), (dO, per)):
        capi.check(lib.cgrt_device_malloc(local, b, C.byref(p)))
    capi.check(lib.cgrt_memcpy_h2d(local, dR, C.c_void_p(rays.ctypes.data), per * 32))
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize(dev)
            best = min(best, allmax(a.elapsed_time(b)))
        return best

    out = dict(config="C4 soup 1 M triangles / 16 777 216 incoherent rays", n_gpus=world, exchange="none (contiguous ray chunks)")
    out["closest_ms"] = round(timed(lambda: capi.check(lib.cgrt_intersect_closest_device(s.h, dR, per, dH, None, st))), 3)
    hits = np.zeros(per, capi.HIT_DTYPE)
    capi.check(lib.cgrt_memcpy_d2h(local, C.c_void_p(hits.ctypes.data), dH, per * 32))
    for label, m in (("any_inf", inf), ("any_u01", md)):
        capi.check(lib.cgrt_memcpy_h2d(local, dM, C.c_void_p(m.ctypes.data), per * 4))
        out[label + "_ms"] = round(timed(lambda: capi.check(lib.cgrt_intersect_any_device(s.h, dR, dM, C.c_float(0.001), per, dO, st))), 3)
    nhit = allsum([float((hits["tri"] >= 0).sum())])[0]
    for k in ("closest", "any_inf", "any_u01"):
        out[k + "_Mrays_s"] = round(n / out[k + "_ms"] / 1e3, 1)
    out["hit_frac"] = round(nhit / n, 4)
    if rank == 0 and world == 1:
        sample = np.random.default_rng(1).choice(per, 100_000, replace=False)
        b = ob.RefLib().scene(flat).bvh(mode=1)
        b.intersect(rays[sample[:2000]], nthreads=nthreads)
        t0 = time.perf_counter()
        g = b.intersect(rays[sample], nthreads=nthreads)
        dt = time.perf_counter() - t0
        out.update(cpu_reference_Mrays_s=round(len(sample) / dt / 1e6, 4), cpu_threads=nthreads, cpu_sample="100 000 of the rays, closest hit",
                   sample_bit_identical=bool(np.array_equal(hits["t"][sample].view(np.uint32), g["t"].view(np.uint32))))
    if rank == 0:
        print(json.dumps(out), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
