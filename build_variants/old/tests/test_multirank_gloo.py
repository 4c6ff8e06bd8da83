"""CPU, world_size 2 over gloo: the multi-rank plumbing of bench.py (tile ownership from the C ABI's host arithmetic, gather
to rank 0, de-interleave) with the CPU oracle standing in for the renderer — no GPU, no compute call into the product."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, W, H, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    import bench
    from conftest import load_golden
    from oracle import bindings as ob
    capi = ge.load_package().capi
    g = load_golden("cornell")
    full, _ = ob.OracleLib().scene(g.flat, g.lights).bvh().render(ob.default_camera(W, H), W, H, trace_limit=2, nthreads=2)
    p = capi.render_params(W, H, 2, rank, world)
    # this rank's tile-major buffer, cut from the oracle frame exactly as the render kernel lays it out
    local = bench.pack_tiles(capi, p, full)
    gathered = bench.gather_tiles(torch.from_numpy(local), rank, world)  # list of per-rank buffers on rank 0
    if rank == 0:
        frame = bench.assemble_on_host(capi, p, [t.numpy() for t in gathered])
        np.save(out_path, np.stack([frame, full]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("W,H", [(96, 64), (101, 37)])
def test_two_rank_gather_and_assemble(tmp_path, W, H):
    port = 29500 + (os.getpid() % 2000)
    out = str(tmp_path / "frames.npy")
    mp.spawn(_worker, args=(2, port, W, H, out), nprocs=2, join=True)
    frame, full = np.load(out)
    assert np.array_equal(frame.view(np.uint32), full.view(np.uint32))
    assert full.any()
