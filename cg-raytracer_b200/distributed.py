"""Multi-GPU frame rendering: one process per GPU (torch.distributed over NCCL for the plumbing), full scene replica per
rank, interleaved screen tiles, tiles gathered to rank 0 over NVLink and de-interleaved there by a kernel of the library.

The path partitions by pixels (src/main.cpp:656-697 has no inter-pixel dependence), so there is exactly one exchange step:
the framebuffer gather. Everything compute-side stays inside libcgrt_b200.so; torch is used for device buffers, the stream
and the collective only.
"""
import ctypes as C

import numpy as np

from . import capi as _capi


# ---- host-side mirrors of the tile layout (numpy; used by the CPU/gloo tests of the plumbing) ---------------------------------
def _tile_geometry(params):
    tw, th = (params.tile_w or 8), (params.tile_h or 8)
    tiles_x = -(-params.width // tw)
    return tw, th, tiles_x


def pack_tiles(capi, params, frame):
    """Cut this rank's tile-major buffer out of a full frame [H,W,3] in Screen layout — the layout cgrt_render_device
    writes for world > 1 (tiles in increasing global id, tile_h*tile_w pixels each, padded to the largest rank)."""
    tw, th, tiles_x = _tile_geometry(params)
    W, H = params.width, params.height
    mine = capi.tile_list(params, params.rank)
    out = np.zeros(capi.tile_buffer_floats(params), np.float32).reshape(-1, th, tw, 3)
    for lt, g in enumerate(mine):
        ty, tx = divmod(int(g), tiles_x)
        for q_y in range(th):
            y = ty * th + q_y
            if y >= H:
                continue
            x0, x1 = tx * tw, min(tx * tw + tw, W)
            out[lt, q_y, : x1 - x0] = frame[H - 1 - y, x0:x1]
    return out.reshape(-1)


def assemble_on_host(capi, params, buffers):
    """numpy restatement of cgrt_assemble_tiles: per-rank tile-major buffers -> frame [H,W,3] in Screen layout."""
    tw, th, tiles_x = _tile_geometry(params)
    W, H = params.width, params.height
    frame = np.zeros((H, W, 3), np.float32)
    for r, buf in enumerate(buffers):
        p = _capi.render_params(W, H, params.trace_limit, r, params.world, params.tile_w, params.tile_h)
        tiles = np.asarray(buf, np.float32).reshape(-1, th, tw, 3)
        for lt, g in enumerate(capi.tile_list(p, r)):
            ty, tx = divmod(int(g), tiles_x)
            for q_y in range(th):
                y = ty * th + q_y
                if y >= H:
                    continue
                x0, x1 = tx * tw, min(tx * tw + tw, W)
                frame[H - 1 - y, x0:x1] = tiles[lt, q_y, : x1 - x0]
    return frame


def gather_tiles(local, rank, world):
    """Gather equal-sized per-rank tile buffers to rank 0 (list of tensors there, None elsewhere)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [local]
    out = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
    dist.gather(local, gather_list=out, dst=0)
    return out


class TiledRenderer:
    """Render frames of one scene on `world` GPUs. Rank r renders its interleaved tiles on its own B200; rank 0 receives all
    tile buffers (NCCL gather over NVLink), de-interleaves them into the Screen layout and, for the end-to-end form, copies
    the frame to pinned host memory."""

    def __init__(self, scene, width, height, trace_limit, rank=0, world=1, device=0, tile=(0, 0)):
        import torch
        self.torch = torch
        self.scene = scene
        self.lib = _capi.load_library()
        self.W, self.H, self.L = width, height, trace_limit
        self.rank, self.world, self.device = rank, world, device
        self.tile = tile
        self.dev = torch.device(f"cuda:{device}")
        self.params = _capi.render_params(width, height, trace_limit, rank, world, tile[0], tile[1])
        n = _capi.tile_buffer_floats(self.params)
        self.local = torch.empty(n, dtype=torch.float32, device=self.dev)
        self.frame = torch.empty(height * width * 3, dtype=torch.float32, device=self.dev) if rank == 0 else None
        self.gathered = torch.empty(n * world, dtype=torch.float32, device=self.dev) if (rank == 0 and world > 1) else None
        self.host_frame = None
        self.launches_last = 0

    def _stream(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def render_device(self, cam, flags=0):
        """Enqueue one frame on the current torch stream; returns the device frame tensor on rank 0 (None elsewhere).
        No host synchronisation."""
        p = self.params
        p.flags = flags
        st = self._stream()
        if self.world == 1:
            self.scene.render_device(cam, p, self.frame.data_ptr(), st)
            self.launches_last = 3 * self.L
            return self.frame
        import torch.distributed as dist
        self.scene.render_device(cam, p, self.local.data_ptr(), st)
        if self.rank == 0:
            chunks = list(self.gathered.chunk(self.world))
            dist.gather(self.local, gather_list=chunks, dst=0)
            _capi.check(self.lib.cgrt_assemble_tiles(self.device, C.byref(p), C.c_void_p(self.gathered.data_ptr()),
                                                     C.c_void_p(self.frame.data_ptr()), C.c_void_p(st)))
            return self.frame
        dist.gather(self.local, gather_list=None, dst=0)
        return None

    def render_to_host(self, cam):
        """End to end: per-frame inputs (camera + lights) go host->device inside the call, the finished frame comes back to
        pinned host memory on rank 0. Returns a numpy view [H,W,3] on rank 0."""
        torch = self.torch
        if self.world == 1:
            if self.host_frame is None:
                self.host_frame = torch.empty(self.H * self.W * 3, dtype=torch.float32).pin_memory()
            p = self.params
            p.flags = 0
            _capi.check(self.lib.cgrt_render(self.scene.h, C.byref(cam), C.byref(p), C.c_void_p(self.host_frame.data_ptr()), None))
            return self.host_frame.numpy().reshape(self.H, self.W, 3)
        frame = self.render_device(cam)
        if self.rank == 0:
            if self.host_frame is None:
                self.host_frame = torch.empty(self.H * self.W * 3, dtype=torch.float32).pin_memory()
            self.host_frame.copy_(frame, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return self.host_frame.numpy().reshape(self.H, self.W, 3) if self.rank == 0 else None
