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


class _DevicePtr:
    """A raw device allocation of the library presented to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (int(ptr), False), "version": 3}


class TiledRenderer:
    """Render frames of one scene on `world` GPUs: full scene replica per rank, interleaved screen tiles, one exchange step.

    mode "p2p" (default for world > 1): rank 0 owns the frame in device memory it exports over CUDA IPC; every other rank maps
    it (NVLink / NVSwitch peer memory) and the kernels of cgrt_render_device store that rank's pixels straight at their final
    Screen position (CGRT_RENDER_SCREEN_LAYOUT) - the gather is fused into the shading stores. What is left of the collective
    is two sequence-number flags per rank: "my pixels of frame k have landed" (rank r -> rank 0, cgrt_flag_signal on the
    peer-mapped flag, cgrt_flag_wait on rank 0) and "frame k-1 has been consumed" (rank 0 -> rank r) so a fast rank cannot
    overwrite a frame that rank 0 is still copying out.
    mode "nccl": every rank renders into a tile-major buffer, NCCL gather to rank 0, cgrt_assemble_tiles de-interleaves.
    torch.distributed is the plumbing in both modes (handle exchange / gather); all compute is inside libcgrt_b200.so."""

    TIMEOUT_MS = 4000

    def __init__(self, scene, width, height, trace_limit, rank=0, world=1, device=0, tile=(0, 0), mode=None):
        import torch
        self.torch = torch
        self.scene = scene
        self.lib = _capi.load_library()
        self.W, self.H, self.L = width, height, trace_limit
        self.rank, self.world, self.device = rank, world, device
        self.tile = tile
        self.dev = torch.device(f"cuda:{device}")
        self.params = _capi.render_params(width, height, trace_limit, rank, world, tile[0], tile[1])
        self.host_frame = None
        self.launches_last = 0
        self.seq = 0
        self.mode = "single" if world == 1 else (mode or "p2p")
        self.fallback_reason = None
        self._owned, self._mapped, self._pinned = [], [], []
        if self.mode == "p2p":
            # _init_p2p runs the same collective sequence on every rank whatever fails locally and agrees on the outcome
            self.fallback_reason = self._init_p2p()
            if self.fallback_reason is not None:  # no peer mapping between these processes: the NCCL gather is the other GPU path
                self._release()
                self.mode = "nccl"
        if self.mode != "p2p":
            n = _capi.tile_buffer_floats(self.params)
            self.local = torch.empty(n, dtype=torch.float32, device=self.dev)
            self.frame = torch.empty(height * width * 3, dtype=torch.float32, device=self.dev) if rank == 0 else None
            self.gathered = torch.empty(n * world, dtype=torch.float32, device=self.dev) if (rank == 0 and world > 1) else None

    # ---- p2p set-up: rank 0 exports frame + arrival flags, every rank exports its "consumed" flag ------------------------
    def _malloc(self, nbytes):
        p = C.c_void_p()
        _capi.check(self.lib.cgrt_device_malloc(self.device, nbytes, C.byref(p)))
        self._owned.append(p)
        _capi.check(self.lib.cgrt_memset_device(self.device, p, 0, nbytes, None))
        return p

    def _export(self, ptr):
        h = (C.c_uint8 * _capi.IPC_HANDLE_BYTES)()
        _capi.check(self.lib.cgrt_peer_export(self.device, ptr, h))
        return bytes(h)

    def _open(self, handle):
        h = (C.c_uint8 * _capi.IPC_HANDLE_BYTES).from_buffer_copy(handle)
        p = C.c_void_p()
        _capi.check(self.lib.cgrt_peer_open(self.device, h, C.byref(p)))
        self._mapped.append(p)
        return p

    def _agree(self, err):
        """One all_reduce after each phase of the set-up: every rank learns whether ANY rank failed, so all of them take the
        same branch and issue the same collectives (a rank that failed locally still takes part)."""
        import torch.distributed as dist
        ok = self.torch.tensor([0 if err else 1], dtype=self.torch.int32, device=self.dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return int(ok.item()) == 1

    def _init_p2p(self):
        """Returns None when every rank has its peer mappings, else the reason for the fallback (the same decision on every rank)."""
        torch = self.torch
        import torch.distributed as dist
        nfl = self.H * self.W * 3
        HB = _capi.IPC_HANDLE_BYTES
        err = None
        mine = [bytes(HB)] * 4
        # phase 1 (local): allocate + export
        try:
            self.consumed = self._malloc(256)  # [0] consumed sequence number of this rank
            # wait-timeout counter: pinned host memory the wait kernel writes through its device alias, so that checking it after
            # a stream synchronisation is a plain host read (no copy in the timed region)
            st = C.c_void_p()
            _capi.check(self.lib.cgrt_host_alloc_pinned(64, C.byref(st)))
            self._pinned.append(st)
            C.memset(st, 0, 64)
            self.status = st
            mine[0] = self._export(self.consumed)
            if self.rank == 0:
                # two frames: frame k lives in buffer k & 1, so that the copy-out of frame k can overlap the rendering of k + 1
                self.frame_ptrs = [self._malloc(nfl * 4), self._malloc(nfl * 4)]
                self.arrive = self._malloc(4 * max(self.world, 64))
                mine[1], mine[2], mine[3] = self._export(self.frame_ptrs[0]), self._export(self.arrive), self._export(self.frame_ptrs[1])
            _capi.check(self.lib.cgrt_device_synchronize(self.device))
        except Exception as e:
            err = f"{type(e).__name__}: {e}"
        if not self._agree(err):
            return err or "a peer rank could not allocate / export its buffers"
        # phase 2 (collective): exchange the handles
        t = torch.tensor(list(b"".join(mine)), dtype=torch.uint8, device=self.dev)
        allh = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(allh, t)
        allh = [bytes(x.cpu().tolist()) for x in allh]
        # phase 3 (local): map the peers' buffers
        try:
            if self.rank == 0:
                self.peer_consumed = [None] + [self._open(allh[r][0:HB]) for r in range(1, self.world)]
                self.frames = [torch.as_tensor(_DevicePtr(q.value, nfl), device=self.dev) for q in self.frame_ptrs]
                self.out_ptrs = self.frame_ptrs
            else:
                self.frames = [None, None]
                self.out_ptrs = [self._open(allh[0][HB:2 * HB]), self._open(allh[0][3 * HB:4 * HB])]
                self.peer_arrive = self._open(allh[0][2 * HB:3 * HB])
        except Exception as e:
            err = f"{type(e).__name__}: {e}"
        if not self._agree(err):
            return err or "a peer rank could not map the exported buffers"
        self._copy_done = [None, None]  # rank 0, streaming: event after the copy-out of the frame that last used the buffer
        dist.barrier(device_ids=[self.device])
        return None

    def _release(self):
        """Unmap peer buffers, free this rank's device allocations and pinned blocks (idempotent)."""
        try:
            self.lib.cgrt_device_synchronize(self.device)
        except Exception:
            pass
        for q in self._mapped:
            self.lib.cgrt_peer_close(self.device, q)
        for q in self._owned:
            self.lib.cgrt_device_free(self.device, q)
        for q in self._pinned:
            self.lib.cgrt_host_free_pinned(q)
        self._mapped, self._owned, self._pinned = [], [], []

    def close(self):
        """Collective-free teardown; call it on every rank after the last frame (peers must have stopped writing: barrier first)."""
        self.frames = [None, None]
        self.frame = None
        self._release()

    def _check_handoff(self):
        """After a stream synchronisation: a hand-off wait that gave up means a frame with missing peer tiles (or a buffer
        overwritten while it was being copied out) - an error, not a statistic."""
        n = self.timeouts()
        if n:
            raise RuntimeError(f"rank {self.rank}: {n} exchange hand-off wait(s) timed out after {self.TIMEOUT_MS} ms; the frame is not trustworthy")

    def _stream(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def _render_p2p(self, cam, flags):
        p = self.params
        p.flags = flags | _capi.RENDER_SCREEN_LAYOUT
        st = C.c_void_p(self._stream())
        lib, dv = self.lib, self.device
        self.seq += 1
        seq = self.seq
        b = seq & 1
        free = max(seq - 2, 0)  # the frame that used buffer b before
        if self.rank == 0:
            # the consumer of frame seq-2 precedes this signal: either it was enqueued on this stream (stream order), or it is
            # the streaming copy-out, whose event the stream waits for first
            if self._copy_done[b] is not None:
                self.torch.cuda.current_stream(self.dev).wait_event(self._copy_done[b])
                self._copy_done[b] = None
            ptrs = (C.c_void_p * (self.world - 1))(*[q.value for q in self.peer_consumed[1:]])
            _capi.check(lib.cgrt_flag_signal(dv, ptrs, self.world - 1, free, st))
            self.scene.render_device(cam, p, self.out_ptrs[b].value, st.value)
            _capi.check(lib.cgrt_flag_wait(dv, C.c_void_p(self.arrive.value + 4), self.world - 1, seq, self.TIMEOUT_MS, self.status, st))
            self.frame = self.frames[b]
            return self.frame
        _capi.check(lib.cgrt_flag_wait(dv, self.consumed, 1, free, self.TIMEOUT_MS, self.status, st))
        self.scene.render_device(cam, p, self.out_ptrs[b].value, st.value)
        ptrs = (C.c_void_p * 1)(self.peer_arrive.value + 4 * self.rank)
        _capi.check(lib.cgrt_flag_signal(dv, ptrs, 1, seq, st))
        return None

    def timeouts(self):
        """Number of hand-off waits that gave up so far (0 in a healthy run). A host read of the pinned counter: meaningful
        after the stream has been synchronised."""
        if self.mode != "p2p":
            return 0
        return int(C.cast(self.status, C.POINTER(C.c_uint32))[0])

    def render_device(self, cam, flags=0):
        """Enqueue one frame on the current torch stream; returns the device frame tensor on rank 0 (None elsewhere).
        No host synchronisation."""
        p = self.params
        st = self._stream()
        if self.world == 1:
            p.flags = flags
            self.scene.render_device(cam, p, self.frame.data_ptr(), st)
            return self.frame
        if self.mode == "p2p":
            return self._render_p2p(cam, flags)
        p.flags = flags
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

    def count_pass(self, cam):
        """One frame through the counting variants (reference test counts for the roofline arithmetic) into a private
        tile-major buffer: no exchange step, no effect on the shared frame. Returns the scene's stats."""
        if getattr(self, "_count_buf", None) is None:
            self._count_buf = self.torch.empty(_capi.tile_buffer_floats(self.params), dtype=self.torch.float32, device=self.dev)
        p = self.params
        p.flags = _capi.RENDER_COUNT
        self.scene.render_device(cam, p, self._count_buf.data_ptr(), self._stream())
        return self.scene.collect_stats()

    def extra_launches_per_frame(self):
        """Library kernels of the exchange step per frame on this rank (flag signal + wait, or the assemble kernel)."""
        if self.world == 1:
            return 0
        if self.mode == "p2p":
            return 2
        return 1 if self.rank == 0 else 0

    def stream_to_host(self, cam, n_frames):
        """Streaming: n_frames frames end to end with two frames in flight, so that the device->host copy of frame k overlaps
        the kernels of frame k+1; returns the last frame [H,W,3] on rank 0 once every frame has been delivered.
        One GPU: cgrt_render_submit / cgrt_render_wait. Several GPUs (p2p mode): rank 0 owns two frames (buffer k & 1), copies
        them out on a second stream, and only frees a buffer for the peers ("consumed" flag) when its copy has finished."""
        torch = self.torch
        if getattr(self, "_stream_bufs", None) is None and self.rank == 0:
            self._stream_bufs = [torch.empty(self.H * self.W * 3, dtype=torch.float32).pin_memory() for _ in range(2)]
        if self.world == 1:
            p = self.params
            p.flags = 0
            for k in range(n_frames):
                self.scene.render_submit(cam, p, self._stream_bufs[k & 1].data_ptr())
            self.scene.render_wait()
            return self._stream_bufs[(n_frames - 1) & 1].numpy().reshape(self.H, self.W, 3)
        if self.mode != "p2p":  # the gather mode has no second frame: synchronous frames
            out = None
            for _ in range(n_frames):
                out = self.render_to_host(cam)
            return out
        main = torch.cuda.current_stream(self.dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        last = None
        for k in range(n_frames):
            frame = self.render_device(cam)
            if self.rank == 0:
                b = self.seq & 1
                done = torch.cuda.Event()
                done.record(main)
                self._copy_stream.wait_event(done)
                with torch.cuda.stream(self._copy_stream):
                    self._stream_bufs[b].copy_(frame, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                self._copy_done[b] = ev
                last = self._stream_bufs[b]
            if (k & 15) == 15:
                main.synchronize()  # bound the host's run-ahead (per-frame parameter ring of the library)
        main.synchronize()
        if self.rank == 0:
            self._copy_stream.synchronize()
        self._check_handoff()
        if self.rank == 0:
            return last.numpy().reshape(self.H, self.W, 3)
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
        if self.mode == "p2p":
            self._check_handoff()
        return self.host_frame.numpy().reshape(self.H, self.W, 3) if self.rank == 0 else None
