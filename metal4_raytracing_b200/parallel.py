"""Multi-GPU frame sharding: one process per GPU, scene + BVH replicated, interleaved 16x16 screen tiles.

The reference is single-device (MetalRaytracing/Renderer.swift:229); the unit of ownership here is its 16x16
threadgroup tile (Renderer.swift:1445-1451): rank g of N renders tiles with tile % N == g. The EMA history of a
pixel lives on the rank that owns it (Raytracing.metal:796-819 only ever reads the same pixel), so the only exchange
is the assembly of the displayed frame:

  * "peer"   — rt_trace stores owned pixels straight into every rank's frame over NVLink peer mappings (CUDA IPC),
               compute and exchange in one kernel; ranks then meet at a stream-ordered NCCL barrier;
  * "gather" — rt_pack_tiles -> torch.distributed all_gather_into_tensor (NCCL) -> rt_unpack_tiles.

A third mode partitions samples instead of pixels (SURVEY.md §8e's alternative for multi-spp frames):

  * "samples" — rank g traces the samples s of every pixel with s % N == g (rt_trace_options.sampleModulo) and writes
               its share of the frame; one NCCL all-reduce (sum) of the rgba32f frame makes every rank hold the frame, which
               is also the history of the next one. Per-rank work is 1/N of the rays over the whole screen (perfectly
               balanced, no ownership pattern), the exchange is the full frame every frame (33 MB at 1080p), and float
               sums are reassociated: equal to the single-GPU frame to rounding, not bit for bit.

The tile arithmetic and the slab layout are plain host logic and are shared with the CPU (gloo) tests.
"""
import ctypes as C

import numpy as np

TILE = 16


def tile_grid(width, height):
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


def owned_tiles(width, height, world_size, rank):
    tx, ty = tile_grid(width, height)
    return list(range(rank, tx * ty, world_size))


def slab_tiles(width, height, world_size):
    tx, ty = tile_grid(width, height)
    return (tx * ty + world_size - 1) // world_size


def owner_mask(width, height, world_size, rank):
    """Boolean (height, width) mask of the pixels rank `rank` owns."""
    tx, _ = tile_grid(width, height)
    ys, xs = np.mgrid[0:height, 0:width]
    tile = (ys // TILE) * tx + (xs // TILE)
    return (tile % world_size) == rank


def pack_tiles_host(image, world_size, rank):
    """numpy restatement of rt_pack_tiles: (H, W, C) -> (slab_tiles, 256, C), missing pixels zero."""
    h, w = image.shape[:2]
    tx, _ = tile_grid(w, h)
    n = slab_tiles(w, h, world_size)
    slab = np.zeros((n, TILE * TILE) + image.shape[2:], image.dtype)
    for k, t in enumerate(owned_tiles(w, h, world_size, rank)):
        x0, y0 = (t % tx) * TILE, (t // tx) * TILE
        block = image[y0:y0 + TILE, x0:x0 + TILE]
        full = np.zeros((TILE, TILE) + image.shape[2:], image.dtype)
        full[:block.shape[0], :block.shape[1]] = block
        slab[k] = full.reshape((TILE * TILE,) + image.shape[2:])
    return slab


def unpack_tiles_host(slabs, width, height, world_size):
    """numpy restatement of rt_unpack_tiles: (world, slab_tiles, 256, C) -> (H, W, C)."""
    tx, ty = tile_grid(width, height)
    out = np.zeros((height, width) + slabs.shape[3:], slabs.dtype)
    for t in range(tx * ty):
        r, k = t % world_size, t // world_size
        x0, y0 = (t % tx) * TILE, (t // tx) * TILE
        block = slabs[r, k].reshape((TILE, TILE) + slabs.shape[3:])
        hh, ww = min(TILE, height - y0), min(TILE, width - x0)
        out[y0:y0 + hh, x0:x0 + ww] = block[:hh, :ww]
    return out


def gather_frame_host(local_image, world_size, rank, dist):
    """Frame assembly with torch.distributed on host tensors (gloo): the N>1 host logic under test on CPU."""
    import torch
    h, w = local_image.shape[:2]
    slab = pack_tiles_host(local_image, world_size, rank)
    raw = np.ascontiguousarray(slab).view(np.uint8).reshape(-1)
    mine = torch.from_numpy(raw.copy())
    parts = [torch.empty_like(mine) for _ in range(world_size)]
    dist.all_gather(parts, mine)
    allslabs = np.stack([p.numpy().view(local_image.dtype).reshape(slab.shape) for p in parts])
    return unpack_tiles_host(allslabs, w, h, world_size)


class FrameExchange:
    """Device-side frame assembly for a `device.Renderer` under torch.distributed (NCCL).

    Stream ordering: the library enqueues on the context's own stream (rt_get_stream), which need not be torch's
    current stream. Every collective here is issued with that stream made current (a torch ExternalStream over the
    same cudaStream_t), so pack -> all-gather -> unpack, and peer stores -> barrier, are ordered on the one stream the
    kernels run on, whatever stream the caller's torch code uses. `close()` unmaps the imported peer images."""

    def __init__(self, renderer, world_size, rank, mode="peer"):
        import torch
        import torch.distributed as dist
        from . import _abi as A
        from . import device as D
        self.r, self.world, self.rank, self.mode = renderer, world_size, rank, mode
        self.dist, self.torch, self.A, self.D = dist, torch, A, D
        self.ctx = renderer.ctx
        self._flag = torch.zeros(1, device=f"cuda:{self.ctx.device}")
        self._peers = None  # [image slot][rank] -> device pointer
        self._imported = []  # peer mappings this rank opened (rt_ipc_close at close())
        L = D.lib()
        L.rt_pack_tiles.argtypes = [C.c_void_p, C.POINTER(A.Image), C.c_void_p, C.c_int, C.c_int]
        L.rt_unpack_tiles.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(A.Image), C.c_int]
        L.rt_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        L.rt_ipc_import.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.rt_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
        if world_size > 1 and mode == "peer":
            self._open_peers()
        if world_size > 1 and mode == "samples":
            info = renderer.image_info(A.TEXTURE_ACCUMULATION)
            if info.format != A.FORMAT_RGBA32_FLOAT:
                raise ValueError('exchange mode "samples" sums shares of the frame: create the Renderer with fp32=True')
            self._sum = torch.empty(renderer.width * renderer.height * 4, dtype=torch.float32,
                                    device=f"cuda:{self.ctx.device}")
        if world_size > 1 and mode == "gather":
            info = renderer.image_info(A.TEXTURE_ACCUMULATION)
            bpp = {A.FORMAT_RGBA16_FLOAT: 8, A.FORMAT_RGBA32_FLOAT: 16}[info.format]
            n = slab_tiles(renderer.width, renderer.height, world_size) * 256 * bpp
            self._slab = torch.empty(n, dtype=torch.uint8, device=f"cuda:{self.ctx.device}")
            self._all = torch.empty(n * world_size, dtype=torch.uint8, device=f"cuda:{self.ctx.device}")

    def _open_peers(self):
        """Exchange CUDA IPC handles of both accumulation images; keyed by local device pointer."""
        A, D, L = self.A, self.D, self.D.lib()
        mine = {}
        for slot in (A.TEXTURE_ACCUMULATION, A.TEXTURE_PREVIOUS_ACCUMULATION):
            ptr = self.r.image_info(slot).data
            buf = C.create_string_buffer(64)
            D._check(L.rt_ipc_export(self.ctx._h, ptr, buf))
            mine[slot] = (ptr, buf.raw)
        everyone = [None] * self.world
        self.dist.all_gather_object(everyone, mine)
        # local pointer of slot s on this rank -> list of every rank's pointer for the same slot
        self._peers = {}
        for slot, (ptr, _) in mine.items():
            ptrs = []
            for rk in range(self.world):
                if rk == self.rank:
                    ptrs.append(ptr)
                else:
                    out = C.c_void_p()
                    D._check(L.rt_ipc_import(self.ctx._h, everyone[rk][slot][1], C.byref(out)))
                    ptrs.append(out.value)
                    self._imported.append(out.value)
            self._peers[ptr] = ptrs

    def _library_stream(self):
        """torch view of the stream the library enqueues on right now (it may have been changed with set_stream)."""
        return self.torch.cuda.ExternalStream(self.ctx.stream, device=self.ctx.device)

    def close(self):
        """Unmaps the peer images (collective: every rank must have finished using them)."""
        if self._imported:
            self.ctx.sync()
            if self.dist.is_initialized():
                self.dist.barrier()
            L = self.D.lib()
            for p in self._imported:
                L.rt_ipc_close(self.ctx._h, p)
            self._imported = []
        self._peers = None

    def draw_partition(self):
        """Keyword arguments for Renderer.draw that give this rank its share of the frame under the chosen mode."""
        if self.world == 1:
            return {}
        if self.mode == "samples":
            return {"sample_modulo": self.world, "sample_remainder": self.rank}
        return {"tile_modulo": self.world, "tile_remainder": self.rank, "peers": self.peers_for_next_draw()}

    def peers_for_next_draw(self):
        """Peer pointers matching the image the next draw writes (TextureIndexPreviousAccumulation)."""
        if self.world == 1 or self.mode != "peer":
            return None
        dst = self.r.image_info(self.A.TEXTURE_PREVIOUS_ACCUMULATION).data
        return self._peers[dst]

    def finish_frame(self):
        """After draw(): make the full frame visible at TextureIndexAccumulation on every rank."""
        if self.world == 1:
            return
        A, L = self.A, self.D.lib()
        with self.torch.cuda.stream(self._library_stream()):  # NCCL orders itself against the *current* stream
            if self.mode == "samples":
                img = self.r.image_info(A.TEXTURE_ACCUMULATION)
                nbytes = self._sum.numel() * 4
                self.ctx.copy(self._sum.data_ptr(), img.data, nbytes)
                self.dist.all_reduce(self._sum)  # sum of the ranks' shares = the frame (and the next frame's history)
                self.ctx.copy(img.data, self._sum.data_ptr(), nbytes)
            elif self.mode == "gather":
                img = self.r.image_info(A.TEXTURE_ACCUMULATION)
                self.D._check(L.rt_pack_tiles(self.ctx._h, C.byref(img), self._slab.data_ptr(), self.world, self.rank))
                self.dist.all_gather_into_tensor(self._all, self._slab)
                self.D._check(L.rt_unpack_tiles(self.ctx._h, self._all.data_ptr(), C.byref(img), self.world))
            else:
                # peer stores are complete when every rank's kernel has finished: stream-ordered barrier
                self.dist.all_reduce(self._flag)
