"""Data-parallel plumbing of the hot path (the reference itself is single-device): one process per GPU,
`torch.distributed` for the rendezvous, and

  * `PeerExchange` -- the ranks' gradient buffers mapped into each other's address space (CUDA IPC over NVLink), used by
    the fused all-reduce + Adam kernel `nmx_allreduce_adam` (include/nmx.h) that replaces
    ncclAllReduce -> scale -> Adam inside the captured training iteration;
  * `ray_tile` / `gather_tiles` -- the contiguous ray-tile partition of a frame and its reassembly, used by
    `rendering.render(..., process_group=)` (rendering/render.py:243-345 is the single-device loop being sharded).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib_loader as L


def ray_tile(n_rays, rank, world):
    """Contiguous tile [start, stop) of `n_rays` rays owned by `rank`: equal tiles of ceil(n/world) rays, the last
    ones shorter (possibly empty)."""
    per = (int(n_rays) + world - 1) // world
    start = min(rank * per, int(n_rays))
    return start, min(start + per, int(n_rays))


def gather_tiles(local, n_rays, group=None):
    """all_gather of per-rank ray tiles [tile, ...] (ragged last tiles) -> the assembled [n_rays, ...] tensor on every
    rank.  Works on any backend (NCCL on the box, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    per = (int(n_rays) + world - 1) // world
    tail = tuple(local.shape[1:])
    pad = torch.zeros((per,) + tail, dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world, per) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather(list(out.unbind(0)), pad, group=group)
    return out.reshape((world * per,) + tail)[:int(n_rays)]


class _DevMem:
    """A device allocation owned by libnmx, exposed to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerExchange:
    """`n_bufs` gradient buffers of `numel` fp32 each plus one flags block in ONE cudaMalloc'd allocation per rank,
    every rank's allocation opened in every other rank through CUDA IPC.  Requires all ranks on one NVLink box."""

    FLAGS_BYTES = 512

    def __init__(self, numel, n_bufs=2, device=None, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise L.NmxError("PeerExchange needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 8:
            raise L.NmxError("PeerExchange supports up to 8 ranks (one NVSwitch box)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.numel = int(numel)
        self.n_bufs = int(n_bufs)
        self.buf_bytes = (self.numel * 4 + 255) // 256 * 256
        self.nbytes = self.FLAGS_BYTES + self.n_bufs * self.buf_bytes
        assert int(L.lib().nmx_p2p_flags_bytes()) <= self.FLAGS_BYTES
        with torch.cuda.device(self.device):
            base = ctypes.c_void_p()
            handle = (ctypes.c_ubyte * 64)()
            L.call("nmx_p2p_alloc", L.i64(self.nbytes), ctypes.byref(base), handle)
            self._base = int(base.value)
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self._peer_base = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self._peer_base.append(self._base)
                    continue
                p = ctypes.c_void_p()
                L.call("nmx_p2p_open", (ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(p))
                self._peer_base.append(int(p.value))
            self._mem = torch.as_tensor(_DevMem(self._base, self.nbytes), device=self.device)
        self.flags = self._mem[:self.FLAGS_BYTES].view(torch.int64)
        self._bufs = [self._mem[self.FLAGS_BYTES + i * self.buf_bytes:][:self.numel * 4].view(torch.float32)
                      for i in range(self.n_bufs)]
        PP = ctypes.c_void_p * self.world
        self._grad_ptrs = [PP(*[b + self.FLAGS_BYTES + i * self.buf_bytes for b in self._peer_base])
                           for i in range(self.n_bufs)]
        self._flag_ptrs = PP(*self._peer_base)
        self._closed = False
        dist.barrier(group=group)  # every rank has mapped every peer before the first kernel touches peer memory

    def buffer(self, i):
        """This rank's gradient buffer i ([numel] fp32, device memory peers can read)."""
        return self._bufs[i]

    def allreduce_adam(self, i, p, m, v, lr, b1, b2, eps, lr_dev=None, g_avg=None):
        """Average buffer i over the ranks and apply the MLX-style Adam update to the local replica (p, m, v)."""
        L.require_cuda(p, m, v, lr_dev, g_avg)
        L.call("nmx_allreduce_adam", self._grad_ptrs[i], self._flag_ptrs, L.i32(self.rank), L.i32(self.world), L.ptr(p),
               L.ptr(m), L.ptr(v), L.ptr(g_avg), L.i64(self.numel), L.f32(lr), L.ptr(lr_dev), L.f32(b1), L.f32(b2),
               L.f32(eps), L.stream())

    def error(self):
        """0 = ok; 1 / 2 = a rank never arrived at the entry / exit barrier of some exchange (kernel timed out)."""
        return int(self.flags[18].item())

    def epoch(self):
        return int(self.flags[16].item())

    def close(self):
        if self._closed:
            return
        self._closed = True
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody unmaps while a peer's kernel may still read
        for r, b in enumerate(self._peer_base):
            if r != self.rank:
                L.lib().nmx_p2p_close(ctypes.c_void_p(b))
        self._bufs, self.flags, self._mem = [], None, None
        L.lib().nmx_p2p_free(ctypes.c_void_p(self._base))
