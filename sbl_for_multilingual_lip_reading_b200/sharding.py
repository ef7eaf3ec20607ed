"""Batch sharding of the visual-encoder path across the GPUs of one box.

Clips are independent in eval mode, so the multi-GPU path is contiguous batch shards with replicated weights and no
data-path collective; the only exchange the reference has is `nn.DataParallel`'s scatter of the clip batch and gather
of the outputs (SBL/train.py:114-115, test.py:163-174).  One process per GPU (`torchrun`); these helpers are
backend-agnostic (NCCL on B200s, gloo in the CPU tests) and carry no arithmetic.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Clip range [begin, end) of `rank`, chunked like `DataParallel.scatter` (= `Tensor.chunk(world)`): chunks of
    ceil(n / world) clips, the last ranks may get fewer or none."""
    if n < 0 or world < 1 or not 0 <= rank < world:
        raise ValueError(f"shard_bounds: bad arguments n={n} world={world} rank={rank}")
    per = -(-n // world) if n else 0
    begin = min(n, rank * per)
    return begin, min(n, begin + per)


def gather_outputs(local_out: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All ranks contribute their shard's encoder output [n_local, T, D] (n_local may differ per rank, `shard_bounds`
    order) and receive the whole batch [n_total, T, D] — the `DataParallel.gather` of the reference on every rank.
    One `all_gather_into_tensor` of equal-sized (zero-padded) blocks."""
    world = dist.get_world_size(group)
    if world == 1:
        return local_out
    per = -(-n_total // world)
    t, d = local_out.shape[1], local_out.shape[2]
    block = local_out
    if local_out.shape[0] != per:
        block = local_out.new_zeros((per, t, d))
        block[:local_out.shape[0]] = local_out
    full = local_out.new_empty((world * per, t, d))
    dist.all_gather_into_tensor(full, block.contiguous(), group=group)
    return full[:n_total]


def max_over_ranks(value: float, device, group=None) -> float:
    """Device-side timing is reported as the MAX over ranks (bench.py)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def set_p2p_timeout_ms(ms: int) -> int:
    """How long a peer-memory gather waits for the other ranks before its watchdog fires (process-wide, default 30 s);
    returns the previous value."""
    from . import _lib
    return int(_lib.load().sblk_set_p2p_timeout_ms(int(ms)))


class P2PGather:
    """One-shot all-gather of every rank's fp32 output block over NVLink / NVSwitch peer memory (`sblk_p2p_gather_fwd`):
    the `DataParallel.gather` of the reference (train.py:114-115) for the one-process-per-GPU layout, as ONE kernel that
    stores the local block into every peer's buffer and returns when all peers' blocks have landed (NCCL all-gather
    completion semantics at a fraction of its latency for MB-sized blocks).  Buffers are cudaMalloc'ed by libsblk and
    mapped into the peers through CUDA IPC handles exchanged over the process group.  Two gather buffers alternate so a
    peer that is one step ahead never overwrites a block that may still be read.

        g = P2PGather(local_elems, device)          # collective: every rank of `group` must call it
        full = g(local_out)                          # -> fp32 [world * local_elems] view of one of the two buffers

    The returned view stays valid only until THIS rank's next call is enqueued: a peer may enter the call after next —
    which stores into the same buffer — as soon as this rank's next gather kernel has published its flag, so every read
    of a result has to be stream-ordered in front of the following call (bench.py copies / consumes it on the same
    stream).  `close()` (collective) unmaps and frees the buffers; the object is also a context manager.  A rank that
    does not show up within `sblk_set_p2p_timeout_ms` (default 30 s) traps the waiting kernel with watchdog code 0x0901.
    """

    def __init__(self, local_elems: int, device, group=None):
        import ctypes
        from . import _lib
        self._lib, self._ct = _lib, ctypes
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device)
        self.local_elems = int(local_elems)
        self.block_bytes = self.local_elems * 4
        if self.block_bytes % 16:
            raise ValueError("P2PGather: the per-rank block must be a multiple of 16 bytes")
        lib = _lib.load()
        self._owned, self._opened = [], []
        # every step below is collective-safe: a rank that fails still takes part in the exchanges, then ALL ranks raise
        mine, err = [], None
        with torch.cuda.device(self.device):
            try:
                for nbytes in (self.world * self.block_bytes, self.world * self.block_bytes, 256, 256):   # 2 buffers, 2 flag arrays
                    ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
                    _lib.check(lib.sblk_p2p_alloc(nbytes, ctypes.byref(ptr), handle), "sblk_p2p_alloc")
                    self._owned.append(ptr.value)
                    mine.append(bytes(handle))
            except Exception as e:  # noqa: BLE001
                err, mine = e, None
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            if any(m is None for m in everyone):
                self._release()
                raise RuntimeError(f"P2PGather: peer-memory buffers could not be allocated on every rank ({err})")
            tables = []
            try:
                for which in range(4):
                    ptrs = []
                    for r in range(self.world):
                        if r == self.rank:
                            ptrs.append(self._owned[which])
                        else:
                            ptr = ctypes.c_void_p()
                            hb = (ctypes.c_ubyte * 64).from_buffer_copy(everyone[r][which])
                            _lib.check(lib.sblk_p2p_open(hb, ctypes.byref(ptr)), "sblk_p2p_open")
                            self._opened.append(ptr.value)
                            ptrs.append(ptr.value)
                    tables.append(torch.tensor(ptrs, dtype=torch.int64, device=self.device))
            except Exception as e:  # noqa: BLE001
                err = e
            oks = [None] * self.world
            dist.all_gather_object(oks, err is None, group=group)
            if not all(oks):
                self._release()
                raise RuntimeError(f"P2PGather: peer buffers could not be mapped on every rank ({err})")
            self._buf_tables, self._flag_tables = tables[0:2], tables[2:4]
            self._counter = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._views = [_DeviceBuffer(self._owned[i], self.world * self.local_elems).as_tensor(self.device)
                           for i in range(2)]
        self._epoch = 0
        dist.barrier(group=group)   # every peer has mapped every buffer before the first store

    def __call__(self, local: torch.Tensor) -> torch.Tensor:
        if local.dtype != torch.float32 or not local.is_cuda or not local.is_contiguous():
            raise RuntimeError("P2PGather: expected a contiguous CUDA fp32 tensor")
        if local.numel() != self.local_elems:
            raise RuntimeError(f"P2PGather: block has {local.numel()} elements, expected {self.local_elems}")
        self._epoch += 1
        b = self._epoch & 1
        with torch.cuda.device(self.device):
            rc = self._lib.load().sblk_p2p_gather_fwd(
                local.data_ptr(), self._buf_tables[b].data_ptr(), self._flag_tables[b].data_ptr(),
                self._counter.data_ptr(), self.rank, self.world, self.block_bytes, (self._epoch + 1) // 2,
                torch.cuda.current_stream().cuda_stream)
        self._lib.check(rc, "sblk_p2p_gather_fwd")
        return self._views[b]

    def _release(self):
        lib = self._lib.load()
        for p in self._opened:
            lib.sblk_p2p_close(p, 1)
        self._opened = []
        dist.barrier(group=self.group)   # nobody frees a buffer a peer still has mapped
        for p in self._owned:
            lib.sblk_p2p_close(p, 0)
        self._owned = []

    def close(self):
        if not self._owned and not self._opened:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self._release()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class _DeviceBuffer:
    """Zero-copy torch view of a raw device allocation (fp32) through __cuda_array_interface__."""

    def __init__(self, ptr, elems):
        self.__cuda_array_interface__ = {"shape": (int(elems),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}

    def as_tensor(self, device):
        return torch.as_tensor(self, device=device)
