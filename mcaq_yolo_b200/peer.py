"""Per-channel range exchange between the GPUs of one node over peer memory (csrc/peer_exchange.cuh).

One `RangeExchange` per scale and rank: a small device buffer that every other rank of the node maps
through CUDA IPC.  K2 of each rank stores its `[min, -max]` vector into every rank's buffer and K3
takes the minimum over ranks, so a batch sharded over R GPUs quantizes with the ranges of the whole
batch (what the single-process reference computes, quantization.py:423-426, 650-654) without a
collective launch between the two kernels.  `torch.distributed` is only used once, at construction,
to swap the 64-byte IPC handles.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

MAX_RANKS = 8


class RangeExchange:
    def __init__(self, C: int, rank: int, world: int, local_ptr: int, peer_ptrs, owner=None):
        self.C, self.rank, self.world = int(C), int(rank), int(world)
        self.local = int(local_ptr)
        self.peers = (ctypes.c_void_p * MAX_RANKS)(*([int(p) for p in peer_ptrs] + [None] * (MAX_RANKS - world)))
        self._owner = owner            # keeps shared state (virtual ranks) alive
        self._opened = []
        self._owns_local = owner is None

    # ------------------------------------------------------------------ construction
    @staticmethod
    def _alloc(C: int, world: int) -> int:
        lib = _lib.load()
        nbytes = lib.mcaq_xchg_bytes(int(C), int(world))
        if nbytes <= 0:
            raise ValueError(f"bad exchange geometry C={C} world={world} (at most {MAX_RANKS} ranks)")
        out = ctypes.c_void_p()
        _lib.check(lib.mcaq_xchg_alloc(nbytes, ctypes.byref(out)), "mcaq_xchg_alloc")
        return int(out.value)

    @classmethod
    def create(cls, C: int, group=None) -> "RangeExchange":
        """Collective over `group` (all ranks on ONE node): allocate, swap IPC handles, map peers."""
        import torch.distributed as dist
        lib = _lib.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())

        def agree(err):
            """All ranks learn whether any of them failed (also a barrier): construction then fails
            on every rank alike, so callers can fall back to the all-reduce path consistently."""
            ok = torch.tensor([0 if err else 1], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            return int(ok.item()) == 1

        local, err = None, None
        handle = ctypes.create_string_buffer(64)
        try:
            local = cls._alloc(C, world)
            _lib.check(lib.mcaq_xchg_export(local, handle), "mcaq_xchg_export")
        except (RuntimeError, ValueError) as e:      # allocation / IPC export failed on this rank
            err = e
        if not agree(err):                           # before all_gather_object: nobody is left waiting in it
            if local is not None:
                lib.mcaq_xchg_free(local)
            raise RuntimeError(f"peer range exchange unavailable on this node ({err or 'failed on another rank'})")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        ptrs, opened, err = [], [], None
        try:
            for r, h in enumerate(handles):
                if r == rank:
                    ptrs.append(local)
                    continue
                out = ctypes.c_void_p()
                _lib.check(lib.mcaq_xchg_open(ctypes.create_string_buffer(h, 64), ctypes.byref(out)), "mcaq_xchg_open")
                ptrs.append(int(out.value))
                opened.append(int(out.value))
        except RuntimeError as e:          # no peer access between two GPUs, IPC disabled, ...
            err = e
        # nobody publishes before every rank has mapped every buffer
        if not agree(err):
            for p in opened:
                lib.mcaq_xchg_close(p)
            lib.mcaq_xchg_free(local)
            raise RuntimeError(f"peer range exchange unavailable on this node ({err or 'failed on another rank'})")
        ex = cls(C, rank, world, local, ptrs)
        ex._opened = opened
        return ex

    @classmethod
    def virtual(cls, C: int, world: int):
        """`world` exchange objects inside ONE process on one GPU (all buffers local): lets a
        single-GPU test drive the protocol with several virtual ranks on one stream."""
        bufs = [cls._alloc(C, world) for _ in range(world)]
        return [cls(C, r, world, bufs[r], bufs, owner=bufs) for r in range(world)]

    # ------------------------------------------------------------------ host-driven halves (tests)
    def publish(self, packed: torch.Tensor):
        from . import ops
        ops._call("mcaq_xchg_publish", ctypes.addressof(self.peers), self.rank, self.world, packed.data_ptr(),
                  self.C, ops._stream())

    def merged(self, device) -> torch.Tensor:
        from . import ops
        out = torch.empty((2 * self.C,), device=device, dtype=torch.float32)
        ops._call("mcaq_xchg_merge", self.local, self.world, self.C, out.data_ptr(), ops._stream())
        return out

    def check(self):
        """Synchronise and raise if a wait of this exchange timed out (a peer died, or the ranks did not
        run the same number of exchange steps); the affected launch used this rank's own ranges."""
        step = ctypes.c_int(0)
        _lib.check(_lib.load().mcaq_xchg_error(self.local, ctypes.byref(step)), "mcaq_xchg_error")
        if step.value:
            raise RuntimeError(f"peer range exchange: wait for step {step.value} timed out on rank {self.rank} "
                               "(peer not running the same step sequence?); that launch used local ranges")

    def close(self):
        """Unmap the peers' buffers and free the local one (peers must have closed their mappings or be
        gone: call after a barrier)."""
        lib = _lib.load()
        for p in self._opened:
            lib.mcaq_xchg_close(p)
        self._opened = []
        if self._owns_local and self.local:
            lib.mcaq_xchg_free(self.local)
            self.local = 0
        # virtual ranks (tests) share their buffers through `_owner`: those few KB are not freed

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
