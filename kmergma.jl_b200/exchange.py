"""Host-side exchange of the per-shard result blocks of a multi-GPU scan on ONE node.

north_star: "per-GPU hit lists are copied back and merged on the host in genome order" -- no collective on the path.  Every
rank (one process per GPU) packs its (run, extension result) block with kgma_result_pack straight into its slice of a POSIX
shared-memory segment and publishes a step number; rank 0 waits for all step numbers and replays the blocks in place
(kgma_replay_packed reads the segment directly).  A few KB per rank and step, no device round trip: the NCCL all-gather this
replaces cost ~0.1 ms per step (H2D of the block, the collective, D2H, a stream synchronisation).

Layout: int64 published[world] | int64 consumed | 2 slots x world blocks of `cap` bytes (steps alternate between the slots, so a
rank may pack step s+1 while rank 0 still reads step s; it waits only if rank 0 is two steps behind).
Stores of one x86 core become visible in program order, and the flag is written after the block, so a reader that sees the
flag sees the block (the box's hosts are x86-64; on a weakly ordered host a release fence would be needed here)."""
import time
from multiprocessing import shared_memory
from typing import Optional

import numpy as np

_HDR_WORDS = 64          # published[<= 62], consumed, reserved  (one cache-line multiple)


class HostExchange:
    def __init__(self, name: str, rank: int, world: int, cap: int, create: bool):
        if world > _HDR_WORDS - 2:
            raise ValueError("at most %d ranks" % (_HDR_WORDS - 2))
        self.rank, self.world, self.cap = rank, world, (int(cap) + 4095) // 4096 * 4096
        size = _HDR_WORDS * 8 + 2 * world * self.cap
        self._shm = shared_memory.SharedMemory(name=name, create=create, size=size if create else 0)
        if not create:
            # the segment belongs to rank 0: keep this process's resource tracker from unlinking it a second time at exit
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        self._buf = np.ndarray((size,), dtype=np.uint8, buffer=self._shm.buf)
        self._hdr = self._buf[:_HDR_WORDS * 8].view(np.int64)
        if create:
            self._hdr[:] = 0
        self._base = self._buf.ctypes.data + _HDR_WORDS * 8
        self.step = 0
        self._owner = create

    # -- every rank -----------------------------------------------------------------------------------------------------
    def begin_step(self, timeout: float = 30.0) -> int:
        """address of this rank's block for the next step (cap bytes); blocks only while rank 0 is two steps behind"""
        self.step += 1
        t0 = None
        while self._hdr[_HDR_WORDS - 2] < self.step - 2:
            t0 = t0 or time.monotonic()
            if time.monotonic() - t0 > timeout:
                raise TimeoutError("rank 0 did not consume step %d" % (self.step - 2))
        return self._base + ((self.step & 1) * self.world + self.rank) * self.cap

    def block(self) -> np.ndarray:
        """numpy view of this rank's block of the current step"""
        o = _HDR_WORDS * 8 + ((self.step & 1) * self.world + self.rank) * self.cap
        return self._buf[o:o + self.cap]

    def publish(self) -> None:
        self._hdr[self.rank] = self.step

    # -- rank 0 ---------------------------------------------------------------------------------------------------------
    def wait_all(self, timeout: float = 30.0) -> int:
        """address of the world blocks of the current step (stride cap), once every rank has published it"""
        pub = self._hdr[:self.world]
        t0 = None
        while int(pub.min()) < self.step:
            t0 = t0 or time.monotonic()
            if time.monotonic() - t0 > timeout:
                raise TimeoutError("ranks %s did not publish step %d" % (np.nonzero(pub < self.step)[0].tolist(), self.step))
        return self._base + (self.step & 1) * self.world * self.cap

    def blocks(self) -> np.ndarray:
        o = _HDR_WORDS * 8 + (self.step & 1) * self.world * self.cap
        return self._buf[o:o + self.world * self.cap].reshape(self.world, self.cap)

    def consumed(self) -> None:
        self._hdr[_HDR_WORDS - 2] = self.step

    def close(self) -> None:
        self._hdr = self._buf = None
        try:
            self._shm.close()
            if self._owner:
                self._shm.unlink()
        except Exception:
            pass
