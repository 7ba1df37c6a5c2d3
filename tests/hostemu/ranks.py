"""Several emulated RANKS in one process: one host thread per rank, each with its own EmuEngine; the emulated
``sic_exchange`` / ``sic_allreduce_sum`` (emu_runtime.cpp) are the meeting points, exactly where the CUDA library
exchanges over NVLink.  TEST INFRASTRUCTURE ONLY.

    results = run_ranks(3, lambda ctx: body(ctx))     # ctx: EmuDistContext (rank, world, comm, all_reduce_sum, ...)

ctypes releases the GIL during library calls, so the ranks really run concurrently and block in the barriers of the
emulated exchange; a rank that raises makes the others time out there (120 s) instead of hanging the test forever.
"""
from __future__ import annotations

import threading

import torch

from . import load


class EmuDistContext:
    """The slice of safeincave_b200.distributed.DistContext the host code uses, for thread-ranks."""

    def __init__(self, rank, world, shared):
        self.rank, self.world = rank, world
        self.device = torch.device("cpu")
        self._shared = shared
        self.comm = load().sic_emu_make_comm(rank, world)
        self.p2p, self.use_p2p = None, False

    def make_p2p(self, part):
        return None

    def barrier(self):
        if self.world > 1:
            self._shared["barrier"].wait(timeout=300)

    def all_reduce_sum(self, t):
        if self.world == 1:
            return t
        sh = self._shared
        sh["slots"][self.rank] = t.clone()
        sh["barrier"].wait(timeout=300)
        tot = sum(sh["slots"][r] for r in range(self.world))
        sh["barrier"].wait(timeout=300)
        t.copy_(tot)
        return t

    def max_over_ranks(self, value):
        if self.world == 1:
            return value
        sh = self._shared
        sh["slots"][self.rank] = float(value)
        sh["barrier"].wait(timeout=300)
        out = max(sh["slots"][r] for r in range(self.world))
        sh["barrier"].wait(timeout=300)
        return out


def run_ranks(world, body):
    """Run ``body(ctx)`` on ``world`` thread-ranks; returns the list of results, re-raises the first exception."""
    shared = {"barrier": threading.Barrier(world), "slots": [None] * world}
    results, errors = [None] * world, [None] * world

    def target(r):
        try:
            results[r] = body(EmuDistContext(r, world, shared))
        except BaseException as e:      # noqa: BLE001 -- reported to the caller below
            errors[r] = e
            shared["barrier"].abort()

    threads = [threading.Thread(target=target, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=1800)
    for e in errors:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in errors:
        if e is not None:
            raise e
    if any(t.is_alive() for t in threads):
        raise TimeoutError("an emulated rank is still running")
    return results
