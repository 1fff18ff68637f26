"""Communicator: the stand-in for ``MPI.Comm`` on the path.

One process per GPU.  The data plane is an NCCL communicator owned by
libpa_b200 (``pa_comm``); ``torch.distributed`` is used only as a side channel
to hand the 128-byte NCCL unique id from rank 0 to the others (any backend,
``gloo`` included) -- the same job ``mpiexec`` does for the reference.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _lib
from ._lib import lib, check


class Comm:
    """``MPI.Comm`` analogue: rank, size and (lazily) the NCCL communicator."""

    def __init__(self, rank: int = 0, size: int = 1, *, handle=None, group=None):
        self.rank = int(rank)
        self.size = int(size)
        self._handle = handle  # pa_comm*
        self.group = group     # torch.distributed group used for bootstrap / test utilities

    # MPI.Comm_rank / MPI.Comm_size
    def Comm_rank(self):
        return self.rank

    def Comm_size(self):
        return self.size

    @property
    def handle(self):
        return self._handle

    def __del__(self):
        try:
            if self._handle:
                lib.pa_comm_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def __repr__(self):
        return f"Comm(rank={self.rank}, size={self.size}, nccl={'yes' if self._handle else 'no'})"


COMM_SELF = Comm(0, 1)
_WORLD = None


def comm_world(init_nccl: bool | None = None) -> Comm:
    """``MPI.COMM_WORLD``: built from the torchrun environment.

    With ``size > 1`` on a GPU box this creates the NCCL communicator (rank 0
    generates the unique id; ``torch.distributed`` broadcasts it).  On a
    CPU-only box only the geometry is available (``init_nccl=False``).
    """
    import torch
    import torch.distributed as dist

    global _WORLD
    if _WORLD is not None and (init_nccl is None or bool(init_nccl) == (_WORLD.handle is not None)):
        return _WORLD  # MPI.COMM_WORLD is a singleton: do not build a second NCCL communicator
    if not dist.is_initialized():
        if "RANK" not in os.environ:
            return COMM_SELF
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend)
    rank, size = dist.get_rank(), dist.get_world_size()
    if init_nccl is None:
        init_nccl = torch.cuda.is_available() and size > 1
    handle = None
    if init_nccl:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", str(rank % max(1, torch.cuda.device_count())))))
        torch.cuda.current_stream().synchronize()  # make sure the primary context exists
        check(lib.pa_set_device(torch.cuda.current_device()))
        buf = C.create_string_buffer(_lib.PA_UNIQUE_ID_BYTES)
        if rank == 0:
            check(lib.pa_comm_unique_id(buf))
        obj = [bytes(buf.raw)]
        dist.broadcast_object_list(obj, src=0)
        h = C.c_void_p()
        check(lib.pa_comm_init_rank(obj[0], size, rank, C.byref(h)))
        handle = h
        # flag window for the NVLink fences of PeerPut / PeerGet (optional: on any
        # failure every rank keeps the NCCL fences)
        fh = C.create_string_buffer(_lib.PA_IPC_HANDLE_BYTES)
        off = C.c_int64()
        ok = lib.pa_comm_flags_export(h, fh, C.byref(off)) == _lib.PA_OK
        allf = [None] * size
        dist.all_gather_object(allf, (rank, bytes(fh.raw), off.value, ok))
        ok = all(o for (_, _, _, o) in allf)
        if ok:
            for (r, hh, oo, _) in allf:
                if r != rank and lib.pa_comm_flags_import(h, r, hh, oo) != _lib.PA_OK:
                    ok = False
                    break
        oks = [None] * size
        dist.all_gather_object(oks, ok)
        if not all(oks):  # all ranks must fence the same way
            check(lib.pa_set_tunable(b"nccl_fences", 1))
    _WORLD = Comm(rank, size, handle=handle)
    return _WORLD
