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

    def __init__(self, rank: int = 0, size: int = 1, *, handle=None, group=None, transport=None):
        self.rank = int(rank)
        self.size = int(size)
        self._handle = handle  # pa_comm*
        self.group = group     # torch.distributed group used for bootstrap / test utilities
        # "nccl": ncclSend/ncclRecv carry the staged methods; "ipc": NCCL-free communicator
        # (peer-mapped memory + flag words only); None: no data plane (geometry only)
        self.transport = transport

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
        return f"Comm(rank={self.rank}, size={self.size}, transport={self.transport})"


COMM_SELF = Comm(0, 1)
_WORLD = None


def _flag_window(h, rank, size, dist, mandatory):
    """Collective: export this rank's flag window, import everybody else's.  On any
    failure every rank keeps the NCCL fences (or, for an NCCL-free communicator,
    raises: the flag window is its only signalling path)."""
    fh = C.create_string_buffer(_lib.PA_IPC_HANDLE_BYTES)
    off = C.c_int64()
    ok = lib.pa_comm_flags_export(h, fh, C.byref(off)) == _lib.PA_OK
    err = None if ok else lib.pa_last_error().decode()
    allf = [None] * size
    dist.all_gather_object(allf, (rank, bytes(fh.raw), off.value, ok))
    ok = all(o for (_, _, _, o) in allf)
    if ok:
        for (r, hh, oo, _) in allf:
            if r != rank and lib.pa_comm_flags_import(h, r, hh, oo) != _lib.PA_OK:
                ok = False
                err = lib.pa_last_error().decode()
                break
    oks = [None] * size
    dist.all_gather_object(oks, (ok, err))
    if not all(o for (o, _) in oks):
        if mandatory:
            raise _lib.DeviceError(_lib.PA_ECUDA, "flag window setup failed: " +
                                   "; ".join(str(e) for (o, e) in oks if not o)[:300])
        check(lib.pa_set_tunable(b"nccl_fences", 1))  # all ranks must fence the same way


def comm_world(init_nccl: bool | None = None, transport: str | None = None) -> Comm:
    """``MPI.COMM_WORLD``: built from the torchrun environment.

    ``transport`` (or ``$PA_B200_TRANSPORT``): ``"nccl"`` -- one GPU per rank, NCCL
    communicator (rank 0 generates the unique id; ``torch.distributed`` broadcasts
    it) plus the flag window of the one-sided methods; ``"ipc"`` -- NCCL-free
    communicator over CUDA IPC mappings and flag words only (several ranks may
    share a GPU; ``torch.distributed`` runs on gloo as the side channel);
    ``"auto"`` (default) -- ``"ipc"`` when there are more ranks than GPUs on the
    box, else ``"nccl"``.  On a CPU-only box only the geometry is available.
    """
    import torch
    import torch.distributed as dist

    global _WORLD
    if _WORLD is not None and (init_nccl is None or bool(init_nccl) == (_WORLD.handle is not None)):
        return _WORLD  # MPI.COMM_WORLD is a singleton: do not build a second communicator
    transport = transport or os.environ.get("PA_B200_TRANSPORT", "auto")
    ngpu = torch.cuda.device_count() if torch.cuda.is_available() else 0
    wsize = int(os.environ.get("WORLD_SIZE", "1"))
    if transport == "auto":
        transport = "ipc" if (ngpu > 0 and wsize > ngpu) else "nccl"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not dist.is_initialized():
        if "RANK" not in os.environ:
            return COMM_SELF
        backend = "nccl" if (ngpu > 0 and transport == "nccl") else "gloo"
        if ngpu > 0:
            torch.cuda.set_device(local % ngpu)
        dist.init_process_group(backend)
    rank, size = dist.get_rank(), dist.get_world_size()
    if init_nccl is None:
        init_nccl = ngpu > 0 and size > 1
    handle = None
    if init_nccl:
        torch.cuda.set_device(local % ngpu)
        torch.cuda.current_stream().synchronize()  # make sure the primary context exists
        check(lib.pa_set_device(torch.cuda.current_device()))
        h = C.c_void_p()
        if transport == "ipc":
            check(lib.pa_comm_init_local(size, rank, C.byref(h)))
        else:
            buf = C.create_string_buffer(_lib.PA_UNIQUE_ID_BYTES)
            if rank == 0:
                check(lib.pa_comm_unique_id(buf))
            obj = [bytes(buf.raw)]
            dist.broadcast_object_list(obj, src=0)
            check(lib.pa_comm_init_rank(obj[0], size, rank, C.byref(h)))
        handle = h
        _flag_window(h, rank, size, dist, mandatory=(transport == "ipc"))
    _WORLD = Comm(rank, size, handle=handle, transport=transport if handle else None)
    return _WORLD
