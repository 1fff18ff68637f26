"""``Transpositions``: the reference module's public surface
(src/Transpositions/Transpositions.jl) over libpa_b200.

    t = Transposition(dest, src; method=PointToPoint())   # :93-118
    transpose_(t; waitall=True)                           # transpose!(t; waitall)   :170-179
    Waitall(t)                                            # MPI.Waitall(t)           :127-130
    transpose_(dest, src; method=...)                     # transpose!(dest, src)    :160-168

Python has no ``!`` in identifiers: ``transpose_`` (torch's in-place naming)
stands for ``transpose!``; ``transpose_bang`` is an alias.  Errors follow the
reference: incompatible pencils / extra dims raise ``ArgumentError``.

All work is enqueued on the CURRENT torch CUDA stream and is asynchronous with
respect to the host, like any other CUDA op.
"""
from __future__ import annotations

import ctypes as C
import threading

import torch

from . import _lib
from ._lib import lib, check, i64arr, ArgumentError, PlanInfo, PeerInfo, Timings
from .arrays import PencilArray
from .pencils import Pencil


class AbstractTransposeMethod:
    def __repr__(self):
        return type(self).__name__

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)


class PointToPoint(AbstractTransposeMethod):  # Transpositions.jl:18
    code = _lib.PA_POINT_TO_POINT


class Alltoallv(AbstractTransposeMethod):  # Transpositions.jl:19
    code = _lib.PA_ALLTOALLV


class PeerPut(AbstractTransposeMethod):
    """B200 extension (no reference counterpart): one-sided puts over NVLink.

    Each remote block is packed by a kernel that stores straight into the
    destination rank's ``dest`` array through a peer mapping -- no ``send_buf``,
    no ``recv_buf``, no unpack pass.  ``Transposition(dest, src; method=PeerPut())``
    is COLLECTIVE over the communicator (like ``MPI_Win_create``): it exchanges
    CUDA IPC handles of ``dest``.  Falls back to the staged PointToPoint
    schedule when ``src`` and ``dest`` alias (in-place transposes).
    """
    code = _lib.PA_PEER_PUT


class PeerGet(AbstractTransposeMethod):
    """Pull flavour of :class:`PeerPut`: the unpack kernel of each remote block
    loads straight out of the source rank's ``src`` array over NVLink.  The
    window is on ``src``; ``waitall=False`` defers only the fence that guards
    the reuse of ``src`` -- the role ``MPI.Waitall(t)`` has in the reference."""
    code = _lib.PA_PEER_GET


class _Plan:
    """Owner of one ``pa_plan`` handle (geometry + launch descriptors + streams)."""

    def __init__(self, pin: Pencil, pout: Pencil, extra_dims, elsize: int, method):
        h = C.c_void_p()
        check(lib.pa_plan_create(pin._h, pout._h, len(extra_dims), i64arr(extra_dims), elsize,
                                 method.code, C.byref(h)))
        self.h = h
        info = PlanInfo()
        check(lib.pa_plan_get_info(h, C.byref(info)))
        self.info = info

    def __del__(self):
        try:
            lib.pa_plan_destroy(self.h)
        except Exception:
            pass

    def peer(self, n: int) -> PeerInfo:
        p = PeerInfo()
        check(lib.pa_plan_get_peer(self.h, n, C.byref(p)))
        return p

    def block(self, op: int, n: int = 1):
        d = _lib.BlockDesc()
        check(lib.pa_plan_get_block(self.h, op, n, C.byref(d)))
        return d


def _get_plan(pin: Pencil, pout: Pencil, extra_dims, elsize, method) -> _Plan:
    # plans are cached on the output pencil: `transpose!(dest, src)` builds a
    # fresh Transposition per call in the reference (:165); re-deriving the
    # geometry is cheap there, re-creating streams/events per call is not here.
    key = (id(pin), tuple(extra_dims), int(elsize), method.code)
    hit = pout._plans.get(key)
    if hit is not None and hit[0] is pin:
        return hit[1]
    plan = _Plan(pin, pout, extra_dims, elsize, method)
    pout._plans[key] = (pin, plan)
    return plan


_bound = threading.local()


def _stream_ptr():
    # keep the library's (statically linked) CUDA runtime on torch's current device
    dev = torch.cuda.current_device()
    if getattr(_bound, "dev", None) != dev:
        check(lib.pa_set_device(dev))
        _bound.dev = dev
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _register_window(plan: _Plan, Ao: PencilArray, Ai: PencilArray, comm):
    """Collective: expose ``Ao``'s device buffer to the peers of the grid line and
    map theirs (``pa_ipc_export`` / ``pa_ipc_import`` / ``pa_plan_set_window``).
    Cached per destination array; SPMD programs hit or miss the cache together."""
    import weakref
    import torch.distributed as dist

    lo, hi = Ao.data_ptr(), Ao.data_ptr() + Ao.data.numel() * Ao.elsize
    si, ei = Ai.data_ptr(), Ai.data_ptr() + Ai.data.numel() * Ai.elsize
    if lo < ei and si < hi:
        return  # aliased (in-place): libpa_b200 takes the staged schedule, no window needed
    reg = plan.__dict__.setdefault("_windows", {})
    ent = reg.get(id(Ao))
    if ent is not None and ent[0]() is Ao and ent[1] == lo:
        return
    if not dist.is_initialized():
        raise ArgumentError(_lib.PA_EINVAL, "PeerPut needs torch.distributed for the handle exchange")
    _stream_ptr()  # binds the library to torch's current device
    h = C.create_string_buffer(_lib.PA_IPC_HANDLE_BYTES)
    off = C.c_int64()
    # failures must not leave the other ranks stuck in the collectives below:
    # every rank always takes part in both exchanges and all raise together
    err = None
    st = lib.pa_ipc_export(C.c_void_p(lo), h, C.byref(off))
    if st != _lib.PA_OK:
        err = f"rank {comm.rank}: export failed: {lib.pa_last_error().decode()}"
    allh = [None] * comm.size
    dist.all_gather_object(allh, (comm.rank, bytes(h.raw), off.value, err))
    errs = [e for (_, _, _, e) in allh if e]
    if not errs:
        byrank = {r: (hh, oo) for (r, hh, oo, _) in allh}
        for n in range(1, plan.info.nproc + 1):
            peer = plan.peer(n)
            if peer.is_self:
                continue
            hh, oo = byrank[peer.world_rank]
            mapped = C.c_void_p()
            st = lib.pa_ipc_import(hh, oo, C.byref(mapped))
            if st == _lib.PA_OK:
                st = lib.pa_plan_set_window(plan.h, C.c_void_p(lo), n, mapped)
            if st != _lib.PA_OK:
                err = f"rank {comm.rank}: import from rank {peer.world_rank} failed: " \
                      f"{lib.pa_last_error().decode()}"
                break
    oks = [None] * comm.size
    dist.all_gather_object(oks, err)
    errs += [e for e in oks if e]
    if errs:
        raise _lib.DeviceError(_lib.PA_ECUDA, "one-sided window setup failed: " + "; ".join(errs[:3]))
    reg[id(Ao)] = (weakref.ref(Ao), lo)


class Transposition:
    """Holds data for transposition between two pencil configurations (:69-119)."""

    def __init__(self, Ao: PencilArray, Ai: PencilArray, *, method=None):
        method = PointToPoint() if method is None else method
        Pi, Po = Ai.pencil, Ao.pencil
        if Ai.extra_dims != Ao.extra_dims:  # :99-103
            raise ArgumentError(_lib.PA_EINVAL,
                                "incompatible number of extra dimensions of PencilArrays: "
                                f"{Ai.extra_dims} != {Ao.extra_dims}")
        if Ai.dtype != Ao.dtype:  # PencilArray{T,N} for both arguments
            raise ArgumentError(_lib.PA_EINVAL, f"element types differ: {Ai.dtype} != {Ao.dtype}")
        if Pi.topology is not Po.topology:  # assert_compatible uses `!==` (:182-184)
            raise ArgumentError(_lib.PA_EINCOMPAT, "pencil topologies must be the same.")
        self.Pi, self.Po, self.Ai, self.Ao = Pi, Po, Ai, Ao
        self.method = method
        self._plan = _get_plan(Pi, Po, Ai.extra_dims, Ai.elsize, method)  # remaining checks in C
        d = self._plan.info.dim
        self.dim = None if d == 0 else d  # :110
        if d != 0 and self._plan.info.nproc > 1:
            if isinstance(method, PeerPut):
                _register_window(self._plan, Ao, Ai, Pi.topology.comm)
            elif isinstance(method, PeerGet):
                _register_window(self._plan, Ai, Ao, Pi.topology.comm)

    @property
    def plan(self) -> _Plan:
        return self._plan

    def timings(self) -> Timings:
        t = Timings()
        check(lib.pa_plan_timings(self._plan.h, C.byref(t)))
        return t

    def enable_timing(self, on=True):
        check(lib.pa_plan_enable_timing(self._plan.h, 1 if on else 0))


def Waitall(t: Transposition):
    """``MPI.Waitall(t::Transposition)`` (:127-130): sends done => send_buf reusable."""
    check(lib.pa_wait(t.plan.h, _stream_ptr()))
    return None


def transpose_(*args, method=None, waitall=True, overlap=True, stage_self=False):
    """``transpose!(dest, src; method)`` or ``transpose!(t; waitall)`` (:160-179)."""
    if len(args) == 1 and isinstance(args[0], Transposition):
        t = args[0]
    elif len(args) == 2:
        dest, src = args
        if dest is src:  # same pencil & same data (:164)
            return dest
        t = Transposition(dest, src, method=method)
        waitall = True
    else:
        raise TypeError("transpose_(dest, src; method) or transpose_(t; waitall)")
    flags = (_lib.PA_WAITALL if waitall else 0) | (0 if overlap else _lib.PA_NO_OVERLAP) | (
        _lib.PA_STAGE_SELF if stage_self else 0)
    comm = t.Pi.topology.comm.handle
    check(lib.pa_transpose(t.plan.h, comm, C.c_void_p(t.Ai.data_ptr()),
                           C.c_void_p(t.Ao.data_ptr()), flags, _stream_ptr()))
    return t if len(args) == 1 else args[0]


transpose_bang = transpose_
