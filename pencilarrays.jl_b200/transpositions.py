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


_tunables = {}


def set_tunable(name: str, value: int):
    """``pa_set_tunable`` + a host-side record (the mirror needs to know whether the
    staged methods run over this library's own NVLink copy kernels)."""
    check(lib.pa_set_tunable(name.encode(), int(value)))
    _tunables[name] = int(value)


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
        self._ipc_handles = []   # every pa_ipc_import this plan holds a reference for
        self._windows = {}       # key -> registered local pointer
        self._arena_ptr = None   # recv_buf the peers' arena windows were registered against

    def __del__(self):
        try:
            lib.pa_plan_destroy(self.h)  # (synchronises nothing: the mappings go after it)
            for hh in self._ipc_handles:
                lib.pa_ipc_release(hh)
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


def _exchange_windows(plan: _Plan, key, ptr: int, comm, setter):
    """Collective over the communicator (like ``MPI_Win_create``): expose the device
    allocation at ``ptr`` (0: this rank owns nothing -- it exports nothing but still
    maps its peers') to the peers of the plan's grid line, map theirs and hand each
    mapped pointer to ``setter(n, mapped)``.

    Whether anything has to be (re)mapped is decided COLLECTIVELY: a rank whose own
    pointer is unchanged still takes part when a peer re-allocated its array.
    Failures never leave ranks stuck in a collective: all raise together."""
    import torch.distributed as dist

    if not dist.is_initialized():
        raise ArgumentError(_lib.PA_EINVAL, "one-sided / peer-memory transposes need "
                            "torch.distributed for the handle exchange")
    _stream_ptr()  # binds the library to torch's current device
    changed = plan._windows.get(key) != ptr
    h = C.create_string_buffer(_lib.PA_IPC_HANDLE_BYTES)
    off = C.c_int64()
    err = None
    if ptr:
        st = lib.pa_ipc_export(C.c_void_p(ptr), h, C.byref(off))
        if st != _lib.PA_OK:
            err = f"rank {comm.rank}: export failed: {lib.pa_last_error().decode()}"
    allh = [None] * comm.size
    dist.all_gather_object(allh, (comm.rank, bytes(h.raw) if ptr else None, off.value, err, changed))
    errs = [e for (_, _, _, e, _) in allh if e]
    if not errs and not any(c for (_, _, _, _, c) in allh):
        return  # every rank still holds current mappings
    if not errs:
        byrank = {r: (hh, oo) for (r, hh, oo, _, _) in allh}
        for n in range(1, plan.info.nproc + 1):
            peer = plan.peer(n)
            if peer.is_self:
                continue
            hh, oo = byrank[peer.world_rank]
            if hh is None:
                continue  # that rank owns nothing: nothing will be put to / got from it
            mapped = C.c_void_p()
            st = lib.pa_ipc_import(hh, oo, C.byref(mapped))
            if st == _lib.PA_OK:
                plan._ipc_handles.append(hh)
                st = setter(n, mapped)
            if st != _lib.PA_OK:
                err = f"rank {comm.rank}: import from rank {peer.world_rank} failed: " \
                      f"{lib.pa_last_error().decode()}"
                break
    oks = [None] * comm.size
    dist.all_gather_object(oks, err)
    errs += [e for e in oks if e]
    if errs:
        raise _lib.DeviceError(_lib.PA_ECUDA, "peer window setup failed: " + "; ".join(errs[:3]))
    plan._windows[key] = ptr


def _register_window(plan: _Plan, Ao: PencilArray, Ai: PencilArray, comm):
    """Window of a one-sided method on ``Ao`` (``dest`` for PeerPut, ``src`` for
    PeerGet): ``pa_ipc_export`` / ``pa_ipc_import`` / ``pa_plan_set_window``."""
    lo, hi = Ao.data_ptr(), Ao.data_ptr() + Ao.data.numel() * Ao.elsize
    si, ei = Ai.data_ptr(), Ai.data_ptr() + Ai.data.numel() * Ai.elsize
    # same rule as libpa_b200 (same base pointer, or overlapping ranges); views of one
    # ManyPencilArray alias on every rank, empty ones included
    aliased = (lo < ei and si < hi) or (lo != 0 and lo == si) or \
        (Ao._owner is not None and Ao._owner is Ai._owner)
    if aliased:
        # in place: libpa_b200 takes the staged schedule.  (Aliasing is a property of the
        # ManyPencilArray, the same on every rank, so skipping the collective is symmetric.)
        return _register_arenas(plan, Ao.pencil if plan.info.method == _lib.PA_PEER_PUT
                                else Ai.pencil, comm)
    ptr = lo if Ao.data.numel() > 0 else 0
    _register_window_ptr(plan, ptr, comm)


def _register_window_ptr(plan: _Plan, ptr: int, comm):
    _exchange_windows(plan, ("win", ptr), ptr, comm,
                      lambda n, mapped: lib.pa_plan_set_window(plan.h, C.c_void_p(ptr), n, mapped))


def _uses_ipc_exchange(comm) -> bool:
    return comm.transport == "ipc" or bool(_tunables.get("ipc_exchange"))


def _register_arenas(plan: _Plan, Po: Pencil, comm):
    """Own-kernel exchange of the staged methods: reserve the arenas at the maximum
    size over the ranks (so that every rank re-allocates -- or not -- at the same
    call), then expose ``recv_buf`` to the peers (``pa_plan_set_recv_window``)."""
    import torch.distributed as dist

    if plan.info.dim == 0 or plan.info.nproc == 1 or not _uses_ipc_exchange(comm):
        return
    _stream_ptr()
    sizes = [None] * comm.size
    dist.all_gather_object(sizes, (plan.info.send_bytes, plan.info.recv_bytes))
    check(lib.pa_pencil_reserve(Po._h, max(1, max(s for s, _ in sizes)),
                                max(1, max(r for _, r in sizes))))
    fam = Po._family
    plans = fam.__dict__.setdefault("_ipc_plans", [])
    if not any(p is plan for p in plans):
        plans.append(plan)
    _, _, rp, _ = Po.buffers()
    for pl in plans:  # every plan sharing these arenas (same list, same order on every rank)
        _exchange_windows(pl, "arena", rp, comm,
                          lambda n, mapped, pl=pl: lib.pa_plan_set_recv_window(pl.h, n, mapped))


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
        if d != 0 and self._plan.info.nproc > 1 and Pi.topology.comm.handle is not None:
            comm = Pi.topology.comm
            if isinstance(method, PeerPut):
                _register_window(self._plan, Ao, Ai, comm)
            elif isinstance(method, PeerGet):
                _register_window(self._plan, Ai, Ao, comm)
            else:
                _register_arenas(self._plan, Po, comm)

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


def transpose_(*args, method=None, waitall=True, overlap=True, stage_self=False, fft=None):
    """``transpose!(dest, src; method)`` or ``transpose!(t; waitall)`` (:160-179).

    ``fft="forward"`` / ``"backward"`` (B200 extension): the unpack and the
    1-d complex FFT along ``dest``'s contiguous dimension -- the step a PencilFFTs-style
    plan runs next -- execute as one kernel; ``dest`` receives the transformed array."""
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
    if fft is not None:
        if fft in ("forward", -1):
            flags |= _lib.PA_FFT_FORWARD
        elif fft in ("backward", 1):
            flags |= _lib.PA_FFT_BACKWARD
        else:
            raise ArgumentError(_lib.PA_EINVAL, "fft must be 'forward' or 'backward'")
    comm = t.Pi.topology.comm.handle
    # (an empty local array may have a null data pointer: the library accepts that)
    check(lib.pa_transpose(t.plan.h, comm, C.c_void_p(t.Ai.data_ptr() or None),
                           C.c_void_p(t.Ao.data_ptr() or None), flags, _stream_ptr()))
    return t if len(args) == 1 else args[0]


transpose_bang = transpose_


def fft_(u: PencilArray, direction="forward"):
    """In-place 1-d complex FFT of ``u`` along its contiguous (first memory) dimension: the
    first step of a PencilFFTs-style 3-d transform (the other two are ``transpose_(t,
    fft=...)``).  Same kernel as the fused unpack+FFT, without a transposition."""
    t = Transposition(u, u)
    return transpose_(t, fft=direction).Ao


def transpose_host_(t: Transposition, host_src: torch.Tensor, host_dst: torch.Tensor):
    """``transpose!`` on HOST arrays through ``pa_transpose_host``: upload, kernels and
    download pipelined inside the library; returns when ``host_dst`` is valid.  Pin the
    tensors (``pin_memory()``) for full PCIe bandwidth."""
    _stream_ptr()
    n_in, n_out = t.plan.info.length_in * t.Ai.elsize, t.plan.info.length_out * t.Ao.elsize
    if host_src.numel() * host_src.element_size() != n_in or \
            host_dst.numel() * host_dst.element_size() != n_out or \
            host_src.is_cuda or host_dst.is_cuda or \
            not host_src.is_contiguous() or not host_dst.is_contiguous():
        raise _lib.DimensionMismatch(_lib.PA_EDIM, "host arrays must be dense CPU tensors of the "
                                     "local array sizes")
    check(lib.pa_transpose_host(t.plan.h, t.Pi.topology.comm.handle,
                                C.c_void_p(host_src.data_ptr() if n_in else None),
                                C.c_void_p(host_dst.data_ptr() if n_out else None), 0))
    return host_dst


class HostChain:
    """A sequence of transpositions applied to host arrays (``pa_host_chain_*``): what a
    PencilFFTs-style plan over ``Array``-backed pencils does around its transposes.

        chain = HostChain([t_xy, t_yz, t_zy, t_yx])
        k = chain.submit(hin, hout)      # asynchronous, double-buffered on the device
        chain.wait(k)                    # hout valid

    Only the pencils / methods of the transpositions are used; the device buffers belong
    to the chain (for one-sided methods their windows are registered here, collectively).
    """

    def __init__(self, transpositions):
        ts = list(transpositions)
        self.ts = ts
        n = len(ts)
        arr = (C.c_void_p * n)(*[t.plan.h.value for t in ts])
        comm = ts[0].Pi.topology.comm
        h = C.c_void_p()
        _stream_ptr()
        check(lib.pa_host_chain_create(n, arr, comm.handle, C.byref(h)))
        self.h = h
        for slot in range(4):
            probe = C.c_void_p()
            check(lib.pa_host_chain_buffer(h, slot, 0, C.byref(probe), None))
            if not probe.value:
                break  # the chain has fewer staging sets (tunable "host_slots")
            for i, t in enumerate(ts):
                info = t.plan.info
                if info.dim == 0 or info.nproc == 1 or comm.handle is None:
                    continue
                if info.method in (_lib.PA_PEER_PUT, _lib.PA_PEER_GET):
                    which = (1 - i % 2) if info.method == _lib.PA_PEER_PUT else (i % 2)
                    p = C.c_void_p()
                    check(lib.pa_host_chain_buffer(h, slot, which, C.byref(p), None))
                    _register_window_ptr(t.plan, p.value, comm)
                else:
                    _register_arenas(t.plan, t.Po, comm)

    def submit(self, host_src: torch.Tensor, host_dst: torch.Tensor) -> int:
        k = C.c_int64()
        check(lib.pa_host_chain_submit(self.h, C.c_void_p(host_src.data_ptr()),
                                       C.c_void_p(host_dst.data_ptr()), C.byref(k)))
        return k.value

    def wait(self, ticket: int = -1):
        check(lib.pa_host_chain_wait(self.h, ticket))

    def time_begin(self):
        check(lib.pa_host_chain_time_begin(self.h))

    def time_end(self) -> float:
        ms = C.c_float()
        check(lib.pa_host_chain_time_end(self.h, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            lib.pa_host_chain_destroy(self.h)
        except Exception:
            pass
