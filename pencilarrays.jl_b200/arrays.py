"""``PencilArray`` / ``ManyPencilArray``: the array wrappers `transpose!`
receives (src/arrays.jl:81-138, src/multiarrays.jl:106-140), over CUDA device
memory.  A ``torch.Tensor`` is used purely as the owner of the device buffer.

Layout contract (identical to the reference): ``parent(u)`` is dense,
column-major (first index fastest), with dims
``(size_local(pencil, MemoryOrder())..., extra_dims...)``.  A row-major torch
tensor with the REVERSED shape has exactly that byte layout, so ``parent(u)``
is a torch tensor of shape ``(*reversed(extra_dims), *reversed(mem_dims))``;
``u.jl()`` gives a (strided) view indexed in Julia dim order and
``u.logical()`` one indexed in logical order (``u[i,j,k]`` of the reference).
"""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import DimensionMismatch, ArgumentError
from .pencils import Pencil, MemoryOrder, LogicalOrder, size_local as _p_size_local, _is_mem
from .permutations import as_tuple


def _device():
    if not torch.cuda.is_available():
        raise _lib.DeviceError(_lib.PA_ENOGPU, "PencilArray storage lives in CUDA device memory")
    return torch.device("cuda", torch.cuda.current_device())


class PencilArray:
    def __init__(self, pencil: Pencil, data: torch.Tensor, extra_dims=None):
        mem = _p_size_local(pencil, MemoryOrder())
        N = len(mem)
        if data.dim() < N:
            raise DimensionMismatch(_lib.PA_EDIM,
                                    f"array has {data.dim()} dimensions, pencil has {N}")
        jl_dims = tuple(reversed(data.shape))  # Julia (column-major) dims of this buffer
        if extra_dims is None:
            extra_dims = jl_dims[N:]
        extra_dims = tuple(int(e) for e in extra_dims)
        if jl_dims != mem + extra_dims:  # arrays.jl:108-114
            raise DimensionMismatch(
                _lib.PA_EDIM, f"array has incorrect dimensions: {jl_dims}; expected "
                f"{mem + extra_dims} (memory order + extra dims)")
        if not data.is_contiguous():
            raise ArgumentError(_lib.PA_EINVAL, "parent array must be dense")
        self.pencil = pencil
        self.data = data
        self.extra_dims = extra_dims
        self.space_dims = _p_size_local(pencil, LogicalOrder())
        # arrays carved out of one ManyPencilArray alias each other BY CONSTRUCTION, also on
        # a rank where one of them is empty (every rank must then take the same schedule)
        self._owner = None
        self._base_ptr = 0

    @classmethod
    def undef(cls, dtype, pencil: Pencil, *extra_dims, device=None):
        """``PencilArray{T}(undef, pencil, extra_dims...)`` (arrays.jl:134-138)."""
        mem = _p_size_local(pencil, MemoryOrder())
        shape = tuple(reversed(mem + tuple(extra_dims)))
        data = torch.empty(shape, dtype=dtype, device=device or _device())
        return cls(pencil, data, tuple(extra_dims))

    # ---- views ----
    def jl(self) -> torch.Tensor:
        """View with Julia's dim order: ``u.jl()[i,j,k] == parent(u)[i+1,j+1,k+1]``."""
        nd = self.data.dim()
        return self.data.permute(*reversed(range(nd)))

    def logical(self) -> torch.Tensor:
        """View indexed in logical order: ``u.logical()[I] == u[I .+ 1]`` (arrays.jl:327-337)."""
        N = len(self.space_dims)
        perm = as_tuple(self.pencil.perm, N)  # memory dim m holds logical dim perm[m]
        v = self.jl()
        order = [0] * v.dim()
        for m, l in enumerate(perm):
            order[l - 1] = m
        for j in range(N, v.dim()):
            order[j] = j
        return v.permute(*order)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def elsize(self):
        return self.data.element_size()

    def data_ptr(self):
        p = self.data.data_ptr()
        return p if p else self._base_ptr  # an empty view still names its buffer

    def __len__(self):
        return self.data.numel()

    def __repr__(self):
        return (f"PencilArray{{{self.dtype}}}(size_local={self.space_dims}, "
                f"extra_dims={self.extra_dims}, perm={self.pencil.perm})")


# ---- accessor functions, named as in arrays.jl / size.jl ----
def parent(u: PencilArray):
    return u.data


def pencil(u: PencilArray):
    return u.pencil


def extra_dims(u: PencilArray):
    return u.extra_dims


def ndims_extra(u: PencilArray):
    return len(u.extra_dims)


def size_local(u, order=LogicalOrder()):
    if isinstance(u, Pencil):
        return _p_size_local(u, order)
    return _p_size_local(u.pencil, order) + u.extra_dims


def similar(u: PencilArray, pencil_: Pencil = None, dtype=None):
    """``similar(u, [T], [pencil])`` (arrays.jl:246-300)."""
    return PencilArray.undef(dtype or u.dtype, pencil_ or u.pencil, *u.extra_dims,
                             device=u.data.device)


class ManyPencilArray:
    """Several PencilArray views over ONE buffer sized for the largest pencil
    (multiarrays.jl:106-140) -- what in-place transposes operate on."""

    def __init__(self, dtype, *pencils, extra_dims=(), device=None):
        extra_dims = tuple(extra_dims)
        n = max(math.prod(_p_size_local(p)) for p in pencils) * math.prod(extra_dims)
        self.data = torch.empty(max(n, 1), dtype=dtype, device=device or _device())
        self.arrays = []
        for p in pencils:
            mem = _p_size_local(p, MemoryOrder())
            cnt = math.prod(mem) * math.prod(extra_dims)
            view = self.data[:cnt].view(tuple(reversed(mem + extra_dims)))
            a = PencilArray(p, view, extra_dims)
            a._owner, a._base_ptr = self, self.data.data_ptr()
            self.arrays.append(a)

    def __getitem__(self, i):  # 1-based like A[1], A[2] of the reference
        if i < 1 or i > len(self.arrays):
            raise IndexError("index must be in 1:%d" % len(self.arrays))
        return self.arrays[i - 1]

    def __len__(self):
        return len(self.arrays)

    def first(self):
        return self.arrays[0]

    def last(self):
        return self.arrays[-1]
