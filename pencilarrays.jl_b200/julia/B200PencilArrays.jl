# B200PencilArrays.jl -- Julia veneer over libpa_b200 (include/pa_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia (and no
# MPI).  The same C ABI is exercised call-for-call by the Python/ctypes host
# mirror and its tests; this file is the binding a PencilArrays.jl maintainer
# would load next to CUDA.jl.  It plugs in at the reference's own seam --
# multiple dispatch on the array type carried by the `Pencil`
# (Pencils.jl:282-304) -- in two ways:
#
#   (A) kernel level: the "generic array" methods of the hot path
#       copy_range!   (Transpositions.jl:568-583)  -> pa_box_copy (K1, no allocation)
#       _permutedims! (Transpositions.jl:648-664)  -> pa_box_copy (K2, ONE pass, no temp)
#       so that the reference's own transpose_send!/transpose_recv! loops and a
#       CUDA-aware MPI keep driving the exchange;
#   (B) whole path: transpose!(t::Transposition; waitall) (Transpositions.jl:170-179)
#       -> pa_transpose (pack / NCCL exchange / unpack pipelined on CUDA streams),
#       MPI.Waitall(t) -> pa_wait.
module B200PencilArrays

using PencilArrays
using PencilArrays: Pencils, Transpositions, MemoryOrder, LogicalOrder
using PencilArrays.Transpositions: Transposition, PointToPoint, Alltoallv, AbstractTransposeMethod
using StaticPermutations
using CUDA
import MPI

const libpa = get(ENV, "PA_B200_LIB", "libpa_b200.so")

# ---- status -> exception (pa_b200.h: pa_status) ------------------------------
function check(status::Cint)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:pa_last_error, libpa), Cstring, ()))
    what = unsafe_string(ccall((:pa_strerror, libpa), Cstring, (Cint,), status))
    status in (1, 2) && throw(ArgumentError("$what: $msg"))       # PA_EINVAL, PA_EINCOMPAT
    status == 3 && throw(DimensionMismatch("$what: $msg"))        # PA_EDIM
    error("libpa_b200: $what: $msg")
end

stream_ptr() = reinterpret(Ptr{Cvoid}, CUDA.stream().handle)
devptr(x::CuArray) = reinterpret(Ptr{Cvoid}, pointer(x))

# ---- (A) kernel-level overrides ------------------------------------------------
# column-major strides (in elements) of dims
function colstrides(dims::NTuple{N,Int}) where {N}
    s = ones(Int64, N)
    for i in 2:N
        s[i] = s[i - 1] * dims[i - 1]
    end
    s
end

# pack: strided sub-box of the memory-order parent -> contiguous send/recv buffer
function Transpositions.copy_range!(
        dest::CuVector{T}, dest_offset::Integer,
        src::PencilArray{T,N,<:CuArray}, src_range_memorder::NTuple,
    ) where {T,N}
    src_p = parent(src)
    exdims = extra_dims(src)
    ranges = (src_range_memorder..., map(Base.OneTo, exdims)...)
    ext = Int64[length(r) for r in ranges]
    sstr = colstrides(size(src_p))
    dstr = colstrides(Tuple(ext))
    soff = sum((first(r) - 1) * s for (r, s) in zip(ranges, sstr))
    check(ccall((:pa_box_copy, libpa), Cint,
        (Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        length(ext), ext, sstr, dstr, sizeof(T),
        devptr(src_p) + soff * sizeof(T), devptr(dest) + dest_offset * sizeof(T),
        stream_ptr(), C_NULL))
    dest
end

# unpack: dense block (dims in Pi memory order) -> permuted sub-box of parent(dst),
# one pass, no temporary (the reference allocates `tmp` and copies twice, :659-661)
function Transpositions._permutedims!(
        ::Type{<:CuArray}, v::SubArray{T,N,<:CuArray}, src::CuArray{T,N}, perm,
    ) where {T,N}
    E = N - length(perm)
    pperm = Tuple(append(perm, Val(E)))          # v[k] = src[j], k[i] = j[pperm[i]]
    ext = Int64[size(src)...]
    sstr = colstrides(size(src))
    pstr = colstrides(size(parent(v)))
    dstr = zeros(Int64, N)
    for i in 1:N
        dstr[pperm[i]] = pstr[i]
    end
    doff = sum((first(r) - 1) * s for (r, s) in zip(parentindices(v), pstr))
    check(ccall((:pa_box_copy, libpa), Cint,
        (Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        N, ext, sstr, dstr, sizeof(T), devptr(src), devptr(parent(v)) + doff * sizeof(T),
        stream_ptr(), C_NULL))
    v
end

# ---- (B) whole-path override ------------------------------------------------------
mutable struct Handles
    topo::Ptr{Cvoid}
    pencils::IdDict{Any,Ptr{Cvoid}}
    plans::Dict{Any,Ptr{Cvoid}}
    comm::Ptr{Cvoid}
end
const HANDLES = IdDict{Any,Handles}()   # keyed by MPITopology (identity, like `!==` at :182)

function handles(topo)
    get!(HANDLES, topo) do
        dims = Int64[size(topo)...]
        rank = MPI.Comm_rank(get_comm(topo))
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:pa_topology_create, libpa), Cint, (Cint, Ptr{Int64}, Cint, Ptr{Ptr{Cvoid}}),
                    length(dims), dims, rank, h))
        # NCCL bootstrap over the MPI communicator the user already has
        id = zeros(UInt8, 128)
        rank == 0 && check(ccall((:pa_comm_unique_id, libpa), Cint, (Ptr{UInt8},), id))
        MPI.Bcast!(id, 0, get_comm(topo))
        c = Ref{Ptr{Cvoid}}()
        check(ccall((:pa_set_device, libpa), Cint, (Cint,), CUDA.deviceid(CUDA.device())))
        check(ccall((:pa_comm_init_rank, libpa), Cint, (Ptr{UInt8}, Cint, Cint, Ptr{Ptr{Cvoid}}),
                    id, MPI.Comm_size(get_comm(topo)), rank, c))
        setup_flag_window!(c[], get_comm(topo))
        Handles(h[], IdDict{Any,Ptr{Cvoid}}(), Dict{Any,Ptr{Cvoid}}(), c[])
    end
end

# Flag window of the one-sided methods (pa_comm_flags_export / _import): every rank
# exports a few 64-bit words per source rank and maps everybody else's; window-open /
# window-close of PeerPut / PeerGet then ride inside the transfer kernel as NVLink
# signals.  Collective; if ANY rank fails, all keep the NCCL fences.
function setup_flag_window!(comm_handle::Ptr{Cvoid}, comm::MPI.Comm)
    rank, nranks = MPI.Comm_rank(comm), MPI.Comm_size(comm)
    handle = zeros(UInt8, 64)
    off = Ref{Int64}(0)
    ok = ccall((:pa_comm_flags_export, libpa), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int64}),
               comm_handle, handle, off) == 0
    handles_all = MPI.Allgather(handle, comm)
    offs_all = MPI.Allgather([off[]], comm)
    ok = MPI.Allreduce(ok, &, comm)
    if ok
        for r in 0:(nranks - 1)
            r == rank && continue
            ok &= ccall((:pa_comm_flags_import, libpa), Cint, (Ptr{Cvoid}, Cint, Ptr{UInt8}, Int64),
                        comm_handle, r, view(handles_all, 64r+1:64r+64), offs_all[r + 1]) == 0
        end
    end
    MPI.Allreduce(ok, &, comm) ||
        check(ccall((:pa_set_tunable, libpa), Cint, (Cstring, Int64), "nccl_fences", 1))
    nothing
end

function pencil_handle(H::Handles, p::Pencil{N}) where {N}
    get!(H.pencils, p) do
        perm = permutation(p)
        permv = isidentity(perm) ? C_NULL : Cint[Tuple(perm)...]
        # pencils sharing send_buf (Pencils.jl:265-270) share the device arenas
        share = C_NULL
        for (q, hq) in H.pencils
            q.send_buf === p.send_buf && (share = hq; break)
        end
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:pa_pencil_create, libpa), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Cint}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
            H.topo, N, Int64[size_global(p)...], Cint[decomposition(p)...], permv, share, h))
        h[]
    end
end

method_code(::PointToPoint) = Cint(0)
method_code(::Alltoallv) = Cint(1)

# B200 extensions (no reference counterpart): one-sided transposition methods.
# `PeerPut`: every remote block is stored by ONE kernel straight into the
# destination rank's `dest` over NVLink; `PeerGet`: pull flavour.  Usage:
#     t = Transposition(dest, src; method = B200PencilArrays.PeerPut())
#     B200PencilArrays.register_window!(t)      # collective, once per (dest | src) array
#     transpose!(t)
struct PeerPut <: AbstractTransposeMethod end
struct PeerGet <: AbstractTransposeMethod end
method_code(::PeerPut) = Cint(2)
method_code(::PeerGet) = Cint(3)

# Collective over the communicator, like MPI_Win_create: exchanges CUDA IPC
# handles of the window array (dest for PeerPut, src for PeerGet).
function register_window!(t::Transposition)
    plan, H = plan_handle(t)
    A = t.method isa PeerGet ? parent(t.Ai) : parent(t.Ao)
    comm = get_comm(t.Pi.topology)
    handle = zeros(UInt8, 64)
    off = Ref{Int64}(0)
    check(ccall((:pa_ipc_export, libpa), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int64}), devptr(A), handle, off))
    handles_all = MPI.Allgather(handle, comm)            # 64 bytes per rank
    offs_all = MPI.Allgather([off[]], comm)
    R = t.dim
    R === nothing && return t
    # ranks (in `comm`) of my grid line along R: topology.ranks maps Cartesian
    # coordinates to ranks (MPITopologies.jl:84-86,208-226); get_remote_indices (:539-549)
    topo = t.Pi.topology
    coords = topo.coords_local
    line = [topo.ranks[ntuple(i -> i == R ? n : coords[i], length(coords))...] for n in 1:size(topo)[R]]
    me = MPI.Comm_rank(comm)
    for (n, r) in enumerate(line)
        r == me && continue
        mapped = Ref{Ptr{Cvoid}}()
        check(ccall((:pa_ipc_import, libpa), Cint, (Ptr{UInt8}, Int64, Ptr{Ptr{Cvoid}}),
                    view(handles_all, 64r+1:64r+64), offs_all[r + 1], mapped))
        check(ccall((:pa_plan_set_window, libpa), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                    plan, devptr(A), n, mapped[]))
        # (each import holds a reference on the mapping: pair it with pa_ipc_release(handle)
        #  when the window array is freed)
    end
    t
end

function plan_handle(t::Transposition{T}) where {T}
    H = handles(t.Pi.topology)
    ex = Int64[extra_dims(t.Ai)...]
    key = (objectid(t.Pi), objectid(t.Po), Tuple(ex), sizeof(T), method_code(t.method))
    plan = get!(H.plans, key) do
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:pa_plan_create, libpa), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int64}, Cint, Cint, Ptr{Ptr{Cvoid}}),
            pencil_handle(H, t.Pi), pencil_handle(H, t.Po), length(ex), ex, sizeof(T),
            method_code(t.method), h))
        h[]
    end
    plan, H
end

const DeviceTransposition{T,N} = Transposition{T,N,<:Pencil,<:Pencil,
    <:PencilArray{T,N,<:CuArray},<:PencilArray{T,N,<:CuArray}}

# transpose!(t; waitall) (Transpositions.jl:170-179).
# `fft = :forward | :backward` (B200 extension, PA_FFT_FORWARD / PA_FFT_BACKWARD): the unpack
# and the 1-d FFT along dest's contiguous dimension -- the next step of a PencilFFTs plan --
# run as ONE kernel (ComplexF64, power-of-two lines of 8..1024 points, staged methods).
function Transpositions.transpose!(t::DeviceTransposition; waitall = true, fft = nothing)
    plan, H = plan_handle(t)
    flags = waitall ? Cuint(1) : Cuint(0)        # PA_WAITALL
    fft === :forward && (flags |= Cuint(8))
    fft === :backward && (flags |= Cuint(16))
    ptr_or_null(A) = isempty(A) ? C_NULL : devptr(A)   # a rank may own nothing (Pencils.jl:193-218)
    check(ccall((:pa_transpose, libpa), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint, Ptr{Cvoid}),
        plan, H.comm, ptr_or_null(parent(t.Ai)), ptr_or_null(parent(t.Ao)), flags, stream_ptr()))
    t
end

# In-place 1-d FFT of `u` along its contiguous dimension (same kernel, no transposition): with
# the two fused transposes this is a whole PencilFFTs-style 3-d transform in three kernels.
fft_lines!(u::PencilArray{T,N,<:CuArray}; direction = :forward) where {T,N} =
    (Transpositions.transpose!(Transposition(u, u); fft = direction); u)

# ---- host arrays: a chain of transpositions on `Array`-backed data ------------------
# pa_host_chain_*: one submit = upload, every transpose! on the device, download;
# asynchronous, double-buffered (download of one submit || upload of the next).
mutable struct HostChain
    h::Ptr{Cvoid}
end
function HostChain(ts::Vector{<:Transposition})
    plans = Ptr{Cvoid}[plan_handle(t)[1] for t in ts]
    H = handles(first(ts).Pi.topology)
    h = Ref{Ptr{Cvoid}}()
    check(ccall((:pa_host_chain_create, libpa), Cint, (Cint, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
                length(plans), plans, H.comm, h))
    finalizer(c -> ccall((:pa_host_chain_destroy, libpa), Cvoid, (Ptr{Cvoid},), c.h), HostChain(h[]))
end
function submit!(c::HostChain, host_dst::Array, host_src::Array)   # pin both with CUDA.pin
    k = Ref{Int64}(0)
    check(ccall((:pa_host_chain_submit, libpa), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}),
                c.h, pointer(host_src), pointer(host_dst), k))
    k[]
end
Base.wait(c::HostChain, ticket = -1) =
    check(ccall((:pa_host_chain_wait, libpa), Cint, (Ptr{Cvoid}, Int64), c.h, ticket))

# ---- PencilIO: MPIIODriver files straight from / into device arrays ------------------
# The reference's set_view! + MPI.File.write_all (mpi_io.jl:338-380) needs host memory;
# these methods keep its file format (and its own add_metadata / JSON sidecar) and move
# the rank's sub-box with pa_io_write / pa_io_read.
using PencilArrays.PencilIO: PencilIO
function PencilIO.write_discontiguous(ff::MPI.FileHandle, x::PencilArray{T,N,<:CuArray};
                                      offset, collective, filename, infokws...) where {T,N}
    H = handles(topology(pencil(x)))
    ex = Int64[extra_dims(x)...]
    check(ccall((:pa_io_write, libpa), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Int64}, Cint, Cint, Ptr{Cvoid}, Cstring, Int64),
        pencil_handle(H, pencil(x)), length(ex), ex, sizeof(T), 0, devptr(parent(x)), filename, offset))
    nothing
end
function PencilIO.read_discontiguous!(ff::MPI.FileHandle, x::PencilArray{T,N,<:CuArray};
                                      offset, collective, filename, infokws...) where {T,N}
    H = handles(topology(pencil(x)))
    ex = Int64[extra_dims(x)...]
    check(ccall((:pa_io_read, libpa), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Int64}, Cint, Cint, Ptr{Cvoid}, Cstring, Int64),
        pencil_handle(H, pencil(x)), length(ex), ex, sizeof(T), 0, devptr(parent(x)), filename, offset))
    x
end
# (the chunks = true pair, write_contiguous / read_contiguous!, passes 1 instead of 0; the
#  caller supplies `filename = get_filename(file)` -- two one-line forwarding methods of
#  setindex!(::MPIFile, …) / _read_mpiio! in the extension)

# MPI.Waitall(t) (Transpositions.jl:127-130)
function MPI.Waitall(t::DeviceTransposition)
    plan, _ = plan_handle(t)
    check(ccall((:pa_wait, libpa), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), plan, stream_ptr()))
    nothing
end

end # module
