/*
 * pa_b200.h -- C ABI of the B200-native global-transposition hot path.
 *
 * This is the drop-in boundary for the `transpose!` path of PencilArrays.jl
 * (reference: src/Transpositions/Transpositions.jl).  The reference has no
 * FFI of its own: its seam is Julia multiple dispatch on the array type
 * carried by a `Pencil` (Pencils.jl:282-304).  A device array type plugs in
 * by providing the "generic" methods `copy_range!` (Transpositions.jl:568-583),
 * `_viewreshape` (:615-623), `_permutedims!` (:648-664) and the transport
 * pair `transpose_send_other!` / `MPI.Alltoallv!` / `MPI.Waitany`
 * (:462-484, :422, :513).  Each entry point below names the reference
 * function it replaces.  `INTEGRATION.md` shows the Julia `ccall` stubs.
 *
 * Conventions
 *  - plain C: opaque handles, `int64_t` sizes, `void*` device pointers,
 *    `void*` for `cudaStream_t`; no torch / C++ types cross this boundary;
 *  - dimension indices, permutations, ranges and peer indices are 1-BASED and
 *    ranges are inclusive, exactly as the Julia reference passes them;
 *  - every function returns a `pa_status`; nothing throws across the ABI.
 *    The Julia veneer maps PA_EINVAL / PA_EINCOMPAT to `ArgumentError` and
 *    PA_EDIM to `DimensionMismatch` (Transpositions.jl:99-103,181-198;
 *    arrays.jl:108-114);
 *  - there is NO CPU fallback: every data-path call fails with PA_ENOGPU /
 *    PA_ECUDA when no device is usable.
 */
#ifndef PA_B200_H
#define PA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_MAX_DIMS 8        /* spatial + extra dims of one array           */
#define PA_MAX_TOPO 7        /* dims of the process grid (M < N)            */
#define PA_UNIQUE_ID_BYTES 128

typedef enum pa_status {
  PA_OK = 0,
  PA_EINVAL = 1,     /* bad argument                      -> ArgumentError      */
  PA_EINCOMPAT = 2,  /* pencils not transposable          -> ArgumentError      */
  PA_EDIM = 3,       /* array does not match its pencil   -> DimensionMismatch  */
  PA_ECUDA = 4,      /* CUDA runtime error (see pa_last_error)                  */
  PA_ENCCL = 5,      /* NCCL error / NCCL library not loadable                  */
  PA_ENOMEM = 6,     /* device or host allocation failed                        */
  PA_ESTATE = 7,     /* call sequence error (e.g. comm missing for Nproc > 1)   */
  PA_ENOGPU = 8      /* no CUDA device: the data path has no CPU fallback       */
} pa_status;

typedef enum pa_method {
  PA_POINT_TO_POINT = 0, /* Transpositions.PointToPoint  (Transpositions.jl:18) */
  PA_ALLTOALLV = 1,      /* Transpositions.Alltoallv     (Transpositions.jl:19) */
  PA_PEER_PUT = 2,       /* B200 extension (no reference counterpart): one-sided puts.
                            The pack kernel of each remote block stores straight into
                            the destination rank's `dest` array over NVLink (peer-mapped
                            memory): no send_buf, no recv_buf, no unpack pass.  Needs
                            pa_plan_set_window for `dest`; falls back to PointToPoint
                            when src and dest alias.                                   */
  PA_PEER_GET = 3        /* same, pull flavour: the unpack kernel of each remote block
                            LOADS straight out of the source rank's `src` array over
                            NVLink and stores permuted into the local `dest`.  Needs
                            pa_plan_set_window for `src`.                              */
} pa_method;

/* flags of pa_transpose */
#define PA_WAITALL   1u  /* transpose!(t; waitall=true)   (Transpositions.jl:170-176) */
#define PA_NO_OVERLAP 2u /* run pack -> exchange -> unpack strictly in sequence      */
#define PA_STAGE_SELF 4u /* self block through recv_buf like the reference (:393-403)
                            instead of the fused src->dest kernel                    */
/* Fused neighbour step: the unpack of ALL blocks and a 1-d complex FFT
 * along the destination's contiguous dimension (the one that has just become local:
 * the reason the pencil is permuted, docs/src/Pencils.md:210-214) run as ONE kernel --
 * `dst` receives fft(transposed array) (forward: exp(-2*pi*i*jk/n); backward: the
 * unnormalised inverse, as FFTW / PencilFFTs).  ComplexF64, power-of-two lines of
 * 8..1024 points, staged methods or local transposes, src and dst not aliased;
 * PA_EINVAL otherwise (transpose, then transform, separately).  A plan between
 * IDENTICAL pencils with src == dst is the plain in-place transform along the
 * contiguous dim -- the first step of a PencilFFTs-style 3-d transform, whose other
 * two steps are the fused transposes.                                               */
#define PA_FFT_FORWARD  8u
#define PA_FFT_BACKWARD 16u

typedef struct pa_topology pa_topology; /* MPITopology   (MPITopologies.jl:72-119) */
typedef struct pa_pencil pa_pencil;     /* Pencil        (Pencils.jl:151-272)      */
typedef struct pa_plan pa_plan;         /* Transposition (Transpositions.jl:69-119)*/
typedef struct pa_comm pa_comm;         /* communicator standing in for
                                           topology.comm / subcomms               */
typedef struct pa_host_chain pa_host_chain; /* asynchronous upload -> transpose!... ->
                                           download pipeline for host arrays      */

/* ---- library ------------------------------------------------------------- */
const char* pa_version(void);
const char* pa_strerror(pa_status s);
/* detail text of the last failure on the calling thread ("" if none) */
const char* pa_last_error(void);
/* number of CUDA kernels this library has launched in this process */
int64_t pa_launch_count(void);
/* number of visible CUDA devices (0 on a CPU-only box); never fails */
int pa_device_count(void);
/* run-time tunables: "remote_ctas" = grid cap of the PeerPut/PeerGet kernels
 * (they are NVLink-bound; a capped, tile-striding grid leaves SMs to the local
 * kernels running beside them; n > 0 = n CTAs, n < 0 = |n| CTAs per SM (default
 * -4), 0 = uncapped), "box_copy_ctas" = same cap for pa_box_copy (benchmarks),
 * "nccl_fences" = 1: one-sided methods fence with NCCL groups, "nccl_register"
 * (default 1; 0 = plain cudaMalloc arenas): staging arenas from ncclMemAlloc +
 * ncclCommRegister, "bulk_rows" = 1: row copies as the TMA bulk-copy pipeline,
 * "transpose_tbq" (0/16 = 256-byte destination runs per tile, the default; 32 = 512),
 * "transpose_y_fastest" (-1 = consecutive CTAs walk along the side with fewer tiles,
 * the default; 0 / 1 = forced along source / destination rows), "small_block_bytes"
 * (accepted, unused since round 2),
 * "multi_put" (default 1): the one-sided methods issue ONE launch over all peers'
 * blocks with the window protocol inside the kernel, "p2p_chunks" (default 1):
 * sub-blocks per peer block flowing pack -> exchange -> unpack independently
 * (must be equal on all ranks), "staged_ctas": grid cap of pack/unpack kernels
 * while an exchange is in flight, "ipc_exchange" = 1: the staged methods move the
 * blocks with this library's own NVLink copy kernels instead of ncclSend/ncclRecv,
 * "fence_timeout_ms" (default 60000): a flag wait longer than this records the
 * failure and traps, "pdl" (default 1): programmatic dependent launch between
 * back-to-back local kernels, "nccl_ctas": ncclCommInitRankConfig min/maxCTAs,
 * "host_chunk_bytes": bytes per pipelined cut of the host paths, "host_slots" (2..4):
 * device staging sets = submits a pa_host_chain keeps in flight, "self_first" = 1:
 * staged schedules run the self block first and beside the packs (round-1 order;
 * default: remote packs first and alone), "oneside_self_ctas": grid cap of the self
 * block while the one-sided puts run beside it (default 0 = uncapped, measured best),
 * "fft_lines" = 4: the fused unpack+FFT uses four lines per CTA for 256/512-point
 * lines too (default eight, measured faster).                                     */
pa_status pa_set_tunable(const char* name, int64_t value);
/* bind the calling thread to a device (one process per GPU: LOCAL_RANK).  All
 * handles created afterwards (streams, staging arenas, communicator) live there. */
pa_status pa_set_device(int device);

/* ---- MPITopology ---------------------------------------------------------
 * Row-major rank grid, identical to MPI_Cart_create(reorder=false)
 * (MPITopologies.jl:125-131,208-226): the LAST coordinate varies fastest.   */
pa_status pa_dims_create(int nprocs, int M, int64_t* dims /* out[M] */);
        /* balanced, non-increasing factorisation: MPI_Dims_create analogue
           (MPITopologies.jl:138-144)                                        */
pa_status pa_topology_create(int M, const int64_t* dims, int world_rank,
                             pa_topology** out);
void pa_topology_destroy(pa_topology* t);
pa_status pa_topology_info(const pa_topology* t, int* M, int64_t* dims /*[M]*/,
                           int* world_rank, int* world_size,
                           int64_t* coords_local /*[M], 1-based*/);
pa_status pa_topology_rank_of(const pa_topology* t,
                              const int64_t* coords /*1-based*/, int* rank);
/* world ranks of the grid line through the local coords along grid dim R
 * (1-based) -- `subcomm_ranks[R]` + `get_remote_indices`
 * (Transpositions.jl:293-299,539-549)                                       */
pa_status pa_topology_line(const pa_topology* t, int R, int* ranks /*[dims[R]]*/);

/* ---- Pencil -------------------------------------------------------------- */
/* `share_with` != NULL reproduces Pencil(p; decomp_dims, permute): the new
 * pencil shares send_buf / recv_buf with `share_with` (Pencils.jl:257-271). */
pa_status pa_pencil_create(pa_topology* topo, int N, const int64_t* size_global,
                           const int* decomp_dims /*[M], 1-based*/,
                           const int* perm /*[N], 1-based; NULL = NoPermutation*/,
                           pa_pencil* share_with, pa_pencil** out);
void pa_pencil_destroy(pa_pencil* p);
/* range_local / range_remote (Pencils.jl:472-497): `coords` NULL = local rank;
 * memory_order != 0 applies the pencil's permutation. lo/hi 1-based inclusive
 * (hi = lo - 1 for an empty range).                                          */
pa_status pa_pencil_range(const pa_pencil* p, const int64_t* coords,
                          int memory_order, int64_t* lo, int64_t* hi);
pa_status pa_pencil_size_local(const pa_pencil* p, int memory_order,
                               int64_t* dims /*[N]*/);
/* staging arenas living in the pencil family (Pencils.jl:185-189).  Device
 * pointers; capacity in bytes; grow-only (Transpositions.jl:313-317).       */
pa_status pa_pencil_buffers(const pa_pencil* p, void** send_buf,
                            int64_t* send_cap, void** recv_buf,
                            int64_t* recv_cap);
pa_status pa_pencil_reserve(pa_pencil* p, int64_t send_bytes, int64_t recv_bytes);

/* ---- Transposition plan --------------------------------------------------
 * Transposition(dest, src; method) (Transpositions.jl:93-118): compatibility
 * checks (assert_compatible, :181-198), `dim` discovery (:110), per-peer
 * ranges/offsets (transpose_send!, :380-416; transpose_recv!, :516-524), all
 * reduced once to kernel launch descriptors.                                */
pa_status pa_plan_create(pa_pencil* pin, pa_pencil* pout, int n_extra,
                         const int64_t* extra_dims, int elsize, pa_method method,
                         pa_plan** out);
void pa_plan_destroy(pa_plan* plan);

typedef struct pa_plan_info {
  int dim;            /* grid dim of the exchange, 1-based; 0 = `nothing` (local) */
  int nproc;          /* topology.dims[dim]  (1 when dim == 0)                  */
  int self_index;     /* my index in the grid line, 1-based                      */
  int same_perm;      /* local path: plain copy! (Transpositions.jl:226-227)     */
  int elsize;
  int method;
  int64_t length_in;       /* length(Ai), elements incl. extra dims             */
  int64_t length_out;      /* length(Ao)                                        */
  int64_t length_self;     /* Transpositions.jl:302-305                         */
  int64_t send_bytes;      /* sizeof(T) * length_send       (:308,313)          */
  int64_t recv_bytes;      /* sizeof(T) * length_recv_total (:309,316)          */
} pa_plan_info;
pa_status pa_plan_get_info(const pa_plan* plan, pa_plan_info* info);

typedef struct pa_peer_info {
  int world_rank;        /* rank of peer n in the world communicator            */
  int is_self;
  int64_t send_offset;   /* bytes into send_buf (self: unused, 0)               */
  int64_t send_count;    /* bytes                                               */
  int64_t recv_offset;   /* bytes into recv_buf; self block sits at the tail
                            (Transpositions.jl:393-403)                         */
  int64_t recv_count;    /* bytes                                               */
  int64_t remote_recv_offset; /* bytes into the PEER's recv_buf where this rank's block
                            lands (= that peer's recv_offset for me): the address the
                            own-kernel exchange stores to (pa_plan_set_recv_window)  */
} pa_peer_info;
pa_status pa_plan_get_peer(const pa_plan* plan, int n /*1-based*/, pa_peer_info* info);

/* Strided-copy descriptor of one block as the kernels see it, exported so
 * tests can compare the C++ plan with the oracle's independent derivation.
 * op: 0 = pack (src parent -> contiguous), 1 = unpack (contiguous -> dest
 * parent), 2 = fused self/local (src parent -> dest parent), 3 = put (src
 * parent -> peer n's dest parent, in the peer's layout; self: empty), 4 = get
 * (peer n's src parent, in its layout -> dest parent).
 * Dims are listed in the source's memory order incl. merged extra dims;
 * strides and offsets in elements.                                           */
typedef struct pa_block_desc {
  int nd;
  int64_t extent[PA_MAX_DIMS];
  int64_t src_stride[PA_MAX_DIMS];
  int64_t dst_stride[PA_MAX_DIMS];
  int64_t src_offset;
  int64_t dst_offset;
  int kernel_class;   /* 0 empty, 1 row copy, 2 tiled transpose, 3 generic scalar,
                         4 contiguous memcpy                                     */
  int vec_bytes;      /* access width chosen for this block                      */
} pa_block_desc;
pa_status pa_plan_get_block(const pa_plan* plan, int op, int n /*1-based*/,
                            pa_block_desc* desc);
/* sub-block `part` (0-based) of `nparts` of a pack (op 0) / unpack (op 1) block as the
 * chunked PointToPoint schedule cuts it (tunable "p2p_chunks"): a sub-range of the
 * block's outermost dimension, i.e. one contiguous piece of the wire block --
 * [*wire_offset, *wire_offset + *wire_bytes) bytes into send_buf / recv_buf.     */
pa_status pa_plan_get_chunk(const pa_plan* plan, int op, int n, int part, int nparts,
                            pa_block_desc* desc, int64_t* wire_offset, int64_t* wire_bytes);

/* ---- kernels (enqueue on `stream`; asynchronous) -------------------------
 * K1 pack: copy_range! (Transpositions.jl:552-583).  `buf` is the base of
 * send_buf (remote peer) or recv_buf (self) -- offsets come from the plan.  */
pa_status pa_pack(pa_plan* plan, int n, const void* src, void* buf, void* stream);
/* K2 unpack: copy_permuted! -> _viewreshape -> _permutedims! (:585-664)     */
pa_status pa_unpack(pa_plan* plan, int n, const void* recv_buf, void* dst, void* stream);
/* K3 fused self block: src parent -> permuted dest parent in one pass
 * (replaces :393-403 followed by :527-529 for the local block)              */
pa_status pa_copy_self(pa_plan* plan, const void* src, void* dst, void* stream);
/* K1-put: block for peer n read from `src` and stored, already permuted, into
 * `peer_dst` = peer n's dest parent array (any pointer this device can write:
 * an IPC/peer mapping, a symmetric heap, or -- for tests -- local memory).
 * Replaces :406-412 + the peer's :527-529 for that block in one pass.         */
pa_status pa_put(pa_plan* plan, int n, const void* src, void* peer_dst, void* stream);
/* K2-get: the block peer n holds for this rank, loaded from `peer_src` = peer
 * n's src parent array (its layout) and stored permuted into `dst`.           */
pa_status pa_get(pa_plan* plan, int n, const void* peer_src, void* dst, void* stream);
/* K1-put / K2-get of EVERY remote block in ONE launch (the kernel the one-sided
 * methods run, here without the window protocol): `peers[n-1]` = peer n's dest
 * (put) / src (get) parent array, the self entry is ignored.  Tiles of the blocks
 * are interleaved round-robin so that all destination links are driven at once.
 * `max_ctas`: grid cap (0 = tunable "remote_ctas").  Falls back to one launch per
 * block when the blocks need different kernel flavours.                        */
pa_status pa_put_all(pa_plan* plan, const void* src, void* const* peers, int max_ctas,
                     void* stream);
pa_status pa_get_all(pa_plan* plan, void* const* peers, void* dst, int max_ctas, void* stream);
/* transpose_impl!(::Nothing) / permute_local! (:213-270).  `scratch` must hold
 * length_out elements when src and dst alias, may be NULL otherwise.        */
pa_status pa_permute_local(pa_plan* plan, const void* src, void* dst,
                           void* scratch, void* stream);
/* raw N-d strided copy used by all of the above (bench / tests)             */
pa_status pa_box_copy(int nd, const int64_t* extent, const int64_t* src_stride,
                      const int64_t* dst_stride, int elsize, const void* src,
                      void* dst, void* stream, pa_block_desc* chosen /*NULL ok*/);

/* ---- communicator --------------------------------------------------------
 * Stands in for MPI.COMM_WORLD + MPI_Cart_sub sub-communicators
 * (MPITopologies.jl:244-251): one NCCL communicator over all ranks, peers of
 * a grid line addressed by world rank.  Bootstrap: rank 0 calls
 * pa_comm_unique_id and distributes the 128 bytes by any side channel.      */
pa_status pa_comm_unique_id(void* id128);
pa_status pa_comm_init_rank(const void* id128, int nranks, int rank, pa_comm** out);
/* NCCL-free communicator: ranks talk only through peer-mapped memory (CUDA IPC) and
 * the flag window below, which is then mandatory.  Serves PA_PEER_PUT / PA_PEER_GET
 * as they are and the staged methods through the library's own NVLink copy kernels
 * (pa_plan_set_recv_window).  Unlike NCCL it also allows several ranks on ONE device
 * (test rigs; a single-GPU box can exercise the whole multi-rank path).          */
pa_status pa_comm_init_local(int nranks, int rank, pa_comm** out);
void pa_comm_destroy(pa_comm* c);
/* Flag window: each rank exports a small device buffer (a few 64-bit words per
 * source rank) and imports everybody else's.  Words only increase (signals are
 * red.max.release.sys over NVLink, waits are ld.acquire.sys polls with the
 * "fence_timeout_ms" time-out): window-open / window-close of PA_PEER_PUT /
 * PA_PEER_GET ride inside the transfer kernel, block arrival of the own-kernel
 * exchange is one word per peer.  Without it (NCCL communicators only) the
 * one-sided methods fence with NCCL send/recv groups.                          */
pa_status pa_comm_flags_export(pa_comm* c, void* handle64, int64_t* offset);
pa_status pa_comm_flags_import(pa_comm* c, int rank, const void* handle64, int64_t offset);

/* ---- PeerPut windows (PA_PEER_PUT) -----------------------------------------
 * The analogue of a collective MPI_Win_create over `dest`: every rank exports
 * an IPC handle of its `dest` array, hands it to the peers of its grid line by
 * any side channel (the host mirror uses the same channel as the NCCL id), and
 * registers the mapped peer pointers with the plan.                          */
#define PA_IPC_HANDLE_BYTES 64
/* handle of the device allocation containing `devptr` + byte offset of devptr in it */
pa_status pa_ipc_export(const void* devptr, void* handle64, int64_t* offset);
/* map a peer's allocation (one mapping per handle, reference-counted) and return
 * base + offset; every successful import is paired with one pa_ipc_release     */
pa_status pa_ipc_import(const void* handle64, int64_t offset, void** mapped);
/* drop one reference; the mapping is closed (cudaIpcCloseMemHandle) with the last */
pa_status pa_ipc_release(const void* handle64);
/* `peer_dst` = peer n's (1-based index in the grid line) dest array as mapped here;
 * `local_dst` = this rank's dest array the window belongs to                   */
pa_status pa_plan_set_window(pa_plan* plan, const void* local_dst, int n, void* peer_dst);
/* own-kernel exchange of the staged methods (NCCL-free communicator, or tunable
 * "ipc_exchange"): `peer_recv_buf` = peer n's recv_buf arena (pa_pencil_buffers
 * after pa_pencil_reserve, exported with pa_ipc_export) as mapped here.  Must be
 * renewed when the arenas are reallocated (PA_ESTATE otherwise).                */
pa_status pa_plan_set_recv_window(pa_plan* plan, int n, void* peer_recv_buf);

/* ---- transpose! ----------------------------------------------------------
 * transpose!(t; waitall) (Transpositions.jl:170-179) for device arrays.
 * Ordered after prior work on `stream`; on return, later work on `stream`
 * sees `dst` complete.  Without PA_WAITALL the send side (send_buf reuse) is
 * only guaranteed after pa_wait -- MPI.Waitall(t) (:127-130).
 * `comm` may be NULL when nproc == 1.                                        */
pa_status pa_transpose(pa_plan* plan, pa_comm* comm, const void* src, void* dst,
                       unsigned flags, void* stream);
pa_status pa_wait(pa_plan* plan, void* stream);
/* same, with HOST arrays: H2D of `src`, transpose!, D2H of `dst`; blocks until
 * `host_dst` is valid.  Device staging is owned by the plan.  A purely local
 * transposition (nproc == 1 or dim == nothing) is cut along the source's outermost
 * dimension and pipelined on three streams: upload(c+1) || kernel(c) ||
 * download(c-1) (the download joins when that dimension is outermost in `dst` too).
 * Pinning the host arrays (cudaHostRegister / pinned allocation) is the caller's job. */
pa_status pa_transpose_host(pa_plan* plan, pa_comm* comm, const void* host_src,
                            void* host_dst, unsigned flags);

/* ---- host chains -----------------------------------------------------------
 * A sequence of transpositions applied to HOST arrays (plan i+1 consumes what
 * plan i produces): one submit = upload of `host_src`, every transpose! on the
 * device, download into `host_dst`.  Submits return immediately and are double-
 * buffered on the device: the download of one overlaps the upload of the next
 * (PCIe is full duplex).  pa_host_chain_wait(ticket) blocks until that submit's
 * `host_dst` is valid (ticket < 0: all).  The host arrays must stay untouched /
 * alive until then.  For the one-sided methods the device buffers the plans will
 * see are exposed through pa_host_chain_buffer(slot 0..1, which 0..1) so that
 * windows can be registered on them (plan i of a submit in slot s reads buffer
 * (s, i % 2) and writes buffer (s, 1 - i % 2)).                                  */
pa_status pa_host_chain_create(int n, pa_plan* const* plans, pa_comm* comm,
                               pa_host_chain** out);
void pa_host_chain_destroy(pa_host_chain* c);
pa_status pa_host_chain_submit(pa_host_chain* c, const void* host_src, void* host_dst,
                               int64_t* ticket /* out, may be NULL */);
pa_status pa_host_chain_wait(pa_host_chain* c, int64_t ticket);
pa_status pa_host_chain_buffer(pa_host_chain* c, int slot, int which, void** devptr,
                               int64_t* bytes);
/* CUDA-event bracket around a run of submits (device-side timing of the pipeline):
 * begin marks the upload stream now, end waits for everything submitted so far and
 * returns the milliseconds between the two.                                      */
pa_status pa_host_chain_time_begin(pa_host_chain* c);
pa_status pa_host_chain_time_end(pa_host_chain* c, float* ms);

/* ---- PencilIO binary layout ---------------------------------------
 * Device arrays <-> the raw binary files of the reference's MPIIODriver
 * (src/PencilIO/mpi_io.jl), byte for byte:
 *   chunks == 0  the dataset is the GLOBAL array in the pencil's memory order, dims
 *                (perm * size_global..., extra_dims...), column-major; this rank owns
 *                the sub-box range_local(x, MemoryOrder()) (:372-380) -- files can be
 *                read back with any other decomposition;
 *   chunks != 0  the ranks' parent arrays one after the other in column-major order of
 *                the process-grid coordinates (:382-424).
 * `offset` = byte offset of the dataset in the file (MPIFile position, :159-164).
 * Every rank calls these for its own part (no collective: ranks pwrite / pread
 * disjoint byte ranges of the same file); the device array moves through a double-
 * buffered pinned staging area on a stream of the library's own: the calls block until
 * the transfer is complete and are NOT ordered against work queued on the caller's
 * streams -- synchronise those first.  The JSON sidecar (:194-211) is written by the host
 * veneer (Julia: the reference's own add_metadata; Python mirror: pencilio.py).     */
pa_status pa_io_sizes(const pa_pencil* p, int n_extra, const int64_t* extra_dims, int elsize,
                      int chunks, int64_t* global_bytes, int64_t* local_bytes, int64_t* nruns,
                      int64_t* run_bytes, int64_t* first_offset);
/* the local array is `nruns` runs of `run_bytes` bytes, contiguous in the array and in the
 * file; run r of the array starts `*file_offset` bytes into the dataset               */
pa_status pa_io_run_offset(const pa_pencil* p, int n_extra, const int64_t* extra_dims, int elsize,
                           int chunks, int64_t run, int64_t* file_offset);
pa_status pa_io_write(const pa_pencil* p, int n_extra, const int64_t* extra_dims, int elsize,
                      int chunks, const void* dev_array, const char* path, int64_t offset);
pa_status pa_io_read(const pa_pencil* p, int n_extra, const int64_t* extra_dims, int elsize,
                     int chunks, void* dev_array, const char* path, int64_t offset);

/* CUDA-event timings (ms) of the last pa_transpose on this plan, named after
 * the reference's TimerOutputs sections (Transpositions.jl:172-175,326,336).
 * Blocks until that transpose has finished.                                  */
typedef struct pa_timings {
  float total_ms;        /* "transpose!"    */
  float pack_ms;         /* "pack data"     */
  float exchange_ms;     /* Isend/Irecv or "MPI.Alltoallv!" */
  float unpack_ms;       /* "unpack data"   */
} pa_timings;
pa_status pa_plan_timings(pa_plan* plan, pa_timings* t);
pa_status pa_plan_enable_timing(pa_plan* plan, int on);

#ifdef __cplusplus
}
#endif
#endif /* PA_B200_H */
