// transpose! driver: streams, events and the NCCL exchange.
//
// Reference flow (Transpositions.jl:281-343): pack every block
// (transpose_send!, :345-430) posting Isend/Irecv per peer as soon as its
// block is packed (:406-412) or one Alltoallv after all packs (:418-427); then
// unpack blocks as they arrive (transpose_recv!, :486-533), self block first.
//
// B200 restatement: three CUDA streams (pack / comm / unpack) joined by
// events.  PointToPoint = one grouped {ncclSend, ncclRecv} per exchange step,
// enqueued the moment that step's pack finishes, unpack gated on that step's
// receive -- so pack(k+1), exchange(k) and unpack(k-1) overlap.  Steps follow
// a rotation (send to me+k, receive from me-k) instead of the reference's
// identical 1..Nproc order on every rank: over NVSwitch all peers are
// equidistant and the rotation keeps every link busy at every step.
// Alltoallv = one group holding every peer's send and receive.
// NCCL is loaded with dlopen so that the library itself has no link-time
// dependency and picks up the libnccl already mapped by the host process.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <set>

#include "pa_internal.hpp"

namespace pa {

#define CU(call)                                                                  \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      set_error("%s failed: %s", #call, cudaGetErrorString(e_));                  \
      return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver)       \
                 ? PA_ENOGPU                                                      \
                 : (e_ == cudaErrorMemoryAllocation ? PA_ENOMEM : PA_ECUDA);      \
    }                                                                             \
  } while (0)

pa_status set_device(int dev) {
  CU(cudaSetDevice(dev));
  return PA_OK;
}

// ---- CUDA IPC: windows of the PeerPut method -----------------------------------
typedef int (*cuMemGetAddressRange_fn)(unsigned long long*, size_t*, unsigned long long);

static cuMemGetAddressRange_fn addr_range_fn() {
  static cuMemGetAddressRange_fn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (cuMemGetAddressRange_fn)f;
  }();
  return fn;
}

pa_status ipc_export(const void* devptr, void* handle64, i64* offset) {
  static_assert(sizeof(cudaIpcMemHandle_t) == PA_IPC_HANDLE_BYTES, "ipc handle size");
  if (device_count() == 0) {
    set_error("no CUDA device");
    return PA_ENOGPU;
  }
  CU(cudaFree(nullptr));  // make sure this runtime instance has its context
  cuMemGetAddressRange_fn fn = addr_range_fn();
  unsigned long long base = 0;
  size_t size = 0;
  if (!fn || fn(&base, &size, (unsigned long long)(uintptr_t)devptr) != 0) {
    set_error("cuMemGetAddressRange failed for %p (not a device allocation?)", devptr);
    return PA_ECUDA;
  }
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
  memcpy(handle64, &h, sizeof h);
  *offset = (i64)((unsigned long long)(uintptr_t)devptr - base);
  return PA_OK;
}

pa_status ipc_import(const void* handle64, i64 offset, void** mapped) {
  static std::mutex mu;
  static std::map<std::string, void*> cache;  // one mapping per peer allocation, kept for the process
  std::lock_guard<std::mutex> lock(mu);
  std::string key((const char*)handle64, PA_IPC_HANDLE_BYTES);
  auto it = cache.find(key);
  void* base = nullptr;
  if (it != cache.end()) {
    base = it->second;
  } else {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    CU(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    cache[key] = base;
  }
  *mapped = (char*)base + offset;
  return PA_OK;
}

pa_status plan_set_window(Plan* P, const void* local_dst, int n0, void* peer_dst) {
  if (P->dim < 0 || n0 < 0 || n0 >= P->nproc) {
    set_error("window peer index out of range");
    return PA_EINVAL;
  }
  std::vector<void*>& v = P->windows[local_dst];
  v.resize(P->nproc, nullptr);
  v[n0] = peer_dst;
  return PA_OK;
}

// ---- NCCL via dlopen ---------------------------------------------------------
struct NcclApi {
  void* h = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  decltype(&ncclMemAlloc) MemAlloc = nullptr;        // optional (user-buffer registration)
  decltype(&ncclMemFree) MemFree = nullptr;
  decltype(&ncclCommRegister) CommRegister = nullptr;
  decltype(&ncclCommDeregister) CommDeregister = nullptr;
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (!api.h) return;
#define LOAD(sym) api.sym = (decltype(api.sym))dlsym(api.h, "nccl" #sym)
    LOAD(GetUniqueId);
    LOAD(CommInitRank);
    LOAD(CommDestroy);
    LOAD(Send);
    LOAD(Recv);
    LOAD(GroupStart);
    LOAD(GroupEnd);
    LOAD(GetErrorString);
    LOAD(GetVersion);
    LOAD(MemAlloc);
    LOAD(MemFree);
    LOAD(CommRegister);
    LOAD(CommDeregister);
#undef LOAD
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv &&
             api.GroupStart && api.GroupEnd && api.GetErrorString;
  });
  return api;
}

#define NC(call)                                                       \
  do {                                                                 \
    ncclResult_t r_ = (call);                                          \
    if (r_ != ncclSuccess) {                                           \
      set_error("%s failed: %s", #call, nccl().GetErrorString(r_));    \
      return PA_ENCCL;                                                 \
    }                                                                  \
  } while (0)

struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 0, rank = 0, device = 0;
  // Flag window of the one-sided paths: one 64-bit sequence word per source
  // rank, written by that rank over NVLink (st.release.sys) and polled locally.
  unsigned long long* flags = nullptr;                // my words, indexed by source rank
  std::vector<unsigned long long*> peer_flags;        // peers' windows as mapped here
  std::vector<unsigned long long> seq_with;           // fences issued with each rank so far
  int* fence_err = nullptr;                           // mapped pinned host word set on time-out
  bool flags_ready = false;
};

// ---- NVLink fence: signal + wait among the ranks of one grid line -------------
constexpr int FENCE_MAX = 64;
struct FenceParams {
  int n;
  unsigned long long* remote[FENCE_MAX];  // peer's word for me
  unsigned long long* local[FENCE_MAX];   // my word for that peer
  unsigned long long seq[FENCE_MAX];
  unsigned long long timeout_ns;
  int* err;
};

__global__ void k_fence(const __grid_constant__ FenceParams fp) {
  const int t = threadIdx.x;
  if (t >= fp.n) return;
  // everything this GPU wrote before (puts into peer memory included) is visible
  // system-wide before the peer can observe the new sequence number
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fp.remote[t]), "l"(fp.seq[t]) : "memory");
  unsigned long long t0, now, v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(fp.local[t]) : "memory");
    if (v >= fp.seq[t]) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (now - t0 > fp.timeout_ns) {  // a peer died: do not hang the GPU
      *fp.err = 1;
      break;
    }
    __nanosleep(200);
  }
}

pa_status comm_unique_id(void* id128) {
  static_assert(sizeof(ncclUniqueId) <= PA_UNIQUE_ID_BYTES, "unique id size");
  if (!nccl().ok) {
    set_error("libnccl.so.2 could not be loaded");
    return PA_ENCCL;
  }
  ncclUniqueId id;
  NC(nccl().GetUniqueId(&id));
  memset(id128, 0, PA_UNIQUE_ID_BYTES);
  memcpy(id128, &id, sizeof id);
  return PA_OK;
}

static std::mutex g_live_mu;
static std::set<void*> g_live_comms;  // communicators that may still hold registrations

pa_status comm_init(const void* id128, int nranks, int rank, Comm** out) {
  if (!nccl().ok) {
    set_error("libnccl.so.2 could not be loaded");
    return PA_ENCCL;
  }
  if (device_count() == 0) {
    set_error("no CUDA device");
    return PA_ENOGPU;
  }
  std::unique_ptr<Comm> c(new Comm);
  c->nranks = nranks;
  c->rank = rank;
  CU(cudaGetDevice(&c->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NC(nccl().CommInitRank(&c->comm, nranks, id, rank));
  {
    std::lock_guard<std::mutex> lock(g_live_mu);
    g_live_comms.insert((void*)c->comm);
  }
  *out = c.release();
  return PA_OK;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  {
    std::lock_guard<std::mutex> lock(g_live_mu);
    g_live_comms.erase((void*)c->comm);
  }
  if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
  if (c->flags) cudaFree(c->flags);
  if (c->fence_err) cudaFreeHost(c->fence_err);
  delete c;
}

// flag window: allocate + export (collective exchange is the caller's job)
pa_status comm_flags_export(Comm* c, void* handle64, i64* offset) {
  if (!c->flags) {
    CU(cudaMalloc((void**)&c->flags, sizeof(unsigned long long) * (size_t)c->nranks));
    CU(cudaMemset(c->flags, 0, sizeof(unsigned long long) * (size_t)c->nranks));
    CU(cudaHostAlloc((void**)&c->fence_err, sizeof(int), cudaHostAllocMapped));
    *c->fence_err = 0;
    CU(cudaDeviceSynchronize());
    c->peer_flags.assign(c->nranks, nullptr);
    c->seq_with.assign(c->nranks, 0);
  }
  return ipc_export(c->flags, handle64, offset);
}

pa_status comm_flags_import(Comm* c, int rank, const void* handle64, i64 offset) {
  if (!c->flags || rank < 0 || rank >= c->nranks) {
    set_error("flag window: export first / rank out of range");
    return PA_ESTATE;
  }
  if (rank == c->rank) return PA_OK;
  void* p = nullptr;
  pa_status s = ipc_import(handle64, offset, &p);
  if (s != PA_OK) return s;
  c->peer_flags[rank] = (unsigned long long*)p;
  bool all = true;
  for (int r = 0; r < c->nranks; ++r)
    if (r != c->rank && !c->peer_flags[r]) all = false;
  c->flags_ready = all;
  return PA_OK;
}

// ---- staging arenas ----------------------------------------------------------
static void buffers_deregister(Buffers& b) {
  if (!b.reg_comm) return;
  bool live;
  {
    std::lock_guard<std::mutex> lock(g_live_mu);
    live = g_live_comms.count(b.reg_comm) != 0;
  }
  if (live && nccl().CommDeregister) {
    if (b.reg_send) nccl().CommDeregister((ncclComm_t)b.reg_comm, b.reg_send);
    if (b.reg_recv) nccl().CommDeregister((ncclComm_t)b.reg_comm, b.reg_recv);
  }
  b.reg_comm = b.reg_send = b.reg_recv = b.reg_send_ptr = b.reg_recv_ptr = nullptr;
}

static void free_arena(void* p, bool from_nccl) {
  if (!p) return;
  if (from_nccl && nccl().MemFree) nccl().MemFree(p);
  else cudaFree(p);
}

Buffers::~Buffers() {
  buffers_deregister(*this);
  free_arena(send, send_nccl);
  free_arena(recv, recv_nccl);
  if (comm_done_event) cudaEventDestroy((cudaEvent_t)comm_done_event);
}

// grow-only, like resize! on the pencil's UInt8 vectors (Transpositions.jl:313-317)
pa_status Buffers::reserve(i64 send_bytes, i64 recv_bytes) {
  auto grow = [this](void*& p, i64& cap, bool& from_nccl, i64 need) -> pa_status {
    if (need <= cap) return PA_OK;
    // a previous exchange may still be reading/writing the old arena
    CU(cudaDeviceSynchronize());
    buffers_deregister(*this);
    free_arena(p, from_nccl);
    p = nullptr;
    cap = 0;
    i64 n = (need + 255) / 256 * 256;
    from_nccl = false;
    if (g_tun.nccl_register && nccl().ok && nccl().MemAlloc && nccl().MemFree &&
        nccl().MemAlloc(&p, (size_t)n) == ncclSuccess && p) {
      from_nccl = true;
    } else {
      p = nullptr;
      CU(cudaMalloc(&p, (size_t)n));
    }
    cap = n;
    return PA_OK;
  };
  pa_status s = grow(send, send_cap, send_nccl, send_bytes);
  if (s != PA_OK) return s;
  return grow(recv, recv_cap, recv_nccl, recv_bytes);
}

// register the arenas with `comm` (no-op unless tunable nccl_register and ncclMemAlloc'ed arenas)
static void buffers_register(Buffers& b, ncclComm_t comm) {
  if (!g_tun.nccl_register || !nccl().CommRegister) return;
  if (b.reg_comm == (void*)comm && b.reg_send_ptr == b.send && b.reg_recv_ptr == b.recv) return;
  buffers_deregister(b);
  b.reg_comm = (void*)comm;
  if (b.send && b.send_nccl &&
      nccl().CommRegister(comm, b.send, (size_t)b.send_cap, &b.reg_send) != ncclSuccess)
    b.reg_send = nullptr;
  if (b.recv && b.recv_nccl &&
      nccl().CommRegister(comm, b.recv, (size_t)b.recv_cap, &b.reg_recv) != ncclSuccess)
    b.reg_recv = nullptr;
  b.reg_send_ptr = b.send;
  b.reg_recv_ptr = b.recv;
}

// ---- per-plan stream/event state --------------------------------------------
struct TransposeState {
  cudaStream_t pack_s = nullptr, comm_s = nullptr, unpack_s = nullptr, host_s = nullptr;
  cudaEvent_t ev_start = nullptr, ev_allpacked = nullptr, ev_comm_done = nullptr,
              ev_unpack_done = nullptr;
  std::vector<cudaEvent_t> ev_packed, ev_recvd;
  bool timing = false;
  bool timed_once = false;
  cudaEvent_t t[8] = {nullptr};  // 0 start,1 pack_end,2 comm0,3 comm1,4 unpack0,5 unpack1,6 end
  bool sends_pending = false;
  char* tok = nullptr;  // 4-byte tokens of the PeerPut line barrier: [0] sent, [1+n] received from n
};

void destroy_state(TransposeState* st) {
  if (!st) return;
  if (st->pack_s) cudaStreamDestroy(st->pack_s);
  if (st->comm_s) cudaStreamDestroy(st->comm_s);
  if (st->unpack_s) cudaStreamDestroy(st->unpack_s);
  if (st->host_s) cudaStreamDestroy(st->host_s);
  for (cudaEvent_t e : {st->ev_start, st->ev_allpacked, st->ev_comm_done, st->ev_unpack_done})
    if (e) cudaEventDestroy(e);
  for (auto e : st->ev_packed) cudaEventDestroy(e);
  for (auto e : st->ev_recvd) cudaEventDestroy(e);
  for (auto e : st->t)
    if (e) cudaEventDestroy(e);
  if (st->tok) cudaFree(st->tok);
  delete st;
}

Plan::~Plan() {
  destroy_state(st);
  if (h_src_dev) cudaFree(h_src_dev);
  if (h_dst_dev) cudaFree(h_dst_dev);
}

static pa_status ensure_state(Plan* P) {
  if (P->st) return PA_OK;
  if (device_count() == 0) {
    set_error("no CUDA device: the transpose! path has no CPU fallback");
    return PA_ENOGPU;
  }
  std::unique_ptr<TransposeState> st(new TransposeState);
  int lo = 0, hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CU(cudaStreamCreateWithPriority(&st->pack_s, cudaStreamNonBlocking, lo));
  CU(cudaStreamCreateWithPriority(&st->comm_s, cudaStreamNonBlocking, hi));  // exchange first
  CU(cudaStreamCreateWithPriority(&st->unpack_s, cudaStreamNonBlocking, lo));
  CU(cudaStreamCreateWithFlags(&st->host_s, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&st->ev_start, &st->ev_allpacked, &st->ev_comm_done, &st->ev_unpack_done})
    CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  st->ev_packed.resize(P->nproc);
  st->ev_recvd.resize(P->nproc);
  for (int i = 0; i < P->nproc; ++i) {
    CU(cudaEventCreateWithFlags(&st->ev_packed[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&st->ev_recvd[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < 7; ++i) CU(cudaEventCreate(&st->t[i]));
  CU(cudaMalloc((void**)&st->tok, 4 * (size_t)(P->nproc + 1)));
  CU(cudaMemset(st->tok, 0, 4 * (size_t)(P->nproc + 1)));
  P->st = st.release();
  return PA_OK;
}

pa_status plan_enable_timing(Plan* P, int on) {
  pa_status s = ensure_state(P);
  if (s != PA_OK) return s;
  P->st->timing = on != 0;
  return PA_OK;
}

pa_status plan_timings(Plan* P, pa_timings* out) {
  memset(out, 0, sizeof *out);
  if (!P->st || !P->st->timed_once) {
    set_error("no timed transpose! has run on this plan (pa_plan_enable_timing first)");
    return PA_ESTATE;
  }
  TransposeState& S = *P->st;
  CU(cudaEventSynchronize(S.t[6]));
  CU(cudaEventElapsedTime(&out->total_ms, S.t[0], S.t[6]));
  if (P->dim >= 0 && P->nproc > 1) {
    CU(cudaEventElapsedTime(&out->pack_ms, S.t[0], S.t[1]));
    CU(cudaEventElapsedTime(&out->exchange_ms, S.t[2], S.t[3]));
    CU(cudaEventElapsedTime(&out->unpack_ms, S.t[4], S.t[5]));
  }
  return PA_OK;
}

static bool ranges_overlap(const void* a, i64 na, const void* b, i64 nb) {
  const char* pa_ = (const char*)a;
  const char* pb = (const char*)b;
  return pa_ < pb + nb && pb < pa_ + na;
}

// transpose_impl!(::Nothing) (Transpositions.jl:213-270)
static pa_status local_transpose(Plan* P, const void* src, void* dst, void* scratch,
                                 cudaStream_t st) {
  const i64 bytes = P->length_out * P->elsize;
  if (P->same_perm) {
    if (src == dst) return PA_OK;  // copy!(uo, ui) onto itself
    return launch_block(P->self_fused, src, dst, st, nullptr);
  }
  const bool inplace = ranges_overlap(src, bytes, dst, bytes);  // Base.mightalias (:249)
  if (!inplace) return launch_block(P->self_fused, src, dst, st, nullptr);
  if (!scratch) {
    set_error("aliased local permutation needs a scratch buffer");
    return PA_EINVAL;
  }
  // permute into the temporary, then copy! to the output (:251-263)
  pa_status s = launch_block(P->self_fused, src, scratch, st, nullptr);
  if (s != PA_OK) return s;
  CU(cudaMemcpyAsync(dst, scratch, (size_t)bytes, cudaMemcpyDeviceToDevice, st));
  return PA_OK;
}

pa_status permute_local(Plan* P, const void* src, void* dst, void* scratch, void* stream) {
  if (P->dim >= 0) {
    set_error("plan needs an exchange (dim = %d); use pa_transpose", P->dim + 1);
    return PA_ESTATE;
  }
  if (device_count() == 0) {
    set_error("no CUDA device");
    return PA_ENOGPU;
  }
  return local_transpose(P, src, dst, scratch, (cudaStream_t)stream);
}

pa_status transpose(Plan* P, Comm* comm, const void* src, void* dst, unsigned flags,
                    void* stream) {
  pa_status rc = ensure_state(P);
  if (rc != PA_OK) return rc;
  TransposeState& S = *P->st;
  cudaStream_t user = (cudaStream_t)stream;
  const i64 ES = P->elsize;
  const bool timing = S.timing;
  if (timing) CU(cudaEventRecord(S.t[0], user));

  if (P->dim < 0) {
    void* scratch = nullptr;
    const i64 bytes = P->length_out * ES;
    if (!P->same_perm && ranges_overlap(src, bytes, dst, bytes)) {
      rc = P->pin->bufs->reserve(0, std::max<i64>(1, bytes));  // reuses Pi.recv_buf (:255)
      if (rc != PA_OK) return rc;
      scratch = P->pin->bufs->recv;
    }
    rc = local_transpose(P, src, dst, scratch, user);
    if (rc != PA_OK) return rc;
    if (timing) {
      CU(cudaEventRecord(S.t[6], user));
      S.timed_once = true;
    }
    return PA_OK;
  }

  const int nproc = P->nproc, me = P->self_index;
  if (nproc > 1 && !comm) {
    set_error("this transposition exchanges data among %d ranks: a communicator is required",
              nproc);
    return PA_ESTATE;
  }
  Buffers& B = *P->pout->bufs;  // Po.send_buf / Po.recv_buf (:313-317)
  const bool aliased = ranges_overlap(src, P->length_in * ES, dst, P->length_out * ES);
  const bool stage_self = aliased || (flags & PA_STAGE_SELF);
  const bool overlap = !(flags & PA_NO_OVERLAP);
  if (!((P->method == PA_PEER_PUT || P->method == PA_PEER_GET) && !stage_self && nproc > 1)) {
    // (one-sided puts/gets need no staging arenas)
    i64 need_send = nproc > 1 ? std::max<i64>(1, P->send_elems * ES) : 0;
    i64 need_recv = (nproc > 1 || stage_self) ? std::max<i64>(1, P->recv_elems * ES) : 0;
    rc = B.reserve(need_send, need_recv);
    if (rc != PA_OK) return rc;
  }
  if (comm && nproc > 1) buffers_register(B, comm->comm);
  char* sbuf = (char*)B.send;
  char* rbuf = (char*)B.recv;
  const Peer& self = P->peers[me];

  if (nproc == 1) {
    // only the self block: no exchange, everything on the caller's stream
    if (stage_self) {
      rc = launch_block(self.pack, src, rbuf, user, nullptr);
      if (rc == PA_OK) rc = launch_block(self.unpack, rbuf, dst, user, nullptr);
    } else {
      rc = launch_block(P->self_fused, src, dst, user, nullptr);
    }
    if (rc != PA_OK) return rc;
    if (timing) {
      CU(cudaEventRecord(S.t[6], user));
      S.timed_once = true;
    }
    return PA_OK;
  }

  if (!B.comm_done_event) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    B.comm_done_event = e;
  }
  cudaEvent_t buf_ev = (cudaEvent_t)B.comm_done_event;

  // fork: the three streams start after the caller's prior work, and after the
  // previous exchange that used these (shared) arenas
  CU(cudaEventRecord(S.ev_start, user));
  for (cudaStream_t s : {S.pack_s, S.comm_s, S.unpack_s}) {
    CU(cudaStreamWaitEvent(s, S.ev_start, 0));
    CU(cudaStreamWaitEvent(s, buf_ev, 0));
  }

  auto send_recv = [&](int to, int from) -> pa_status {
    const Peer& pt = P->peers[to];
    const Peer& pf = P->peers[from];
    if (pt.send_cnt > 0)
      NC(nccl().Send(sbuf + pt.send_off * ES, (size_t)(pt.send_cnt * ES), ncclUint8,
                     pt.world_rank, comm->comm, S.comm_s));
    if (pf.recv_cnt > 0)
      NC(nccl().Recv(rbuf + pf.recv_off * ES, (size_t)(pf.recv_cnt * ES), ncclUint8,
                     pf.world_rank, comm->comm, S.comm_s));
    return PA_OK;
  };

  // ---- PeerPut: one-sided puts over NVLink, no staging ------------------------
  // Each remote block is read from `src` and stored, already permuted, into the
  // destination rank's `dest` through its peer mapping; two tiny grouped
  // send/recv rounds among the line's ranks act as the window fences
  // ("every dest may be overwritten" / "every put has landed").
  // PeerGet is the pull flavour: the remote block is LOADED out of the source
  // rank's `src` (window on src) and stored permuted into the local `dest`.
  if ((P->method == PA_PEER_PUT || P->method == PA_PEER_GET) && !stage_self) {
    const bool get = P->method == PA_PEER_GET;
    auto w = P->windows.find(get ? src : (const void*)dst);
    if (w == P->windows.end()) {
      set_error("one-sided transpose: `%s` has no registered window (pa_plan_set_window)",
                get ? "src" : "dest");
      return PA_ESTATE;
    }
    const std::vector<void*>& win = w->second;
    for (int n = 0; n < nproc; ++n)
      if (n != me && (get ? P->peers[n].recv_cnt : P->peers[n].send_cnt) > 0 && !win[n]) {
        set_error("one-sided transpose: window of peer %d is missing", n + 1);
        return PA_ESTATE;
      }
    auto line_barrier = [&]() -> pa_status {
      if (comm->flags_ready && !g_tun.nccl_fences) {
        // NVLink flag fence: one tiny kernel signals every peer of the line and
        // waits for their signals (a few microseconds instead of an NCCL group)
        if (*comm->fence_err) {
          set_error("an NVLink fence timed out earlier: a peer rank is gone");
          return PA_ECUDA;
        }
        for (int base = 1; base < nproc; base += FENCE_MAX) {
          FenceParams fp;
          fp.n = 0;
          for (int k = base; k < nproc && fp.n < FENCE_MAX; ++k) {
            const int wr = P->peers[(me + k) % nproc].world_rank;
            fp.remote[fp.n] = comm->peer_flags[wr] + comm->rank;
            fp.local[fp.n] = comm->flags + wr;
            fp.seq[fp.n] = ++comm->seq_with[wr];
            ++fp.n;
          }
          fp.timeout_ns = 10ull * 1000 * 1000 * 1000;
          fp.err = comm->fence_err;
          k_fence<<<1, FENCE_MAX, 0, S.comm_s>>>(fp);
          CU(cudaGetLastError());
        }
        return PA_OK;
      }
      NC(nccl().GroupStart());
      for (int k = 1; k < nproc; ++k) {
        const int to = (me + k) % nproc, from = (me - k + nproc) % nproc;
        NC(nccl().Send(S.tok, 4, ncclUint8, P->peers[to].world_rank, comm->comm, S.comm_s));
        NC(nccl().Recv(S.tok + 4 * (1 + from), 4, ncclUint8, P->peers[from].world_rank, comm->comm,
                       S.comm_s));
      }
      NC(nccl().GroupEnd());
      return PA_OK;
    };
    if (timing) CU(cudaEventRecord(S.t[4], S.unpack_s));
    rc = launch_block(P->self_fused, src, dst, S.unpack_s, nullptr);
    if (rc != PA_OK) return rc;
    CU(cudaEventRecord(S.ev_unpack_done, S.unpack_s));
    if (timing) CU(cudaEventRecord(S.t[5], S.unpack_s));
    if (timing) CU(cudaEventRecord(S.t[2], S.comm_s));
    rc = line_barrier();
    if (rc != PA_OK) return rc;
    // fence, remote kernels and closing fence run in order on the high-priority
    // comm stream; the remote kernels' grid is capped (g_tun.remote_ctas) so the
    // self block on the low-priority stream keeps SMs while NVLink is the limit
    for (int k = 1; k < nproc; ++k) {
      const int to = (me + k) % nproc, from = (me - k + nproc) % nproc;
      rc = get ? launch_block(P->peers[from].get, win[from], dst, S.comm_s, nullptr, g_tun.remote_ctas)
               : launch_block(P->peers[to].put, src, win[to], S.comm_s, nullptr, g_tun.remote_ctas);
      if (rc != PA_OK) return rc;
    }
    CU(cudaEventRecord(S.ev_allpacked, S.comm_s));
    if (timing) CU(cudaEventRecord(S.t[1], S.comm_s));
    rc = line_barrier();
    if (rc != PA_OK) return rc;
    CU(cudaEventRecord(S.ev_comm_done, S.comm_s));
    if (timing) CU(cudaEventRecord(S.t[3], S.comm_s));
    CU(cudaStreamWaitEvent(user, S.ev_allpacked, 0));
    CU(cudaStreamWaitEvent(user, S.ev_unpack_done, 0));
    // put: dest is complete only after the closing fence (peers' stores have landed);
    // get: dest is complete once my loads are done, the closing fence only guards
    //      the reuse of `src` -- exactly MPI.Waitall(t)'s role (:127-130).
    S.sends_pending = true;
    if (!get || (flags & PA_WAITALL)) {
      CU(cudaStreamWaitEvent(user, S.ev_comm_done, 0));
      S.sends_pending = false;
    }
    if (timing) {
      CU(cudaEventRecord(S.t[6], user));
      S.timed_once = true;
    }
    return PA_OK;
  }

  // ---- 1. pack (+ exchange) -------------------------------------------------
  if (stage_self) {
    rc = launch_block(self.pack, src, rbuf, S.pack_s, nullptr);  // tail of recv_buf (:393-403)
  } else {
    rc = launch_block(P->self_fused, src, dst, S.unpack_s, nullptr);  // K3, one pass
  }
  if (rc != PA_OK) return rc;

  const bool p2p = P->method != PA_ALLTOALLV;  // PeerPut with aliased arrays runs staged, as PointToPoint
  for (int k = 1; k < nproc; ++k) {
    const int to = (me + k) % nproc;
    rc = launch_block(P->peers[to].pack, src, sbuf, S.pack_s, nullptr);
    if (rc != PA_OK) return rc;
    CU(cudaEventRecord(S.ev_packed[k], S.pack_s));
  }
  CU(cudaEventRecord(S.ev_allpacked, S.pack_s));
  if (timing) CU(cudaEventRecord(S.t[1], S.pack_s));

  if (p2p) {
    for (int k = 1; k < nproc; ++k) {
      const int to = (me + k) % nproc, from = (me - k + nproc) % nproc;
      CU(cudaStreamWaitEvent(S.comm_s, overlap ? S.ev_packed[k] : S.ev_allpacked, 0));
      if (timing && k == 1) CU(cudaEventRecord(S.t[2], S.comm_s));
      NC(nccl().GroupStart());
      rc = send_recv(to, from);
      NC(nccl().GroupEnd());
      if (rc != PA_OK) return rc;
      CU(cudaEventRecord(S.ev_recvd[k], S.comm_s));
    }
  } else {
    // one collective-like group after all packs (MPI.Alltoallv!, :418-427)
    CU(cudaStreamWaitEvent(S.comm_s, S.ev_allpacked, 0));
    if (timing) CU(cudaEventRecord(S.t[2], S.comm_s));
    NC(nccl().GroupStart());
    for (int k = 1; k < nproc && rc == PA_OK; ++k)
      rc = send_recv((me + k) % nproc, (me - k + nproc) % nproc);
    NC(nccl().GroupEnd());
    if (rc != PA_OK) return rc;
  }
  CU(cudaEventRecord(S.ev_comm_done, S.comm_s));
  CU(cudaEventRecord(buf_ev, S.comm_s));
  if (timing) CU(cudaEventRecord(S.t[3], S.comm_s));

  // ---- 2. unpack ------------------------------------------------------------
  // With aliased src/dest no unpack may start before every block is packed:
  // the reference finishes transpose_send! before transpose_recv! (:326-340).
  if (stage_self) CU(cudaStreamWaitEvent(S.unpack_s, S.ev_allpacked, 0));
  if (timing) CU(cudaEventRecord(S.t[4], S.unpack_s));
  if (stage_self) {
    rc = launch_block(self.unpack, rbuf, dst, S.unpack_s, nullptr);  // local data first (:511)
    if (rc != PA_OK) return rc;
  }
  if (p2p && overlap) {
    for (int k = 1; k < nproc; ++k) {
      const int from = (me - k + nproc) % nproc;
      CU(cudaStreamWaitEvent(S.unpack_s, S.ev_recvd[k], 0));
      rc = launch_block(P->peers[from].unpack, rbuf, dst, S.unpack_s, nullptr);
      if (rc != PA_OK) return rc;
    }
  } else {
    CU(cudaStreamWaitEvent(S.unpack_s, S.ev_comm_done, 0));
    if (timing && !stage_self) CU(cudaEventRecord(S.t[4], S.unpack_s));
    for (int n = 0; n < nproc; ++n) {  // n = 1..Nproc in order (:508-509)
      if (n == me) continue;
      rc = launch_block(P->peers[n].unpack, rbuf, dst, S.unpack_s, nullptr);
      if (rc != PA_OK) return rc;
    }
  }
  CU(cudaEventRecord(S.ev_unpack_done, S.unpack_s));
  if (timing) CU(cudaEventRecord(S.t[5], S.unpack_s));

  // ---- join -----------------------------------------------------------------
  CU(cudaStreamWaitEvent(user, S.ev_allpacked, 0));    // src may be reused by the caller
  CU(cudaStreamWaitEvent(user, S.ev_unpack_done, 0));  // dst complete
  S.sends_pending = true;
  if (flags & PA_WAITALL) {
    CU(cudaStreamWaitEvent(user, S.ev_comm_done, 0));  // MPI.Waitall(t) (:174-176)
    S.sends_pending = false;
  }
  if (timing) {
    CU(cudaEventRecord(S.t[6], user));
    S.timed_once = true;
  }
  return PA_OK;
}

// MPI.Waitall(t::Transposition) (Transpositions.jl:127-130)
pa_status wait_sends(Plan* P, void* stream) {
  if (!P->st || !P->st->sends_pending) return PA_OK;
  CU(cudaStreamWaitEvent((cudaStream_t)stream, P->st->ev_comm_done, 0));
  P->st->sends_pending = false;
  return PA_OK;
}

pa_status transpose_host(Plan* P, Comm* comm, const void* hsrc, void* hdst, unsigned flags) {
  pa_status rc = ensure_state(P);
  if (rc != PA_OK) return rc;
  const size_t nin = (size_t)(P->length_in * P->elsize), nout = (size_t)(P->length_out * P->elsize);
  if (!P->h_src_dev) CU(cudaMalloc(&P->h_src_dev, std::max<size_t>(nin, 1)));
  if (!P->h_dst_dev) CU(cudaMalloc(&P->h_dst_dev, std::max<size_t>(nout, 1)));
  cudaStream_t s = P->st->host_s;
  CU(cudaMemcpyAsync(P->h_src_dev, hsrc, nin, cudaMemcpyHostToDevice, s));
  rc = transpose(P, comm, P->h_src_dev, P->h_dst_dev, flags | PA_WAITALL, s);
  if (rc != PA_OK) return rc;
  CU(cudaMemcpyAsync(hdst, P->h_dst_dev, nout, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return PA_OK;
}

}  // namespace pa
