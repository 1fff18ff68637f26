// extern "C" surface of libpa_b200 (see include/pa_b200.h).  Converts between
// the reference's 1-based / inclusive conventions and the 0-based internals.
#include <algorithm>
#include <new>
#include <vector>

#include "pa_internal.hpp"

using namespace pa;

#define GUARD(expr)                    \
  try {                                \
    expr                               \
  } catch (const std::bad_alloc&) {    \
    set_error("out of host memory");   \
    return PA_ENOMEM;                  \
  } catch (...) {                      \
    set_error("unexpected exception"); \
    return PA_EINVAL;                  \
  }

extern "C" {

const char* pa_version(void) { return "pa_b200 0.2.0 (sm_100a)"; }

const char* pa_strerror(pa_status s) {
  switch (s) {
    case PA_OK: return "success";
    case PA_EINVAL: return "invalid argument";
    case PA_EINCOMPAT: return "pencil configurations are not compatible for transposition";
    case PA_EDIM: return "array dimensions do not match the pencil";
    case PA_ECUDA: return "CUDA error";
    case PA_ENCCL: return "NCCL error";
    case PA_ENOMEM: return "out of memory";
    case PA_ESTATE: return "invalid call sequence";
    case PA_ENOGPU: return "no CUDA device (there is no CPU fallback)";
  }
  return "unknown status";
}

const char* pa_last_error(void) { return last_error(); }
int64_t pa_launch_count(void) { return launch_count(); }
int pa_device_count(void) { return device_count(); }
pa_status pa_set_device(int device) {
  if (device_count() == 0) {
    set_error("no CUDA device");
    return PA_ENOGPU;
  }
  return set_device(device);
}

pa_status pa_set_tunable(const char* name, int64_t value) {
  if (!name) return PA_EINVAL;
  if (!strcmp(name, "remote_ctas")) g_tun.remote_ctas = (int)value;
  else if (!strcmp(name, "box_copy_ctas")) g_tun.box_copy_ctas = (int)value;
  else if (!strcmp(name, "nccl_fences")) g_tun.nccl_fences = (int)value;
  else if (!strcmp(name, "bulk_rows")) g_tun.bulk_rows = (int)value;
  else if (!strcmp(name, "nccl_register")) g_tun.nccl_register = (int)value;
  else if (!strcmp(name, "transpose_tbq")) g_tun.transpose_tbq = (int)value;
  else if (!strcmp(name, "small_block_bytes")) g_tun.small_block_bytes = value;
  else if (!strcmp(name, "transpose_y_fastest")) g_tun.transpose_y_fastest = (int)value;
  else if (!strcmp(name, "multi_put")) g_tun.multi_put = (int)value;
  else if (!strcmp(name, "fft_lines")) g_tun.fft_lines = (int)value;
  else if (!strcmp(name, "oneside_self_ctas")) g_tun.oneside_self_ctas = (int)value;
  else if (!strcmp(name, "p2p_chunks")) g_tun.p2p_chunks = (int)value;
  else if (!strcmp(name, "staged_ctas")) g_tun.staged_ctas = (int)value;
  else if (!strcmp(name, "self_first")) g_tun.self_first = (int)value;
  else if (!strcmp(name, "ipc_exchange")) g_tun.ipc_exchange = (int)value;
  else if (!strcmp(name, "fence_timeout_ms")) g_tun.fence_timeout_ms = value;
  else if (!strcmp(name, "pdl")) g_tun.pdl = (int)value;
  else if (!strcmp(name, "nccl_ctas")) g_tun.nccl_ctas = (int)value;
  else if (!strcmp(name, "host_chunk_bytes")) g_tun.host_chunk_bytes = value;
  else if (!strcmp(name, "host_slots")) g_tun.host_slots = (int)value;
  else {
    set_error("unknown tunable '%s'", name);
    return PA_EINVAL;
  }
  return PA_OK;
}

// ---- topology -----------------------------------------------------------------
pa_status pa_dims_create(int nprocs, int M, int64_t* dims) {
  if (nprocs < 1 || M < 1 || M > PA_MAX_TOPO || !dims) {
    set_error("pa_dims_create: bad arguments");
    return PA_EINVAL;
  }
  // balanced factorisation: hand the prime factors, largest first, to the
  // currently smallest dimension; report in non-increasing order
  for (int i = 0; i < M; ++i) dims[i] = 1;
  int64_t primes[64];
  int np = 0;
  int n = nprocs;
  for (int f = 2; (int64_t)f * f <= n; ++f)
    while (n % f == 0) {
      primes[np++] = f;
      n /= f;
    }
  if (n > 1) primes[np++] = n;
  for (int k = np - 1; k >= 0; --k) {
    int j = 0;
    for (int i = 1; i < M; ++i)
      if (dims[i] < dims[j]) j = i;
    dims[j] *= primes[k];
  }
  std::sort(dims, dims + M, [](int64_t a, int64_t b) { return a > b; });
  return PA_OK;
}

pa_status pa_topology_create(int M, const int64_t* dims, int world_rank, pa_topology** out) {
  GUARD({
    if (M < 1 || M > PA_MAX_TOPO || !dims || !out) {
      set_error("topology must have 1..%d dimensions", PA_MAX_TOPO);
      return PA_EINVAL;
    }
    auto t = std::make_shared<Topology>();
    t->M = M;
    int64_t size = 1;
    for (int i = 0; i < M; ++i) {
      if (dims[i] < 1) {
        set_error("process grid dimensions must be >= 1");
        return PA_EINVAL;
      }
      t->dims[i] = dims[i];
      size *= dims[i];
    }
    if (world_rank < 0 || world_rank >= size) {
      set_error("rank %d outside of a %lld-process grid", world_rank, (long long)size);
      return PA_EINVAL;
    }
    t->size = (int)size;
    t->rank = world_rank;
    t->coords_of(world_rank, t->coords);
    *out = new pa_topology{t};
    return PA_OK;
  })
}

void pa_topology_destroy(pa_topology* t) { delete t; }

pa_status pa_topology_info(const pa_topology* t, int* M, int64_t* dims, int* world_rank,
                           int* world_size, int64_t* coords_local) {
  if (!t) return PA_EINVAL;
  const Topology& T = *t->p;
  if (M) *M = T.M;
  if (world_rank) *world_rank = T.rank;
  if (world_size) *world_size = T.size;
  for (int i = 0; i < T.M; ++i) {
    if (dims) dims[i] = T.dims[i];
    if (coords_local) coords_local[i] = T.coords[i] + 1;
  }
  return PA_OK;
}

pa_status pa_topology_rank_of(const pa_topology* t, const int64_t* coords, int* rank) {
  if (!t || !coords || !rank) return PA_EINVAL;
  const Topology& T = *t->p;
  int64_t c[PA_MAX_TOPO];
  for (int i = 0; i < T.M; ++i) {
    if (coords[i] < 1 || coords[i] > T.dims[i]) {
      set_error("coordinate %d out of range", i + 1);
      return PA_EINVAL;
    }
    c[i] = coords[i] - 1;
  }
  *rank = T.rank_of(c);
  return PA_OK;
}

pa_status pa_topology_line(const pa_topology* t, int R, int* ranks) {
  if (!t || !ranks) return PA_EINVAL;
  const Topology& T = *t->p;
  if (R < 1 || R > T.M) {
    set_error("grid dimension %d out of range 1:%d", R, T.M);
    return PA_EINVAL;
  }
  int64_t c[PA_MAX_TOPO];
  for (int i = 0; i < T.M; ++i) c[i] = T.coords[i];
  for (int64_t n = 0; n < T.dims[R - 1]; ++n) {
    c[R - 1] = n;
    ranks[n] = T.rank_of(c);
  }
  return PA_OK;
}

// ---- pencil -------------------------------------------------------------------
pa_status pa_pencil_create(pa_topology* topo, int N, const int64_t* size_global,
                           const int* decomp_dims, const int* perm, pa_pencil* share_with,
                           pa_pencil** out) {
  GUARD({
    if (!topo || !size_global || !decomp_dims || !out) return PA_EINVAL;
    const int M = topo->p->M;
    if (N < 1 || N > PA_MAX_DIMS) {
      set_error("number of dimensions must be in 1:%d", PA_MAX_DIMS);
      return PA_EINVAL;
    }
    // _check_selected_dimensions (Pencils.jl:397-412)
    if (M > N) {
      set_error("number of decomposed dimensions `M` cannot be larger than N = %d (got M = %d)", N,
                M);
      return PA_EINVAL;
    }
    auto p = std::make_shared<Pencil>();
    p->topo = topo->p;
    p->N = N;
    for (int d = 0; d < N; ++d) {
      if (size_global[d] < 0) {
        set_error("negative global size");
        return PA_EINVAL;
      }
      p->size_global[d] = size_global[d];
    }
    for (int i = 0; i < M; ++i) {
      if (decomp_dims[i] < 1 || decomp_dims[i] > N) {
        set_error("dimensions must be in 1:%d", N);
        return PA_EINVAL;
      }
      for (int j = 0; j < i; ++j)
        if (decomp_dims[j] == decomp_dims[i]) {
          set_error("dimensions may not be repeated");
          return PA_EINVAL;
        }
      p->decomp[i] = decomp_dims[i] - 1;
    }
    bool seen[PA_MAX_DIMS] = {false};
    p->perm_identity = true;
    for (int d = 0; d < N; ++d) {
      int v = perm ? perm[d] : d + 1;
      if (v < 1 || v > N || seen[v - 1]) {  // check_permutation (Pencils.jl:384-387)
        set_error("invalid permutation of dimensions");
        return PA_EINVAL;
      }
      seen[v - 1] = true;
      p->perm[d] = v - 1;
      if (v - 1 != d) p->perm_identity = false;
    }
    p->bufs = share_with ? share_with->p->bufs : std::make_shared<Buffers>();
    *out = new pa_pencil{p};
    return PA_OK;
  })
}

void pa_pencil_destroy(pa_pencil* p) { delete p; }

pa_status pa_pencil_range(const pa_pencil* p, const int64_t* coords, int memory_order, int64_t* lo,
                          int64_t* hi) {
  if (!p || !lo || !hi) return PA_EINVAL;
  const Pencil& P = *p->p;
  int64_t c[PA_MAX_TOPO], l[PA_MAX_DIMS], h[PA_MAX_DIMS];
  for (int i = 0; i < P.topo->M; ++i) {
    if (coords) {
      if (coords[i] < 1 || coords[i] > P.topo->dims[i]) {
        set_error("coordinate %d out of range", i + 1);
        return PA_EINVAL;
      }
      c[i] = coords[i] - 1;
    } else {
      c[i] = P.topo->coords[i];
    }
  }
  P.range_of(c, l, h);
  for (int m = 0; m < P.N; ++m) {
    int d = memory_order ? P.perm[m] : m;  // (perm * t)[m] = t[perm[m]]
    lo[m] = l[d] + 1;
    hi[m] = h[d];
  }
  return PA_OK;
}

pa_status pa_pencil_size_local(const pa_pencil* p, int memory_order, int64_t* dims) {
  if (!p || !dims) return PA_EINVAL;
  int64_t lo[PA_MAX_DIMS], hi[PA_MAX_DIMS];
  pa_status s = pa_pencil_range(p, nullptr, memory_order, lo, hi);
  if (s != PA_OK) return s;
  for (int d = 0; d < p->p->N; ++d) dims[d] = hi[d] - lo[d] + 1;
  return PA_OK;
}

pa_status pa_pencil_buffers(const pa_pencil* p, void** send_buf, int64_t* send_cap,
                            void** recv_buf, int64_t* recv_cap) {
  if (!p) return PA_EINVAL;
  const Buffers& b = *p->p->bufs;
  if (send_buf) *send_buf = b.send;
  if (send_cap) *send_cap = b.send_cap;
  if (recv_buf) *recv_buf = b.recv;
  if (recv_cap) *recv_cap = b.recv_cap;
  return PA_OK;
}

pa_status pa_pencil_reserve(pa_pencil* p, int64_t send_bytes, int64_t recv_bytes) {
  if (!p) return PA_EINVAL;
  if (device_count() == 0) {
    set_error("no CUDA device");
    return PA_ENOGPU;
  }
  return p->p->bufs->reserve(send_bytes, recv_bytes);
}

// ---- plan ---------------------------------------------------------------------
pa_status pa_plan_create(pa_pencil* pin, pa_pencil* pout, int n_extra, const int64_t* extra_dims,
                         int elsize, pa_method method, pa_plan** out) {
  GUARD({
    if (!pin || !pout || !out || (n_extra > 0 && !extra_dims)) return PA_EINVAL;
    Plan* P = nullptr;
    pa_status s = build_plan(pin->p, pout->p, n_extra, extra_dims, elsize, (int)method, &P);
    if (s != PA_OK) return s;
    *out = new pa_plan{P};
    return PA_OK;
  })
}

void pa_plan_destroy(pa_plan* plan) {
  if (!plan) return;
  delete plan->p;
  delete plan;
}

pa_status pa_plan_get_info(const pa_plan* plan, pa_plan_info* info) {
  if (!plan || !info) return PA_EINVAL;
  const Plan& P = *plan->p;
  info->dim = P.dim + 1;
  info->nproc = P.nproc;
  info->self_index = P.self_index + 1;
  info->same_perm = P.same_perm;
  info->elsize = P.elsize;
  info->method = P.method;
  info->length_in = P.length_in;
  info->length_out = P.length_out;
  info->length_self = P.length_self;
  info->send_bytes = P.send_elems * P.elsize;
  info->recv_bytes = P.recv_elems * P.elsize;
  return PA_OK;
}

pa_status pa_plan_get_peer(const pa_plan* plan, int n, pa_peer_info* info) {
  if (!plan || !info) return PA_EINVAL;
  const Plan& P = *plan->p;
  if (P.dim < 0 || n < 1 || n > P.nproc) {
    set_error("peer index %d out of range", n);
    return PA_EINVAL;
  }
  const Peer& pr = P.peers[n - 1];
  const int64_t S = P.elsize;
  info->world_rank = pr.world_rank;
  info->is_self = pr.is_self;
  info->send_offset = pr.send_off * S;
  info->send_count = pr.send_cnt * S;
  info->recv_offset = pr.recv_off * S;
  info->recv_count = pr.recv_cnt * S;
  info->remote_recv_offset = pr.remote_recv_off * S;
  return PA_OK;
}

static void export_block(const BlockCopy& b, pa_block_desc* d) {
  memset(d, 0, sizeof *d);
  d->nd = b.nd_raw;
  for (int i = 0; i < b.nd_raw; ++i) {
    d->extent[i] = b.raw[i].e;
    d->src_stride[i] = b.raw[i].ss;
    d->dst_stride[i] = b.raw[i].ds;
  }
  d->src_offset = b.src_off;
  d->dst_offset = b.dst_off;
  d->kernel_class = b.klass;
  d->vec_bytes = b.stride_align;
}

pa_status pa_plan_get_block(const pa_plan* plan, int op, int n, pa_block_desc* desc) {
  if (!plan || !desc) return PA_EINVAL;
  const Plan& P = *plan->p;
  if (op == 2) {
    export_block(P.self_fused, desc);
    return PA_OK;
  }
  if (P.dim < 0 || n < 1 || n > P.nproc || op < 0 || op > 4) {
    set_error("bad block selector (op=%d, n=%d)", op, n);
    return PA_EINVAL;
  }
  const Peer& pr = P.peers[n - 1];
  export_block(op == 0 ? pr.pack : op == 1 ? pr.unpack : op == 3 ? pr.put : pr.get, desc);
  return PA_OK;
}

pa_status pa_plan_get_chunk(const pa_plan* plan, int op, int n, int part, int nparts,
                            pa_block_desc* desc, int64_t* wire_offset, int64_t* wire_bytes) {
  if (!plan || !desc) return PA_EINVAL;
  const Plan& P = *plan->p;
  if (P.dim < 0 || n < 1 || n > P.nproc || op < 0 || op > 1 || nparts < 1 || part < 0 ||
      part >= nparts) {
    set_error("bad chunk selector (op=%d, n=%d, part=%d/%d)", op, n, part, nparts);
    return PA_EINVAL;
  }
  const Peer& pr = P.peers[n - 1];
  const BlockCopy c = sub_block(op == 0 ? pr.pack : pr.unpack, part, nparts, op == 0, nullptr, nullptr);
  export_block(c, desc);
  if (wire_offset) *wire_offset = (op == 0 ? c.dst_off : c.src_off) * c.elsize;
  if (wire_bytes) *wire_bytes = c.count * c.elsize;
  return PA_OK;
}

// ---- kernels --------------------------------------------------------------------
static pa_status need_gpu() {
  if (device_count() == 0) {
    set_error("no CUDA device: the transpose! path has no CPU fallback");
    return PA_ENOGPU;
  }
  return PA_OK;
}

pa_status pa_pack(pa_plan* plan, int n, const void* src, void* buf, void* stream) {
  if (!plan) return PA_EINVAL;
  Plan& P = *plan->p;
  if (P.dim < 0 || n < 1 || n > P.nproc) {
    set_error("peer index %d out of range", n);
    return PA_EINVAL;
  }
  pa_status s = need_gpu();
  if (s != PA_OK) return s;
  return launch_block(P.peers[n - 1].pack, src, buf, stream, nullptr);
}

pa_status pa_unpack(pa_plan* plan, int n, const void* recv_buf, void* dst, void* stream) {
  if (!plan) return PA_EINVAL;
  Plan& P = *plan->p;
  if (P.dim < 0 || n < 1 || n > P.nproc) {
    set_error("peer index %d out of range", n);
    return PA_EINVAL;
  }
  pa_status s = need_gpu();
  if (s != PA_OK) return s;
  return launch_block(P.peers[n - 1].unpack, recv_buf, dst, stream, nullptr);
}

pa_status pa_put(pa_plan* plan, int n, const void* src, void* peer_dst, void* stream) {
  if (!plan) return PA_EINVAL;
  Plan& P = *plan->p;
  if (P.dim < 0 || n < 1 || n > P.nproc || P.peers[n - 1].is_self) {
    set_error("pa_put: peer index %d out of range (or self)", n);
    return PA_EINVAL;
  }
  pa_status s = need_gpu();
  if (s != PA_OK) return s;
  return launch_block(P.peers[n - 1].put, src, peer_dst, stream, nullptr);
}

pa_status pa_get(pa_plan* plan, int n, const void* peer_src, void* dst, void* stream) {
  if (!plan) return PA_EINVAL;
  Plan& P = *plan->p;
  if (P.dim < 0 || n < 1 || n > P.nproc || P.peers[n - 1].is_self) {
    set_error("pa_get: peer index %d out of range (or self)", n);
    return PA_EINVAL;
  }
  pa_status s = need_gpu();
  if (s != PA_OK) return s;
  return launch_block(P.peers[n - 1].get, peer_src, dst, stream, nullptr);
}

static pa_status all_blocks(pa_plan* plan, bool get, const void* local, void* const* peers,
                            int max_ctas, void* stream) {
  if (!plan || !peers) return PA_EINVAL;
  Plan& P = *plan->p;
  if (P.dim < 0) {
    set_error("plan has no exchange");
    return PA_EINVAL;
  }
  pa_status s = need_gpu();
  if (s != PA_OK) return s;
  std::vector<const BlockCopy*> blocks;
  std::vector<const void*> srcs;
  std::vector<void*> dsts;
  for (int k = 1; k < P.nproc; ++k) {
    const int n = get ? (P.self_index - k + P.nproc) % P.nproc : (P.self_index + k) % P.nproc;
    const BlockCopy& b = get ? P.peers[n].get : P.peers[n].put;
    if (b.klass == KC_EMPTY) continue;
    blocks.push_back(&b);
    srcs.push_back(get ? (const void*)peers[n] : local);
    dsts.push_back(get ? (void*)local : peers[n]);
  }
  if (blocks.empty()) return PA_OK;
  if (max_ctas == 0) max_ctas = g_tun.remote_ctas;
  s = launch_multi((int)blocks.size(), blocks.data(), srcs.data(), dsts.data(), stream, max_ctas,
                   nullptr);
  if (s != PA_EINCOMPAT) return s;
  for (size_t i = 0; i < blocks.size(); ++i) {
    s = launch_block(*blocks[i], srcs[i], dsts[i], stream, nullptr, max_ctas);
    if (s != PA_OK) return s;
  }
  return PA_OK;
}

pa_status pa_put_all(pa_plan* plan, const void* src, void* const* peers, int max_ctas,
                     void* stream) {
  GUARD({ return all_blocks(plan, false, src, peers, max_ctas, stream); })
}

pa_status pa_get_all(pa_plan* plan, void* const* peers, void* dst, int max_ctas, void* stream) {
  GUARD({ return all_blocks(plan, true, dst, peers, max_ctas, stream); })
}

pa_status pa_copy_self(pa_plan* plan, const void* src, void* dst, void* stream) {
  if (!plan) return PA_EINVAL;
  pa_status s = need_gpu();
  if (s != PA_OK) return s;
  return launch_block(plan->p->self_fused, src, dst, stream, nullptr);
}

pa_status pa_permute_local(pa_plan* plan, const void* src, void* dst, void* scratch,
                           void* stream) {
  if (!plan) return PA_EINVAL;
  return permute_local(plan->p, src, dst, scratch, stream);
}

pa_status pa_box_copy(int nd, const int64_t* extent, const int64_t* src_stride,
                      const int64_t* dst_stride, int elsize, const void* src, void* dst,
                      void* stream, pa_block_desc* chosen) {
  if (nd < 0 || nd > PA_MAX_DIMS || (nd > 0 && (!extent || !src_stride || !dst_stride))) {
    set_error("pa_box_copy: 0..%d dimensions", PA_MAX_DIMS);
    return PA_EINVAL;
  }
  if (elsize != 1 && elsize != 2 && elsize != 4 && elsize != 8 && elsize != 16) {
    set_error("pa_box_copy: element size must be 1, 2, 4, 8 or 16");
    return PA_EINVAL;
  }
  BlockCopy b;
  b.elsize = elsize;
  b.nd_raw = nd;
  for (int i = 0; i < nd; ++i) {
    if (extent[i] < 0 || src_stride[i] < 0 || dst_stride[i] < 0) {
      set_error("pa_box_copy: negative extent or stride");
      return PA_EINVAL;
    }
    b.raw[i] = Dim{extent[i], src_stride[i], dst_stride[i]};
  }
  canonicalize(b);
  int vec = 0;
  pa_status s = PA_OK;
  if (src || dst) {
    s = need_gpu();
    if (s != PA_OK) return s;
    s = launch_block(b, src, dst, stream, &vec, g_tun.box_copy_ctas);
  }
  if (chosen) {
    export_block(b, chosen);
    if (vec) chosen->vec_bytes = vec;
  }
  return s;
}

// ---- communicator ----------------------------------------------------------------
pa_status pa_comm_unique_id(void* id128) {
  if (!id128) return PA_EINVAL;
  return comm_unique_id(id128);
}

pa_status pa_comm_init_rank(const void* id128, int nranks, int rank, pa_comm** out) {
  GUARD({
    if (!id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) return PA_EINVAL;
    Comm* c = nullptr;
    pa_status s = comm_init(id128, nranks, rank, &c);
    if (s != PA_OK) return s;
    *out = new pa_comm{c};
    return PA_OK;
  })
}

pa_status pa_comm_init_local(int nranks, int rank, pa_comm** out) {
  GUARD({
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) return PA_EINVAL;
    Comm* c = nullptr;
    pa_status s = comm_init_local(nranks, rank, &c);
    if (s != PA_OK) return s;
    *out = new pa_comm{c};
    return PA_OK;
  })
}

pa_status pa_comm_flags_export(pa_comm* c, void* handle64, int64_t* offset) {
  if (!c || !handle64 || !offset) return PA_EINVAL;
  return comm_flags_export(c->p, handle64, offset);
}

pa_status pa_comm_flags_import(pa_comm* c, int rank, const void* handle64, int64_t offset) {
  GUARD({
    if (!c || !handle64) return PA_EINVAL;
    return comm_flags_import(c->p, rank, handle64, offset);
  })
}

void pa_comm_destroy(pa_comm* c) {
  if (!c) return;
  comm_destroy(c->p);
  delete c;
}

// ---- PeerPut windows ---------------------------------------------------------------
pa_status pa_ipc_export(const void* devptr, void* handle64, int64_t* offset) {
  if (!devptr || !handle64 || !offset) return PA_EINVAL;
  return ipc_export(devptr, handle64, offset);
}

pa_status pa_ipc_import(const void* handle64, int64_t offset, void** mapped) {
  GUARD({
    if (!handle64 || !mapped) return PA_EINVAL;
    return ipc_import(handle64, offset, mapped);
  })
}

pa_status pa_ipc_release(const void* handle64) {
  GUARD({
    if (!handle64) return PA_EINVAL;
    return ipc_release_handle(handle64);
  })
}

pa_status pa_plan_set_window(pa_plan* plan, const void* local_dst, int n, void* peer_dst) {
  GUARD({
    // local_dst may be NULL: a rank that owns nothing of `dest` still puts into its peers
    if (!plan) return PA_EINVAL;
    return plan_set_window(plan->p, local_dst, n - 1, peer_dst);
  })
}

pa_status pa_plan_set_recv_window(pa_plan* plan, int n, void* peer_recv_buf) {
  GUARD({
    if (!plan) return PA_EINVAL;
    return plan_set_recv_window(plan->p, n - 1, peer_recv_buf);
  })
}

// ---- transpose! ------------------------------------------------------------------
pa_status pa_transpose(pa_plan* plan, pa_comm* comm, const void* src, void* dst, unsigned flags,
                       void* stream) {
  GUARD({
    if (!plan) {
      set_error("pa_transpose: null plan");
      return PA_EINVAL;
    }
    // src / dst may be NULL on a rank whose local array is empty (checked against the plan)
    return transpose(plan->p, comm ? comm->p : nullptr, src, dst, flags, stream);
  })
}

pa_status pa_wait(pa_plan* plan, void* stream) {
  if (!plan) return PA_EINVAL;
  return wait_sends(plan->p, stream);
}

pa_status pa_transpose_host(pa_plan* plan, pa_comm* comm, const void* host_src, void* host_dst,
                            unsigned flags) {
  GUARD({
    if (!plan) return PA_EINVAL;
    return transpose_host(plan->p, comm ? comm->p : nullptr, host_src, host_dst, flags);
  })
}

pa_status pa_host_chain_create(int n, pa_plan* const* plans, pa_comm* comm, pa_host_chain** out) {
  GUARD({
    if (n < 1 || n > 64 || !plans || !out) {
      set_error("pa_host_chain_create: 1..64 plans");
      return PA_EINVAL;
    }
    Plan* ps[64];
    for (int i = 0; i < n; ++i) {
      if (!plans[i]) return PA_EINVAL;
      ps[i] = plans[i]->p;
    }
    HostChain* c = nullptr;
    pa_status s = host_chain_create(n, ps, comm ? comm->p : nullptr, &c);
    if (s != PA_OK) return s;
    *out = new pa_host_chain{c};
    return PA_OK;
  })
}

void pa_host_chain_destroy(pa_host_chain* c) {
  if (!c) return;
  host_chain_destroy(c->p);
  delete c;
}

pa_status pa_host_chain_submit(pa_host_chain* c, const void* host_src, void* host_dst,
                               int64_t* ticket) {
  GUARD({
    if (!c) return PA_EINVAL;
    return host_chain_submit(c->p, host_src, host_dst, ticket);
  })
}

pa_status pa_host_chain_wait(pa_host_chain* c, int64_t ticket) {
  if (!c) return PA_EINVAL;
  return host_chain_wait(c->p, ticket);
}

pa_status pa_host_chain_time_begin(pa_host_chain* c) {
  if (!c) return PA_EINVAL;
  return host_chain_time_begin(c->p);
}

pa_status pa_host_chain_time_end(pa_host_chain* c, float* ms) {
  if (!c || !ms) return PA_EINVAL;
  return host_chain_time_end(c->p, ms);
}

pa_status pa_host_chain_buffer(pa_host_chain* c, int slot, int which, void** devptr,
                               int64_t* bytes) {
  if (!c) return PA_EINVAL;
  return host_chain_buffer(c->p, slot, which, devptr, bytes);
}

// ---- PencilIO binary layout --------------------------------------------------------
pa_status pa_io_sizes(const pa_pencil* p, int n_extra, const int64_t* extra_dims, int elsize,
                      int chunks, int64_t* global_bytes, int64_t* local_bytes, int64_t* nruns,
                      int64_t* run_bytes, int64_t* first_offset) {
  GUARD({
    if (!p || (n_extra > 0 && !extra_dims)) return PA_EINVAL;
    return io_sizes(*p->p, n_extra, extra_dims, elsize, chunks, global_bytes, local_bytes, nruns,
                    run_bytes, first_offset);
  })
}

pa_status pa_io_run_offset(const pa_pencil* p, int n_extra, const int64_t* extra_dims, int elsize,
                           int chunks, int64_t run, int64_t* file_offset) {
  GUARD({
    if (!p || !file_offset || (n_extra > 0 && !extra_dims)) return PA_EINVAL;
    return io_run_offset(*p->p, n_extra, extra_dims, elsize, chunks, run, file_offset);
  })
}

pa_status pa_io_write(const pa_pencil* p, int n_extra, const int64_t* extra_dims, int elsize,
                      int chunks, const void* dev_array, const char* path, int64_t offset) {
  GUARD({
    if (!p || !path || offset < 0 || (n_extra > 0 && !extra_dims)) return PA_EINVAL;
    return io_transfer(*p->p, n_extra, extra_dims, elsize, chunks, (void*)dev_array, path, offset, true);
  })
}

pa_status pa_io_read(const pa_pencil* p, int n_extra, const int64_t* extra_dims, int elsize,
                     int chunks, void* dev_array, const char* path, int64_t offset) {
  GUARD({
    if (!p || !path || offset < 0 || (n_extra > 0 && !extra_dims)) return PA_EINVAL;
    return io_transfer(*p->p, n_extra, extra_dims, elsize, chunks, dev_array, path, offset, false);
  })
}

pa_status pa_plan_timings(pa_plan* plan, pa_timings* t) {
  if (!plan || !t) return PA_EINVAL;
  return plan_timings(plan->p, t);
}

pa_status pa_plan_enable_timing(pa_plan* plan, int on) {
  if (!plan) return PA_EINVAL;
  return plan_enable_timing(plan->p, on);
}

}  // extern "C"
