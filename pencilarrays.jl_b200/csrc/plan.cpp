// Host-side geometry and planning: MPITopology / Pencil restated as integer
// arithmetic, and the Transposition plan reduced to strided-box-copy
// descriptors.  No CUDA in this file.
//
// Reference behaviour followed (never copied; the reference is Julia):
//   data_ranges.jl:4-9,15-45      block partition, axes_all
//   MPITopologies.jl:125-131      row-major rank grid (reorder=false)
//   Pencils.jl:221-236,579-587    axes_local, to_local, memory order = perm * logical
//   Transpositions.jl:93-118      compatibility checks + `dim` discovery
//   Transpositions.jl:302-317     length_self / send / recv sizes
//   Transpositions.jl:380-416     per-peer ranges, buffer offsets, self block at the tail
//   Transpositions.jl:516-529     unpack geometry (o_range_iperm, relative permutation)
#include <algorithm>
#include <cstdarg>
#include <cstdio>

#include "pa_internal.hpp"

namespace pa {

// ---- errors -----------------------------------------------------------------
static thread_local char g_err[512] = {0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ---- Topology ---------------------------------------------------------------
void Topology::coords_of(int r, i64* c) const {
  for (int i = M - 1; i >= 0; --i) {
    c[i] = r % dims[i];
    r /= (int)dims[i];
  }
}
int Topology::rank_of(const i64* c) const {
  i64 r = 0;
  for (int i = 0; i < M; ++i) r = r * dims[i] + c[i];
  return (int)r;
}
bool Topology::same_as(const Topology& o) const {
  if (M != o.M || rank != o.rank || size != o.size) return false;
  for (int i = 0; i < M; ++i)
    if (dims[i] != o.dims[i]) return false;
  return true;
}

// ---- Pencil -----------------------------------------------------------------
void Pencil::range_of(const i64* coords, i64* lo, i64* hi) const {
  for (int d = 0; d < N; ++d) {
    lo[d] = 0;
    hi[d] = size_global[d];
  }
  // topology dim i decomposes array dim decomp[i]; the ORDER of decomp matters
  // (complete_dims, data_ranges.jl:15-26)
  for (int i = 0; i < topo->M; ++i) {
    int d = decomp[i];
    local_data_range(coords[i], topo->dims[i], size_global[d], &lo[d], &hi[d]);
  }
}

// ---- canonical form of a strided box copy -----------------------------------
static int pow2_divisor(i64 x, int cap) {
  if (x == 0) return cap;
  if (x < 0) x = -x;
  int a = 1;
  while (a < cap && (x % (2 * a)) == 0) a *= 2;
  return a;
}

void canonicalize(BlockCopy& b) {
  b.count = 1;
  for (int i = 0; i < b.nd_raw; ++i) b.count *= b.raw[i].e;
  b.nd = 0;
  b.klass = KC_EMPTY;
  b.contiguous_both = false;
  if (b.count == 0) return;

  Dim t[PA_MAX_DIMS + 1];
  int n = 0;
  for (int i = 0; i < b.nd_raw; ++i)
    if (b.raw[i].e != 1) t[n++] = b.raw[i];
  if (n == 0) {
    t[0] = Dim{1, 1, 1};
    n = 1;
  }
  std::stable_sort(t, t + n, [](const Dim& a, const Dim& c) { return a.ss < c.ss; });
  // merge dims that are contiguous with their predecessor on BOTH sides
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (m > 0 && t[i].ss == t[m - 1].ss * t[m - 1].e && t[i].ds == t[m - 1].ds * t[m - 1].e) {
      t[m - 1].e *= t[i].e;
    } else {
      t[m++] = t[i];
    }
  }
  n = m;
  // Y = dim with the smallest destination stride
  int j = 0;
  for (int i = 1; i < n; ++i)
    if (t[i].ds < t[j].ds) j = i;
  if (j == 0) {
    b.klass = (t[0].ss == 1 && t[0].ds == 1) ? KC_ROWS : KC_TILE_SCALAR;
  } else {
    Dim y = t[j];
    for (int i = j; i > 1; --i) t[i] = t[i - 1];
    t[1] = y;
    b.klass = (t[0].ss == 1 && t[1].ds == 1) ? KC_TRANSPOSE : KC_TILE_SCALAR;
  }
  b.nd = n;
  for (int i = 0; i < n; ++i) b.d[i] = t[i];
  b.contiguous_both = (n == 1 && t[0].ss == 1 && t[0].ds == 1);

  // largest power-of-two access width every stride / run length allows, per side:
  // the source side constrains the loads, the destination side the stores
  const i64 S = b.elsize;
  int sa = 16, da = 16;
  if (b.klass == KC_ROWS) {
    sa = da = pow2_divisor(t[0].e * S, 16);
    for (int i = 1; i < n; ++i) {
      sa = std::min(sa, pow2_divisor(t[i].ss * S, 16));
      da = std::min(da, pow2_divisor(t[i].ds * S, 16));
    }
  } else if (b.klass == KC_TRANSPOSE) {
    // whole 16-byte vectors along X on the source side, along Y on the destination side
    sa = std::min(sa, pow2_divisor(t[0].e * S, 16));
    sa = std::min(sa, pow2_divisor(t[1].ss * S, 16));
    da = std::min(da, pow2_divisor(t[1].e * S, 16));
    da = std::min(da, pow2_divisor(t[0].ds * S, 16));
    for (int i = 2; i < n; ++i) {
      sa = std::min(sa, pow2_divisor(t[i].ss * S, 16));
      da = std::min(da, pow2_divisor(t[i].ds * S, 16));
    }
  } else {
    sa = da = (int)std::min<i64>(S, 16);
  }
  b.src_align = sa;
  b.dst_align = da;
  b.stride_align = std::min(sa, da);
}

// Chunk `part` of `nparts` along the outermost raw dim with extent > 1.  The dense
// side of a pack (dst) / unpack (src) block lists its dims in the same order, so
// the chunk is one contiguous sub-range of it, and sender and receiver -- who see
// the same box -- cut it at the same places.
BlockCopy sub_block(const BlockCopy& b, int part, int nparts, bool dense_is_dst, i64* dense_off,
                    i64* dense_cnt) {
  BlockCopy c = b;
  int j = -1;
  for (int i = b.nd_raw - 1; i >= 0; --i)
    if (b.raw[i].e > 1) {
      j = i;
      break;
    }
  if (b.count == 0 || j < 0) {
    // nothing to cut: the whole block is part 0
    if (part != 0) {
      for (int i = 0; i < c.nd_raw; ++i) c.raw[i].e = (i == 0) ? 0 : c.raw[i].e;
      if (c.nd_raw == 0) {
        c.nd_raw = 1;
        c.raw[0] = Dim{0, 1, 1};
      }
    }
    canonicalize(c);
    if (dense_off) *dense_off = 0;
    if (dense_cnt) *dense_cnt = c.count;
    return c;
  }
  const i64 e = b.raw[j].e;
  const i64 c0 = e * part / nparts, c1 = e * (part + 1) / nparts;
  c.raw[j].e = c1 - c0;
  c.src_off += c0 * b.raw[j].ss;
  c.dst_off += c0 * b.raw[j].ds;
  canonicalize(c);
  const i64 dense_stride = dense_is_dst ? b.raw[j].ds : b.raw[j].ss;
  if (dense_off) *dense_off = c0 * dense_stride;
  if (dense_cnt) *dense_cnt = (c1 - c0) * dense_stride;
  return c;
}

// ---- plan -------------------------------------------------------------------
namespace {

struct LocalLayout {
  i64 lo[PA_MAX_DIMS], hi[PA_MAX_DIMS], len[PA_MAX_DIMS];
  i64 stride[PA_MAX_DIMS];  // element stride of LOGICAL dim l in the parent (memory-order) array
  i64 total;                // prod(len)
};

// parent(A) has dims (perm * size_local) in memory order (Pencils.jl:229;
// arrays.jl:108-114): memory dim m holds logical dim perm[m].
LocalLayout layout_of(const Pencil& p, const i64* coords = nullptr) {
  LocalLayout L;
  p.range_of(coords ? coords : p.topo->coords, L.lo, L.hi);
  L.total = 1;
  for (int d = 0; d < p.N; ++d) L.len[d] = L.hi[d] - L.lo[d];
  i64 run = 1;
  for (int m = 0; m < p.N; ++m) {
    int l = p.perm[m];
    L.stride[l] = run;
    run *= L.len[l];
  }
  L.total = run;
  return L;
}

}  // namespace

pa_status build_plan(std::shared_ptr<Pencil> pin, std::shared_ptr<Pencil> pout, int n_extra,
                     const i64* extra, int elsize, int method, Plan** out) {
  const Pencil& Pi = *pin;
  const Pencil& Po = *pout;
  // assert_compatible (Transpositions.jl:181-198)
  if (!(Pi.topo.get() == Po.topo.get() || Pi.topo->same_as(*Po.topo))) {
    set_error("pencil topologies must be the same.");
    return PA_EINCOMPAT;
  }
  if (Pi.N != Po.N) {
    set_error("pencils have different dimensionality: %d != %d", Pi.N, Po.N);
    return PA_EINCOMPAT;
  }
  for (int d = 0; d < Pi.N; ++d)
    if (Pi.size_global[d] != Po.size_global[d]) {
      set_error("global data sizes must be the same between different pencil configurations.");
      return PA_EINCOMPAT;
    }
  const int M = Pi.topo->M;
  int ndiff = 0, dim = -1;
  for (int i = 0; i < M; ++i)
    if (Pi.decomp[i] != Po.decomp[i]) {
      if (dim < 0) dim = i;  // findfirst (Transpositions.jl:110)
      ++ndiff;
    }
  if (ndiff > 1) {
    set_error("pencil decompositions must differ in at most one dimension.");
    return PA_EINCOMPAT;
  }
  if (n_extra < 0 || Pi.N + n_extra > PA_MAX_DIMS) {
    set_error("too many dimensions: N + n_extra must be <= %d", PA_MAX_DIMS);
    return PA_EINVAL;
  }
  if (elsize <= 0) {
    set_error("invalid element size %d", elsize);
    return PA_EINVAL;
  }
  {
    // an element size that is not a power of two <= 16 costs one more (innermost) dimension
    int w = 1;
    while (w < 16 && elsize % (2 * w) == 0) w *= 2;
    if (elsize / w > 1 && Pi.N + n_extra + 1 > PA_MAX_DIMS) {
      set_error("too many dimensions for %d-byte elements: N + n_extra must be <= %d", elsize,
                PA_MAX_DIMS - 1);
      return PA_EINVAL;
    }
  }
  if (method < PA_POINT_TO_POINT || method > PA_PEER_GET) {
    set_error("unknown transposition method %d", method);
    return PA_EINVAL;
  }

  std::unique_ptr<Plan> P(new Plan);
  P->pin = pin;
  P->pout = pout;
  P->n_extra = n_extra;
  P->prod_extra = 1;
  for (int j = 0; j < n_extra; ++j) {
    if (extra[j] < 0) {
      set_error("negative extra dimension");
      return PA_EINVAL;
    }
    P->extra[j] = extra[j];
    P->prod_extra *= extra[j];
  }
  // Elements whose size is not a power of two <= 16 move as several
  // power-of-two words: an innermost pseudo-dimension of `sub` words.
  int word = elsize, sub = 1;
  {
    int w = 1;
    while (w < 16 && elsize % (2 * w) == 0) w *= 2;
    word = w;
    sub = elsize / w;
  }
  P->elsize = elsize;
  P->method = method;
  P->dim = dim;
  const int N = Pi.N;
  const LocalLayout Li = layout_of(Pi);
  const LocalLayout Lo = layout_of(Po);
  P->length_in = Li.total * P->prod_extra;
  P->length_out = Lo.total * P->prod_extra;
  P->same_perm = true;
  for (int m = 0; m < N; ++m)
    if (Pi.perm[m] != Po.perm[m]) P->same_perm = false;

  // helper: append one block's dims in SOURCE memory order.
  // `src_contig` / `dst_contig`: that side is a dense buffer whose dims are the
  // box extents in Pi memory order (wire layout, Transpositions.jl:552-565).
  const LocalLayout* Ldst = &Lo;  // destination parent layout used by make_block (a peer's for K1-put)
  const LocalLayout* Lsrc = &Li;  // source parent layout (a peer's for K2-get)
  auto make_block = [&](const i64* blo, const i64* bhi, bool src_contig, bool dst_contig,
                        i64 src_base, i64 dst_base) {
    BlockCopy b;
    b.elsize = word;
    int k = 0;
    if (sub > 1) b.raw[k++] = Dim{(i64)sub, 1, 1};
    i64 run = sub;  // running product for the contiguous side(s), in words
    i64 soff = 0, doff = 0;
    for (int m = 0; m < N; ++m) {
      int l = Pi.perm[m];
      i64 e = std::max<i64>(0, bhi[l] - blo[l]);
      i64 ss = src_contig ? run : Lsrc->stride[l] * sub;
      i64 ds = dst_contig ? run : Ldst->stride[l] * sub;
      if (!src_contig) soff += (blo[l] - Lsrc->lo[l]) * Lsrc->stride[l] * sub;
      if (!dst_contig) doff += (blo[l] - Ldst->lo[l]) * Ldst->stride[l] * sub;
      b.raw[k++] = Dim{e, ss, ds};
      run *= e;
    }
    i64 es = Lsrc->total * sub, ed = Ldst->total * sub;  // strides of the first extra dim in a parent
    for (int j = 0; j < n_extra; ++j) {
      i64 e = P->extra[j];
      b.raw[k++] = Dim{e, src_contig ? run : es, dst_contig ? run : ed};
      run *= e;
      es *= e;
      ed *= e;
    }
    b.nd_raw = k;
    b.src_off = src_base * sub + soff;
    b.dst_off = dst_base * sub + doff;
    canonicalize(b);
    return b;
  };

  if (dim < 0) {
    // transpose_impl!(::Nothing): same decomposition, copy or local permute
    // (Transpositions.jl:213-233).  size_local(Ai) === size_local(Ao) holds.
    P->nproc = 1;
    P->self_index = 0;
    P->length_self = P->length_in;
    P->self_fused = make_block(Li.lo, Li.hi, false, false, 0, 0);
    P->send_elems = 0;
    P->recv_elems = P->length_out;
    *out = P.release();
    return PA_OK;
  }

  const Topology& T = *Pi.topo;
  const int nproc = (int)T.dims[dim];
  P->nproc = nproc;
  P->self_index = (int)T.coords[dim];
  P->peers.resize(nproc);

  // length_self (Transpositions.jl:302-305)
  {
    i64 n = P->prod_extra;
    for (int d = 0; d < N; ++d)
      n *= std::max<i64>(0, std::min(Li.hi[d], Lo.hi[d]) - std::max(Li.lo[d], Lo.lo[d]));
    P->length_self = n;
  }
  P->send_elems = P->length_in - P->length_self;
  P->recv_elems = P->length_out;
  const i64 length_recv = P->length_out - P->length_self;  // data from other processes (:372)

  i64 isend = 0, irecv = 0;
  i64 coords[PA_MAX_TOPO];
  for (int i = 0; i < T.M; ++i) coords[i] = T.coords[i];
  for (int n = 0; n < nproc; ++n) {
    Peer& pr = P->peers[n];
    coords[dim] = n;  // get_remote_indices (:539-549)
    pr.world_rank = T.rank_of(coords);
    pr.is_self = (n == P->self_index);
    i64 olo[PA_MAX_DIMS], ohi[PA_MAX_DIMS], ilo[PA_MAX_DIMS], ihi[PA_MAX_DIMS];
    Po.range_of(coords, olo, ohi);
    Pi.range_of(coords, ilo, ihi);
    // srange = Pi.axes_local ∩ Po.axes_all[n]   (:382)
    i64 slo[PA_MAX_DIMS], shi[PA_MAX_DIMS], rlo[PA_MAX_DIMS], rhi[PA_MAX_DIMS];
    i64 sc = P->prod_extra, rc = P->prod_extra;
    for (int d = 0; d < N; ++d) {
      slo[d] = std::max(Li.lo[d], olo[d]);
      shi[d] = std::max(slo[d], std::min(Li.hi[d], ohi[d]));
      sc *= shi[d] - slo[d];
      // rrange = Po.axes_local ∩ Pi.axes_all[n]  (:387)
      rlo[d] = std::max(Lo.lo[d], ilo[d]);
      rhi[d] = std::max(rlo[d], std::min(Lo.hi[d], ihi[d]));
      rc *= rhi[d] - rlo[d];
    }
    pr.send_cnt = sc;
    pr.recv_cnt = rc;
    if (pr.is_self) {
      // self block: packed straight into the tail of recv_buf (:393-403)
      pr.send_off = 0;
      pr.recv_off = length_recv;
      pr.pack = make_block(slo, shi, false, true, 0, pr.recv_off);
    } else {
      pr.send_off = isend;
      pr.recv_off = irecv;
      pr.pack = make_block(slo, shi, false, true, 0, pr.send_off);
      isend += sc;
      irecv += rc;
    }
    pr.unpack = make_block(rlo, rhi, true, false, pr.recv_off, 0);
    if (pr.is_self) {
      P->self_fused = make_block(slo, shi, false, false, 0, 0);
    } else {
      // K1-put: my send box written straight into peer n's dest parent, which has
      // the PEER's local layout (uneven blocks: its extents differ from mine)
      const LocalLayout Lpeer = layout_of(Po, coords);
      Ldst = &Lpeer;
      pr.put = make_block(slo, shi, false, false, 0, 0);
      Ldst = &Lo;
      // K2-get: the block peer n holds for me, read straight out of ITS src parent
      // (its Pi layout) and stored permuted into my dest parent
      const LocalLayout Lpin = layout_of(Pi, coords);
      Lsrc = &Lpin;
      pr.get = make_block(rlo, rhi, false, false, 0, 0);
      Lsrc = &Li;
      // where my block lands inside peer n's recv_buf: behind the blocks of the
      // sources that precede me in the line (remote blocks in increasing source
      // index, :380-416 as seen from the receiver)
      i64 off = 0, c2[PA_MAX_TOPO];
      for (int i = 0; i < T.M; ++i) c2[i] = T.coords[i];
      for (int m = 0; m < P->self_index; ++m) {
        if (m == n) continue;
        c2[dim] = m;
        i64 mlo[PA_MAX_DIMS], mhi[PA_MAX_DIMS];
        Pi.range_of(c2, mlo, mhi);
        i64 cnt = P->prod_extra;
        for (int d = 0; d < N; ++d)
          cnt *= std::max<i64>(0, std::min(ohi[d], mhi[d]) - std::max(olo[d], mlo[d]));
        off += cnt;
      }
      pr.remote_recv_off = off;
    }
  }
  if (isend != P->send_elems || irecv != length_recv) {
    set_error("internal error: block sizes do not tile the local arrays (%lld/%lld, %lld/%lld)",
              (long long)isend, (long long)P->send_elems, (long long)irecv, (long long)length_recv);
    return PA_EINVAL;
  }
  *out = P.release();
  return PA_OK;
}

}  // namespace pa
