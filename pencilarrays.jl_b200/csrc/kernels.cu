// sm_100a kernels of the transpose! hot path.
//
// Every data movement of the path -- K1 pack (copy_range!,
// Transpositions.jl:552-583), K2 unpack+permute (copy_permuted! ->
// _permutedims!, :585-664), K3 fused self block / permute_local! (:235-270),
// and the one-sided K1-put / K2-get variants whose other side is a peer GPU's
// memory over NVLink -- is one primitive: an N-d strided box copy
//     dst[sum k_i ds_i] = src[sum k_i ss_i],  k in box,
// canonicalised on the host (plan.cpp) into tile dims X (source-fastest),
// Y (destination-fastest, or the next dim) and up to 6 outer dims.
//
// Three kernels, all bandwidth-bound pure byte movers (bit-exact by construction):
//   k_rows<VT>          X contiguous on both sides: vectorised row copy, 8 x
//                       128-bit loads in flight per thread, streaming hints.
//   k_transpose_vec<S>  X contiguous in src, Y contiguous in dst: 512-byte
//                       coalesced 128-bit loads, VxV register micro-transpose,
//                       padded (conflict-free) shared tile in 16-byte items,
//                       512-byte coalesced 128-bit stores.
//   k_tile_scalar<ET>   any strides / alignment (odd sizes, tiny boxes):
//                       32x32 element tile through padded shared memory.
// Each exists in two flavours: LOOP=false, one tile per CTA (HBM-bound local
// work: the grid is the whole problem), and LOOP=true, a capped grid whose CTAs
// stride over the tiles (NVLink-bound remote work: a few CTAs per SM saturate
// the links and the rest of the SM stays available to concurrent local kernels).
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

#include "pa_internal.hpp"

namespace pa {

Tunables g_tun;
static std::atomic<i64> g_launches{0};
i64 launch_count() { return g_launches.load(); }

constexpr int MAXO = PA_MAX_DIMS - 2;

struct KParams {
  const char* src;  // offset to the box origin
  char* dst;
  long long ex, ey;                  // tile-dim extents (k_rows: ex in vectors)
  long long sx_s, sx_d, sy_s, sy_d;  // byte strides of X and Y
  unsigned tiles_x, tiles_y;
  unsigned long long total;  // tiles in the launch
  int no;                    // outer dims
  long long oe[MAXO], os[MAXO], od[MAXO];
  int lx_log2, ux_log2;  // k_rows thread/unroll shape
};

__device__ __forceinline__ void decode_tile(const KParams& p, unsigned long long bid, unsigned& tx,
                                            unsigned& ty, const char*& s, char*& d) {
  tx = (unsigned)(bid % p.tiles_x);
  bid /= p.tiles_x;
  ty = (unsigned)(bid % p.tiles_y);
  bid /= p.tiles_y;
  s = p.src;
  d = p.dst;
#pragma unroll 1
  for (int i = 0; i < p.no; ++i) {
    long long k = (long long)(bid % (unsigned long long)p.oe[i]);
    bid /= (unsigned long long)p.oe[i];
    s += k * p.os[i];
    d += k * p.od[i];
  }
}

// streaming (evict-first) accesses: every byte is touched exactly once
template <typename T>
__device__ __forceinline__ T ld_stream(const char* p) {
  return __ldcs(reinterpret_cast<const T*>(p));
}
template <typename T>
__device__ __forceinline__ void st_stream(char* p, const T& v) {
  __stcs(reinterpret_cast<T*>(p), v);
}

// ---------------------------------------------------------------------------
// K_rows: runs contiguous on both sides.
// 256 threads as LX x LY, each thread moves 8 vectors laid out UX x UY.
template <typename VT>
__device__ __forceinline__ void rows_tile(const KParams& p, unsigned long long bid) {
  constexpr int W = sizeof(VT);
  constexpr int U = 8;
  const int lxl = p.lx_log2, uxl = p.ux_log2;
  const int LX = 1 << lxl, LY = 256 >> lxl;
  const int lx = threadIdx.x & (LX - 1), ly = threadIdx.x >> lxl;
  unsigned tx, ty;
  const char* s;
  char* d;
  decode_tile(p, bid, tx, ty, s, d);
  const long long xv0 = (long long)tx * ((long long)LX << uxl) + lx;
  const long long y0 = (long long)ty * ((long long)LY << (3 - uxl)) + ly;
  VT v[U];
  long long so[U], dof[U];
  bool ok[U];
#pragma unroll
  for (int i = 0; i < U; ++i) {
    const int ux = i & ((1 << uxl) - 1), uy = i >> uxl;
    const long long xv = xv0 + (long long)ux * LX, y = y0 + (long long)uy * LY;
    ok[i] = (xv < p.ex) && (y < p.ey);
    so[i] = y * p.sy_s + xv * W;
    dof[i] = y * p.sy_d + xv * W;
  }
#pragma unroll
  for (int i = 0; i < U; ++i)
    if (ok[i]) v[i] = ld_stream<VT>(s + so[i]);
#pragma unroll
  for (int i = 0; i < U; ++i)
    if (ok[i]) st_stream<VT>(d + dof[i], v[i]);
}

template <typename VT, bool LOOP>
__global__ void __launch_bounds__(256) k_rows(const __grid_constant__ KParams p) {
  if constexpr (LOOP) {
    for (unsigned long long bid = blockIdx.x; bid < p.total; bid += gridDim.x) rows_tile<VT>(p, bid);
  } else {
    rows_tile<VT>(p, blockIdx.x);
  }
}

// ---------------------------------------------------------------------------
// K_rows_bulk: the row copy as a pure DMA pipeline (TMA bulk copies, SASS
// UBLKCP).  One elected thread per CTA drives a ring of STAGES shared-memory
// buffers: cp.async.bulk global->shared completing on an mbarrier, then
// cp.async.bulk shared->global in a bulk group; the CTA strides over
// (row, chunk) units.  No registers or LSU slots are spent on the payload, and
// the stores leave the SM as whole bulk transactions -- which is what the
// NVLink-bound put path wants.
constexpr int BULK_STAGES = 4;
constexpr int BULK_CHUNK = 16384;  // bytes per stage

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}

__global__ void __launch_bounds__(32) k_rows_bulk(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) char bulk_smem[];
  __shared__ __align__(8) unsigned long long full[BULK_STAGES];
  if (threadIdx.x != 0) return;  // a single thread owns the whole pipeline
  const unsigned chunk = (unsigned)p.lx_log2;  // bytes per unit (<= BULK_CHUNK), multiple of 16
  const long long run = p.ex;                  // bytes per row
  for (int s = 0; s < BULK_STAGES; ++s)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  auto unit = [&](unsigned long long bid, const char*& s, char*& d, unsigned& bytes) {
    unsigned tx, ty;
    decode_tile(p, bid, tx, ty, s, d);
    const long long x = (long long)tx * chunk;
    s += (long long)ty * p.sy_s + x;
    d += (long long)ty * p.sy_d + x;
    bytes = (unsigned)((run - x) < (long long)chunk ? (run - x) : (long long)chunk);
  };
  auto load = [&](int st, unsigned long long bid) {
    const char* s;
    char* d;
    unsigned bytes;
    unit(bid, s, d, bytes);
    const unsigned bar = smem_u32(&full[st]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(bulk_smem + st * BULK_CHUNK)),
        "l"(s), "r"(bytes), "r"(bar)
        : "memory");
  };

  unsigned long long next_load = blockIdx.x;
  int ls = 0;
  for (int i = 0; i < BULK_STAGES && next_load < p.total; ++i) {
    load(ls, next_load);
    next_load += gridDim.x;
    ls = (ls + 1) % BULK_STAGES;
  }
  int ss = 0, prev = -1;
  unsigned parity = 0;
  for (unsigned long long bid = blockIdx.x; bid < p.total; bid += gridDim.x) {
    const unsigned bar = smem_u32(&full[ss]);
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, q;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
    }
    const char* s;
    char* d;
    unsigned bytes;
    unit(bid, s, d, bytes);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d),
                 "r"(smem_u32(bulk_smem + ss * BULK_CHUNK)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (prev >= 0 && next_load < p.total) {
      // the PREVIOUS stage may be refilled once its store has finished READING it
      // (all but the newest bulk group done reading)
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      load(prev, next_load);
      next_load += gridDim.x;
    }
    prev = ss;
    ss = (ss + 1) % BULK_STAGES;
    if (ss == 0) parity ^= 1;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------
// K_transpose_vec: element size S in {4,8,16}, V = 16/S elements per vector.
// Tile: TA = 32*V elements along X (512 B of source row), TB = TBQ*V elements
// along Y (TBQ*16 B of destination row).
template <int S>
__device__ __forceinline__ uint4 gather_col(const uint4 (&r)[16 / S], int c);
template <>
__device__ __forceinline__ uint4 gather_col<16>(const uint4 (&r)[1], int) {
  return r[0];
}
template <>
__device__ __forceinline__ uint4 gather_col<8>(const uint4 (&r)[2], int c) {
  return c == 0 ? make_uint4(r[0].x, r[0].y, r[1].x, r[1].y)
                : make_uint4(r[0].z, r[0].w, r[1].z, r[1].w);
}
__device__ __forceinline__ unsigned comp(const uint4& v, int c) {
  return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w;
}
template <>
__device__ __forceinline__ uint4 gather_col<4>(const uint4 (&r)[4], int c) {
  return make_uint4(comp(r[0], c), comp(r[1], c), comp(r[2], c), comp(r[3], c));
}

template <int S, int TBQ>
__device__ __forceinline__ void transpose_tile(const KParams& p, unsigned long long bid, uint4* sm) {
  constexpr int V = 16 / S;
  constexpr int TA = 32 * V;
  constexpr int TB = TBQ * V;
  constexpr int PITCH = TBQ + 1;  // odd pitch in 16-byte items: conflict-free both phases
  constexpr int QI = TBQ / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned tx, ty;
  const char* s;
  char* d;
  decode_tile(p, bid, tx, ty, s, d);
  const long long x0 = (long long)tx * TA, y0 = (long long)ty * TB;
  const long long xl = x0 + lane * V;
  const bool xok = xl < p.ex;

  uint4 r[QI][V];
#pragma unroll
  for (int qi = 0; qi < QI; ++qi) {
#pragma unroll
    for (int rr = 0; rr < V; ++rr) {
      const long long y = y0 + (long long)(warp + 8 * qi) * V + rr;
      r[qi][rr] = (xok && y < p.ey) ? ld_stream<uint4>(s + y * p.sy_s + xl * S)
                                    : make_uint4(0u, 0u, 0u, 0u);
    }
  }
#pragma unroll
  for (int qi = 0; qi < QI; ++qi) {
    const int q = warp + 8 * qi;
#pragma unroll
    for (int c = 0; c < V; ++c) sm[(c * 32 + lane) * PITCH + q] = gather_col<S>(r[qi], c);
  }
  __syncthreads();
  constexpr int IT = TA * TBQ / 256;
#pragma unroll
  for (int k = 0; k < IT; ++k) {
    const int idx = threadIdx.x + 256 * k;
    const int xr = idx / TBQ, qq = idx % TBQ;
    const long long x = x0 + (long long)(xr & 31) * V + (xr >> 5);
    const long long y = y0 + (long long)qq * V;
    if (x < p.ex && y < p.ey) st_stream<uint4>(d + x * p.sx_d + y * S, sm[xr * PITCH + qq]);
  }
}

template <int S, int TBQ, bool LOOP>
__global__ void __launch_bounds__(256) k_transpose_vec(const __grid_constant__ KParams p) {
  __shared__ uint4 sm[(32 * 16 / S) * (TBQ + 1)];
  if constexpr (LOOP) {
    for (unsigned long long bid = blockIdx.x; bid < p.total; bid += gridDim.x) {
      transpose_tile<S, TBQ>(p, bid, sm);
      __syncthreads();  // the shared tile is reused by the next iteration
    }
  } else {
    transpose_tile<S, TBQ>(p, blockIdx.x, sm);
  }
}

// ---------------------------------------------------------------------------
// K_tile_scalar: general strides, element-wise accesses.
template <typename ET>
__device__ __forceinline__ void scalar_tile(const KParams& p, unsigned long long bid, ET (*sm)[33]) {
  const int a = threadIdx.x & 31, b = threadIdx.x >> 5;
  unsigned tx, ty;
  const char* s;
  char* d;
  decode_tile(p, bid, tx, ty, s, d);
  const long long x0 = (long long)tx * 32, y0 = (long long)ty * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long x = x0 + a, y = y0 + b + 8 * k;
    if (x < p.ex && y < p.ey)
      sm[b + 8 * k][a] = *reinterpret_cast<const ET*>(s + x * p.sx_s + y * p.sy_s);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long x = x0 + b + 8 * k, y = y0 + a;
    if (x < p.ex && y < p.ey)
      *reinterpret_cast<ET*>(d + x * p.sx_d + y * p.sy_d) = sm[a][b + 8 * k];
  }
}

template <typename ET, bool LOOP>
__global__ void __launch_bounds__(256) k_tile_scalar(const __grid_constant__ KParams p) {
  __shared__ ET sm[32][33];
  if constexpr (LOOP) {
    for (unsigned long long bid = blockIdx.x; bid < p.total; bid += gridDim.x) {
      scalar_tile<ET>(p, bid, sm);
      __syncthreads();
    }
  } else {
    scalar_tile<ET>(p, blockIdx.x, sm);
  }
}

// ---------------------------------------------------------------------------
static int pow2_of_ptr(const void* a, const void* b) {
  uintptr_t v = (uintptr_t)a | (uintptr_t)b;
  int al = 1;
  while (al < 16 && (v % (2 * al)) == 0) al *= 2;
  return al;
}
static int ceil_log2(long long x) {
  int l = 0;
  while ((1LL << l) < x) ++l;
  return l;
}
static long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

static int sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

// max_ctas: 0 = one tile per CTA; > 0 = capped, tile-striding grid;
//           < 0 = -max_ctas CTAs per SM (resolved against the device here)
template <typename K1, typename KL>
static pa_status do_launch(K1 kern_one, KL kern_loop, KParams& p, cudaStream_t st, int max_ctas) {
  unsigned long long tiles = (unsigned long long)p.tiles_x * p.tiles_y;
  for (int i = 0; i < p.no; ++i) tiles *= (unsigned long long)p.oe[i];
  if (tiles == 0) return PA_OK;
  if (tiles > 0x7fffffffULL) {
    set_error("block too large for one launch (%llu tiles)", tiles);
    return PA_EINVAL;
  }
  p.total = tiles;
  if (max_ctas < 0) max_ctas = -max_ctas * sm_count();
  if (max_ctas > 0 && tiles > (unsigned long long)max_ctas)
    kern_loop<<<(unsigned)max_ctas, 256, 0, st>>>(p);
  else
    kern_one<<<(unsigned)tiles, 256, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("kernel launch failed: %s", cudaGetErrorString(e));
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? PA_ENOGPU : PA_ECUDA;
  }
  g_launches.fetch_add(1);
  return PA_OK;
}

#define LAUNCH(K, ...) do_launch(K<__VA_ARGS__, false>, K<__VA_ARGS__, true>, p, st, max_ctas)

pa_status launch_block(const BlockCopy& b, const void* src, void* dst, void* stream,
                       int* vec_used, int max_ctas) {
  if (vec_used) *vec_used = 0;
  if (b.klass == KC_EMPTY) return PA_OK;
  if (!src || !dst) {
    set_error("null array pointer");
    return PA_EINVAL;
  }
  const long long S = b.elsize;
  const char* s = (const char*)src + b.src_off * S;
  char* d = (char*)dst + b.dst_off * S;
  const int pal = pow2_of_ptr(s, d);
  if (pal < (S > 16 ? 16 : S)) {
    set_error("array pointers must be aligned to the element word size (%lld)", S);
    return PA_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  KParams p;
  memset(&p, 0, sizeof p);
  p.src = s;
  p.dst = d;
  const Dim X = b.d[0];
  const Dim Y = b.nd > 1 ? b.d[1] : Dim{1, 0, 0};
  p.ex = X.e;
  p.ey = Y.e;
  p.sx_s = X.ss * S;
  p.sx_d = X.ds * S;
  p.sy_s = Y.ss * S;
  p.sy_d = Y.ds * S;
  p.no = b.nd > 2 ? b.nd - 2 : 0;
  for (int i = 0; i < p.no; ++i) {
    p.oe[i] = b.d[i + 2].e;
    p.os[i] = b.d[i + 2].ss * S;
    p.od[i] = b.d[i + 2].ds * S;
  }

  if (b.klass == KC_ROWS && g_tun.bulk_rows && std::min(b.stride_align, pal) == 16) {
    // TMA bulk-copy pipeline (tunable "bulk_rows")
    static bool attr_set[64] = {false};  // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      cudaFuncSetAttribute(k_rows_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           BULK_STAGES * BULK_CHUNK);
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const long long run = X.e * S;
    const long long chunk = std::min<long long>(BULK_CHUNK, run);
    p.ex = run;
    p.lx_log2 = (int)chunk;
    p.tiles_x = (unsigned)cdiv(run, chunk);
    p.tiles_y = (unsigned)Y.e;
    unsigned long long units = (unsigned long long)p.tiles_x * p.tiles_y;
    for (int i = 0; i < p.no; ++i) units *= (unsigned long long)p.oe[i];
    p.total = units;
    long long ctas = max_ctas > 0 ? max_ctas : (max_ctas < 0 ? -max_ctas : 3) * (long long)sm_count();
    if ((unsigned long long)ctas > units) ctas = (long long)units;
    if (vec_used) *vec_used = 16;
    k_rows_bulk<<<(unsigned)ctas, 32, BULK_STAGES * BULK_CHUNK, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("kernel launch failed: %s", cudaGetErrorString(e));
      return PA_ECUDA;
    }
    g_launches.fetch_add(1);
    return PA_OK;
  }

  if (b.klass == KC_ROWS) {
    const int W = std::min(b.stride_align, pal);
    const long long exv = X.e * S / W;
    p.ex = exv;
    int lxl = std::min(8, ceil_log2(exv));
    const long long LX = 1LL << lxl, LY = 256 >> lxl;
    int uxl = std::min(3, ceil_log2(cdiv(exv, LX)));
    p.lx_log2 = lxl;
    p.ux_log2 = uxl;
    p.tiles_x = (unsigned)cdiv(exv, LX << uxl);
    p.tiles_y = (unsigned)cdiv(Y.e, LY << (3 - uxl));
    if (vec_used) *vec_used = W;
    switch (W) {
      case 16: return LAUNCH(k_rows, uint4);
      case 8: return LAUNCH(k_rows, uint2);
      case 4: return LAUNCH(k_rows, uint32_t);
      case 2: return LAUNCH(k_rows, uint16_t);
      default: return LAUNCH(k_rows, uint8_t);
    }
  }

  if (b.klass == KC_TRANSPOSE && std::min(b.stride_align, pal) == 16 &&
      (S == 4 || S == 8 || S == 16)) {
    if (vec_used) *vec_used = 16;
    // destination-run length TBQ*16 B: 512 B by default; 256 B for small blocks
    // (more, smaller tiles -> shorter tail), tunable "transpose_tbq" overrides
    int tbq = g_tun.transpose_tbq;
    if (tbq == 0) tbq = (b.count * S < g_tun.small_block_bytes) ? 16 : 32;
    if (S == 16) {
      p.tiles_x = (unsigned)cdiv(X.e, 32);
      // (a 1-KiB-run tile, TBQ = 64, measured 14 % slower: profiles/r1_analysis.md; not kept)
      p.tiles_y = (unsigned)cdiv(Y.e, tbq == 16 ? 16 : 32);
      if (tbq == 16) return LAUNCH(k_transpose_vec, 16, 16);
      return LAUNCH(k_transpose_vec, 16, 32);
    } else if (S == 8) {
      p.tiles_x = (unsigned)cdiv(X.e, 64);
      p.tiles_y = (unsigned)cdiv(Y.e, tbq == 16 ? 32 : 64);
      if (tbq == 16) return LAUNCH(k_transpose_vec, 8, 16);
      return LAUNCH(k_transpose_vec, 8, 32);
    } else {
      p.tiles_x = (unsigned)cdiv(X.e, 128);
      p.tiles_y = (unsigned)cdiv(Y.e, 64);
      return LAUNCH(k_transpose_vec, 4, 16);
    }
  }

  // general path
  p.tiles_x = (unsigned)cdiv(X.e, 32);
  p.tiles_y = (unsigned)cdiv(Y.e, 32);
  if (vec_used) *vec_used = (int)S;
  switch (S) {
    case 16: return LAUNCH(k_tile_scalar, uint4);
    case 8: return LAUNCH(k_tile_scalar, uint2);
    case 4: return LAUNCH(k_tile_scalar, uint32_t);
    case 2: return LAUNCH(k_tile_scalar, uint16_t);
    case 1: return LAUNCH(k_tile_scalar, uint8_t);
    default:
      set_error("unsupported element word size %lld", S);
      return PA_EINVAL;
  }
}

int device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

}  // namespace pa
