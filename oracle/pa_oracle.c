/*
 * CPU ORACLE / CPU BASELINE (test infrastructure -- never linked into, loaded
 * by or called from the product path).
 *
 * Plain-C restatement of the reference's `transpose!` for ALL ranks of a
 * process grid inside one process: one worker thread per emulated MPI rank
 * (extra threads split a rank's loops), the exchange is a memcpy between the
 * ranks' buffers standing in for MPICH's shared-memory transport.  Used by
 *   - tests/test_oracle.py (checked against the NumPy oracle, bit-exact),
 *   - bench.py `cpu_baseline` and `--impl reference` (timed on host cores).
 * Parity status: pinned by the reference's test properties only (Julia + MPI
 * are not available here, see oracle/pencil_oracle.py) -- kind = "port".
 *
 * Follows, function by function:
 *   local_data_range           src/Pencils/data_ranges.jl:4-9
 *   axes of a rank             data_ranges.jl:15-45, MPITopologies.jl:125-131 (row-major ranks)
 *   to_local / memory order    src/Pencils/Pencils.jl:229,579-587
 *   transpose_send! + offsets  src/Transpositions/Transpositions.jl:345-430
 *   copy_range!  (pack)        Transpositions.jl:552-565
 *   transpose_recv!            Transpositions.jl:486-533
 *   copy_permuted! (unpack)    Transpositions.jl:585-645 (Strided.jl blocked copy restated
 *                              as a 2-d cache-tiled loop nest)
 *   permute_local!             Transpositions.jl:213-270
 * Indices are 1-based inclusive ranges [a, b] like the Julia source.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 8
typedef int64_t i64;

typedef struct { i64 a, b; } rng_t; /* a:b, empty when b < a */
static inline i64 rlen(rng_t r) { return r.b >= r.a ? r.b - r.a + 1 : 0; }
static inline rng_t isect(rng_t x, rng_t y) {
  rng_t r = { x.a > y.a ? x.a : y.a, x.b < y.b ? x.b : y.b };
  if (r.b < r.a) r.b = r.a - 1;
  return r;
}

/* data_ranges.jl:4-9 */
static rng_t local_data_range(i64 p, i64 P, i64 N) {
  rng_t r = { (N * (p - 1)) / P + 1, (N * p) / P };
  return r;
}

typedef struct {
  int M, N;
  i64 pdims[MAXD], size_global[MAXD];
  int decomp[MAXD]; /* 1-based array dim per grid dim */
  int perm[MAXD];   /* 1-based; memory dim i holds logical dim perm[i] */
} pencil_t;

/* MPI_Cart_create(reorder=false): row-major, last coordinate fastest; 1-based coords */
static void coords_of(const pencil_t* p, int rank, i64* c) {
  for (int i = p->M - 1; i >= 0; --i) { c[i] = rank % p->pdims[i] + 1; rank /= (int)p->pdims[i]; }
}
static int rank_of(const pencil_t* p, const i64* c) {
  i64 r = 0;
  for (int i = 0; i < p->M; ++i) r = r * p->pdims[i] + (c[i] - 1);
  return (int)r;
}
/* axes_all[coords] (data_ranges.jl:30-45) */
static void axes_of(const pencil_t* p, const i64* coords, rng_t* ax) {
  for (int d = 0; d < p->N; ++d) { ax[d].a = 1; ax[d].b = p->size_global[d]; }
  for (int i = 0; i < p->M; ++i) {
    int d = p->decomp[i] - 1;
    ax[d] = local_data_range(coords[i], p->pdims[i], p->size_global[d]);
  }
}

/* Strided column-major box copy: dst[sum k_i ds_i] = src[sum k_i ss_i].
 * dims are given in SOURCE memory order; nd <= MAXD.  Tiled over the source-
 * fastest dim (0) and the destination-fastest dim. */
static void box_copy(int nd, const i64* e, const i64* ss, const i64* ds, int es, const char* src,
                     char* dst, int tid, int nth) {
  i64 total = 1;
  for (int i = 0; i < nd; ++i) total *= e[i];
  if (total == 0) return;
  int b = 0;
  for (int i = 1; i < nd; ++i)
    if (ds[i] < ds[b]) b = i;
  if (b == 0 || nd == 1) {
    /* dim 0 fastest on both sides: runs (memcpy when contiguous), odometer over the rest */
    i64 rows = total / e[0];
    i64 r0 = rows * tid / nth, r1 = rows * (tid + 1) / nth;
    for (i64 r = r0; r < r1; ++r) {
      i64 t = r, so = 0, dof = 0;
      for (int i = 1; i < nd; ++i) { i64 k = t % e[i]; t /= e[i]; so += k * ss[i]; dof += k * ds[i]; }
      if (ss[0] == 1 && ds[0] == 1) memcpy(dst + dof * es, src + so * es, (size_t)(e[0] * es));
      else for (i64 k = 0; k < e[0]; ++k) memcpy(dst + (dof + k * ds[0]) * es, src + (so + k * ss[0]) * es, es);
    }
    return;
  }
  /* transpose-like: tile (dim 0, dim b) */
  enum { T = 32 };
  i64 ta = (e[0] + T - 1) / T, tb = (e[b] + T - 1) / T;
  i64 outer = total / (e[0] * e[b]);
  i64 ntiles = ta * tb * outer;
  i64 q0 = ntiles * tid / nth, q1 = ntiles * (tid + 1) / nth;
  for (i64 q = q0; q < q1; ++q) {
    i64 t = q;
    i64 ia = t % ta; t /= ta;
    i64 ib = t % tb; t /= tb;
    i64 so = 0, dof = 0;
    for (int i = 1; i < nd; ++i) {
      if (i == b) continue;
      i64 k = t % e[i]; t /= e[i]; so += k * ss[i]; dof += k * ds[i];
    }
    i64 a0 = ia * T, a1 = a0 + T < e[0] ? a0 + T : e[0];
    i64 b0 = ib * T, b1 = b0 + T < e[b] ? b0 + T : e[b];
    for (i64 x = a0; x < a1; ++x) {
      const char* sp = src + (so + x * ss[0] + b0 * ss[b]) * es;
      char* dp = dst + (dof + x * ds[0] + b0 * ds[b]) * es;
      if (es == 8) for (i64 y = b0; y < b1; ++y, sp += ss[b] * 8, dp += ds[b] * 8) *(uint64_t*)dp = *(const uint64_t*)sp;
      else if (es == 4) for (i64 y = b0; y < b1; ++y, sp += ss[b] * 4, dp += ds[b] * 4) *(uint32_t*)dp = *(const uint32_t*)sp;
      else if (es == 16) for (i64 y = b0; y < b1; ++y, sp += ss[b] * 16, dp += ds[b] * 16) { ((uint64_t*)dp)[0] = ((const uint64_t*)sp)[0]; ((uint64_t*)dp)[1] = ((const uint64_t*)sp)[1]; }
      else for (i64 y = b0; y < b1; ++y, sp += ss[b] * es, dp += ds[b] * es) memcpy(dp, sp, es);
    }
  }
}

typedef struct {
  rng_t ax_in[MAXD], ax_out[MAXD];      /* axes_local of Pi / Po */
  i64 len_in[MAXD], len_out[MAXD];
  i64 str_in[MAXD], str_out[MAXD];      /* stride of LOGICAL dim d in the memory-order parent */
  i64 tot_in, tot_out;
} local_t;

static void local_of(const pencil_t* pi, const pencil_t* po, int rank, local_t* L) {
  i64 c[MAXD];
  coords_of(pi, rank, c);
  axes_of(pi, c, L->ax_in);
  axes_of(po, c, L->ax_out);
  i64 run = 1;
  for (int m = 0; m < pi->N; ++m) { int d = pi->perm[m] - 1; L->len_in[d] = rlen(L->ax_in[d]); L->str_in[d] = run; run *= L->len_in[d]; }
  L->tot_in = run;
  run = 1;
  for (int m = 0; m < po->N; ++m) { int d = po->perm[m] - 1; L->len_out[d] = rlen(L->ax_out[d]); L->str_out[d] = run; run *= L->len_out[d]; }
  L->tot_out = run;
}

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* transpose!(dest, src) for every rank of the grid.
 * src[r], dst[r]: parent arrays (memory order) of rank r.  send[r] / recv[r]:
 * staging buffers of at least send_elems / recv_elems elements (query with
 * pao_sizes).  phase_s[3] = seconds in pack, exchange, unpack (max over threads).
 * Returns 0, or -1 for incompatible pencils (ArgumentError). */
int pao_transpose(int M, const i64* pdims, int N, const i64* size_global, const int* decomp_in,
                  const int* perm_in, const int* decomp_out, const int* perm_out, int n_extra,
                  const i64* extra, int es, void** src, void** dst, void** send, void** recv,
                  int nthreads, double* phase_s) {
  pencil_t pi, po;
  memset(&pi, 0, sizeof pi);
  pi.M = M; pi.N = N;
  for (int i = 0; i < M; ++i) { pi.pdims[i] = pdims[i]; pi.decomp[i] = decomp_in[i]; }
  for (int d = 0; d < N; ++d) { pi.size_global[d] = size_global[d]; pi.perm[d] = perm_in ? perm_in[d] : d + 1; }
  po = pi;
  for (int i = 0; i < M; ++i) po.decomp[i] = decomp_out[i];
  for (int d = 0; d < N; ++d) po.perm[d] = perm_out ? perm_out[d] : d + 1;
  int R = -1, ndiff = 0; /* Transpositions.jl:110,192 */
  for (int i = 0; i < M; ++i) if (pi.decomp[i] != po.decomp[i]) { if (R < 0) R = i; ++ndiff; }
  if (ndiff > 1) return -1;
  i64 pe = 1;
  for (int j = 0; j < n_extra; ++j) pe *= extra[j];
  int nranks = 1;
  for (int i = 0; i < M; ++i) nranks *= (int)pdims[i];
  if (nthreads < 1) nthreads = 1;
  int tpr = nthreads / nranks; /* threads per rank */
  if (tpr < 1) tpr = 1;
  int nworkers = nthreads < nranks ? nthreads : nranks * tpr;
  double t_pack = 0, t_exch = 0, t_unpack = 0;

#pragma omp parallel num_threads(nworkers)
  {
#ifdef _OPENMP
    int w = omp_get_thread_num();
#else
    int w = 0;
#endif
    for (int phase = 0; phase < 3; ++phase) {
      double t0 = now();
      /* rank loop: worker w serves ranks r with (r % nranks_per_pass) pattern */
      for (int r = (nthreads < nranks ? w : w / tpr); r < nranks; r += (nthreads < nranks ? nworkers : nranks)) {
        int tid = nthreads < nranks ? 0 : w % tpr, nth = nthreads < nranks ? 1 : tpr;
        local_t L;
        local_of(&pi, &po, r, &L);
        const char* s = (const char*)src[r];
        char* d = (char*)dst[r];
        i64 e[MAXD], ss[MAXD], ds[MAXD];
        if (R < 0) { /* transpose_impl!(::Nothing): copy! or permute_local! (:213-270) */
          if (phase != 0) continue;
          int k = 0;
          for (int m = 0; m < N; ++m) { int l = pi.perm[m] - 1; e[k] = L.len_in[l]; ss[k] = L.str_in[l]; ds[k] = L.str_out[l]; ++k; }
          i64 xs = L.tot_in, xd = L.tot_out;
          for (int j = 0; j < n_extra; ++j) { e[k] = extra[j]; ss[k] = xs; ds[k] = xd; xs *= extra[j]; xd *= extra[j]; ++k; }
          box_copy(k, e, ss, ds, es, s, d, tid, nth);
          continue;
        }
        i64 c[MAXD];
        coords_of(&pi, r, c);
        int Nproc = (int)pdims[R];
        /* length_self (:302-305) */
        i64 length_self = pe;
        for (int dd = 0; dd < N; ++dd) length_self *= rlen(isect(L.ax_in[dd], L.ax_out[dd]));
        i64 length_recv = L.tot_out * pe - length_self; /* :372 */
        i64 isend = 0, irecv = 0;
        for (int n = 1; n <= Nproc; ++n) { /* enumerate(remote_inds) (:380) */
          i64 cn[MAXD];
          memcpy(cn, c, sizeof cn);
          cn[R] = n;
          rng_t oax[MAXD], iax[MAXD];
          axes_of(&po, cn, oax);
          axes_of(&pi, cn, iax);
          rng_t sr[MAXD], rr[MAXD];
          i64 ls = pe, lr = pe;
          for (int dd = 0; dd < N; ++dd) {
            sr[dd] = isect(L.ax_in[dd], oax[dd]);  /* srange (:382) */
            rr[dd] = isect(L.ax_out[dd], iax[dd]); /* rrange (:387) */
            ls *= rlen(sr[dd]); lr *= rlen(rr[dd]);
          }
          int peer = rank_of(&pi, cn);
          int self = (peer == r);
          i64 soff = self ? 0 : isend, roff = self ? length_recv : irecv; /* :389,398 */
          if (phase == 0) {
            /* copy_range! (:552-565): box of the memory-order parent -> contiguous */
            int k = 0; i64 run = 1, so = 0;
            for (int m = 0; m < N; ++m) { int l = pi.perm[m] - 1; e[k] = rlen(sr[l]); ss[k] = L.str_in[l]; ds[k] = run; run *= e[k]; so += (sr[l].a - L.ax_in[l].a) * L.str_in[l]; ++k; }
            i64 xs = L.tot_in;
            for (int j = 0; j < n_extra; ++j) { e[k] = extra[j]; ss[k] = xs; ds[k] = run; run *= extra[j]; xs *= extra[j]; ++k; }
            char* out = self ? (char*)recv[r] + roff * es : (char*)send[r] + soff * es;
            box_copy(k, e, ss, ds, es, s + so * es, out, tid, nth);
          } else if (phase == 1) {
            /* Isend/Irecv pair or Alltoallv (:418-427,462-476) as a memcpy INTO my recv_buf:
             * the block peer `peer` packed for me sits at its send offset for rank r. */
            if (!self && lr > 0 && tid == 0) {
              /* recompute the sender's offset: blocks of lower-numbered peers it packed before mine */
              local_t Lp;
              local_of(&pi, &po, peer, &Lp);
              i64 poff = 0;
              int mypos = (int)c[R];
              for (int q = 1; q < mypos; ++q) {
                if (q == n) continue; /* the sender skips itself (self block goes to its recv_buf) */
                i64 cq[MAXD];
                memcpy(cq, c, sizeof cq);
                cq[R] = q;
                rng_t oq[MAXD];
                axes_of(&po, cq, oq);
                i64 l = pe;
                for (int dd = 0; dd < N; ++dd) l *= rlen(isect(Lp.ax_in[dd], oq[dd]));
                poff += l;
              }
              memcpy((char*)recv[r] + roff * es, (const char*)send[peer] + poff * es, (size_t)(lr * es));
            }
          } else {
            /* copy_permuted! (:585-645): contiguous block (dims in Pi memory order) -> dest box */
            int k = 0; i64 run = 1, dof = 0;
            for (int m = 0; m < N; ++m) { int l = pi.perm[m] - 1; e[k] = rlen(rr[l]); ss[k] = run; run *= e[k]; ds[k] = L.str_out[l]; dof += (rr[l].a - L.ax_out[l].a) * L.str_out[l]; ++k; }
            i64 xd = L.tot_out;
            for (int j = 0; j < n_extra; ++j) { e[k] = extra[j]; ss[k] = run; run *= extra[j]; ds[k] = xd; xd *= extra[j]; ++k; }
            box_copy(k, e, ss, ds, es, (const char*)recv[r] + roff * es, d + dof * es, tid, nth);
          }
          if (!self) { isend += ls; irecv += lr; }
        }
      }
#pragma omp barrier
      double dt = now() - t0;
#pragma omp critical
      {
        if (phase == 0 && dt > t_pack) t_pack = dt;
        if (phase == 1 && dt > t_exch) t_exch = dt;
        if (phase == 2 && dt > t_unpack) t_unpack = dt;
      }
#pragma omp barrier
    }
  }
  if (phase_s) { phase_s[0] = t_pack; phase_s[1] = t_exch; phase_s[2] = t_unpack; }
  return 0;
}

/* element counts of rank r: parent in/out, send_buf, recv_buf */
int pao_sizes(int M, const i64* pdims, int N, const i64* size_global, const int* decomp_in,
              const int* decomp_out, int n_extra, const i64* extra, int rank, i64* out4) {
  pencil_t pi, po;
  memset(&pi, 0, sizeof pi);
  pi.M = M; pi.N = N;
  for (int i = 0; i < M; ++i) { pi.pdims[i] = pdims[i]; pi.decomp[i] = decomp_in[i]; }
  for (int d = 0; d < N; ++d) { pi.size_global[d] = size_global[d]; pi.perm[d] = d + 1; }
  po = pi;
  for (int i = 0; i < M; ++i) po.decomp[i] = decomp_out[i];
  local_t L;
  local_of(&pi, &po, rank, &L);
  i64 pe = 1;
  for (int j = 0; j < n_extra; ++j) pe *= extra[j];
  i64 self = pe;
  for (int d = 0; d < N; ++d) self *= rlen(isect(L.ax_in[d], L.ax_out[d]));
  out4[0] = L.tot_in * pe;
  out4[1] = L.tot_out * pe;
  out4[2] = L.tot_in * pe - self; /* length_send (:308) */
  out4[3] = L.tot_out * pe;       /* length_recv_total (:309) */
  return 0;
}

int pao_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
