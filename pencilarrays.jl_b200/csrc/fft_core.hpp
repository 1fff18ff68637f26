// Radix passes of the in-shared-memory 1-d complex FFT that the fused
// unpack+FFT kernel (fft.cu) runs on every line it has just transposed.
// the permutation of a pencil exists to make the FFT axis
// contiguous (/root/reference/docs/src/Pencils.md:210-214,
// docs/src/Transpositions.md:7-9) -- here the transform is applied before the
// line ever leaves the SM.
//
// Algorithm: decimation in frequency (Gentleman-Sande), mixed radix
// {2, 4, 8}, in place.  A pass of radix R over sub-length M takes, for every
// group g and offset j < M/R, the R points x[g*M + j + q*M/R], replaces them
// by their R-point DFT and multiplies output s by W_M^(j*s).  After all
// passes the element at position p = s1*L/R1 + s2*L/(R1*R2) + ... holds
// X[s1 + R1*(s2 + R2*(...))]: fft_position_of() inverts that for the final
// natural-order read-out.  Twiddles come from a table W_L^k = exp(-2*pi*i*k/L),
// k < L, computed on the host in extended precision; the backward transform
// conjugates them (unnormalised, like FFTW / PencilFFTs).
//
// Everything is __host__ __device__ so that tests/fft_host_check.cpp can run
// the very same index arithmetic and butterflies on the CPU against numpy.fft.
#pragma once

#if defined(__CUDACC__)
#define PA_HD __host__ __device__ __forceinline__
#else
#define PA_HD inline
#endif

namespace pa_fft {

struct cplx {
  double x, y;
};

PA_HD cplx cadd(cplx a, cplx b) { return cplx{a.x + b.x, a.y + b.y}; }
PA_HD cplx csub(cplx a, cplx b) { return cplx{a.x - b.x, a.y - b.y}; }
PA_HD cplx cmul(cplx a, cplx b) { return cplx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
// multiplication by -i (forward) / +i (backward): sign = -1 forward, +1 backward
PA_HD cplx mul_i(cplx a, int sign) { return sign < 0 ? cplx{a.y, -a.x} : cplx{-a.y, a.x}; }

// storage index of logical position i: one pad element after every 8, 64 and 512
// elements, so that accesses at every power-of-8 stride (the strides of the radix-8
// passes and of the digit-reversed read-out) fall into different 16-byte bank groups
PA_HD int pad_index(int i) { return i + (i >> 3) + (i >> 6) + (i >> 9); }
// elements per padded line, forced to 1 mod 8 (the C lines of a CTA are written
// column-wise by 8 consecutive threads)
PA_HD int padded_pitch(int L) {
  int n = pad_index(L - 1) + 1;
  while ((n & 7) != 1) ++n;
  return n;
}

constexpr int MAX_PASSES = 5;
struct Radices {
  int n;
  int r[MAX_PASSES];
};

// small radix first (its stride is the longest: conflict-free), then radix 8
PA_HD Radices radices_of(int logL) {
  Radices R;
  R.n = 0;
  int rem = logL % 3;
  if (rem == 1) R.r[R.n++] = 2;
  if (rem == 2) R.r[R.n++] = 4;
  for (int i = 0; i < logL / 3; ++i) R.r[R.n++] = 8;
  return R;
}

// same, with a given first radix (the fused kernel's in-register first pass)
PA_HD Radices radices_with_first(int logL, int logR1) {
  Radices R;
  R.n = 0;
  if (logR1 > 0) R.r[R.n++] = 1 << logR1;
  const Radices rest = radices_of(logL - logR1);
  for (int i = 0; i < rest.n; ++i) R.r[R.n++] = rest.r[i];
  return R;
}

// position (after all passes) of output frequency k
PA_HD int fft_position_of(int k, int L, const Radices& R) {
  int p = 0, stride = L;
  for (int i = 0; i < R.n; ++i) {
    stride /= R.r[i];
    p += (k % R.r[i]) * stride;
    k /= R.r[i];
  }
  return p;
}

// R-point DFT in registers, a[q] -> A[s] = sum_q a[q] * w^(q*s), w = exp(sign*2*pi*i/R)
PA_HD void dft2(cplx* a) {
  cplx t = a[0];
  a[0] = cadd(t, a[1]);
  a[1] = csub(t, a[1]);
}
PA_HD void dft4(cplx* a, int sign) {
  cplx t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
  cplx t2 = cadd(a[1], a[3]), t3 = mul_i(csub(a[1], a[3]), sign);
  a[0] = cadd(t0, t2);
  a[1] = cadd(t1, t3);
  a[2] = csub(t0, t2);
  a[3] = csub(t1, t3);
}
PA_HD void dft8(cplx* a, int sign) {
  const double h = 0.70710678118654752440;
  // even / odd 4-point transforms
  cplx e[4] = {a[0], a[2], a[4], a[6]};
  cplx o[4] = {a[1], a[3], a[5], a[7]};
  dft4(e, sign);
  dft4(o, sign);
  // o[s] *= w8^s, w8 = exp(sign*i*pi/4)
  cplx w1 = cplx{h, sign * h}, w3 = cplx{-h, sign * h};
  o[1] = cmul(o[1], w1);
  o[2] = mul_i(o[2], sign);
  o[3] = cmul(o[3], w3);
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int s = 0; s < 4; ++s) {
    a[s] = cadd(e[s], o[s]);
    a[s + 4] = csub(e[s], o[s]);
  }
}

// 16 = 4 x 4: a[4*q1 + q2] -> A[s1 + 4*s2]
PA_HD void dft16(cplx* a, int sign) {
  const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;  // cos, sin(pi/8)
  const double h = 0.70710678118654752440;
  cplx b[4][4];
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int q2 = 0; q2 < 4; ++q2) {
    cplx t[4] = {a[q2], a[4 + q2], a[8 + q2], a[12 + q2]};
    dft4(t, sign);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) b[q2][k] = t[k];
  }
  // b[q2][s1] *= w16^(q2*s1), w16 = exp(sign*i*pi/8)
  const cplx w1 = cplx{c1, sign * s1}, w2 = cplx{h, sign * h}, w3 = cplx{s1, sign * c1};
  const cplx w6 = cplx{-h, sign * h}, w9 = cplx{-c1, -sign * s1};
  b[1][1] = cmul(b[1][1], w1);
  b[1][2] = cmul(b[1][2], w2);
  b[1][3] = cmul(b[1][3], w3);
  b[2][1] = cmul(b[2][1], w2);
  b[2][2] = mul_i(b[2][2], sign);   // w4 = sign*i
  b[2][3] = cmul(b[2][3], w6);
  b[3][1] = cmul(b[3][1], w3);
  b[3][2] = cmul(b[3][2], w6);
  b[3][3] = cmul(b[3][3], w9);
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    cplx t[4] = {b[0][k], b[1][k], b[2][k], b[3][k]};
    dft4(t, sign);
    a[k] = t[0];
    a[k + 4] = t[1];
    a[k + 8] = t[2];
    a[k + 12] = t[3];
  }
}

// the R-1 twiddles W^(step*s), s = 1..R-1, of one butterfly from the table entries at
// step, 2 step, 4 step (and 8 step): the others are one or two complex products away
template <int R, class Tw>
PA_HD void butterfly_twiddles(cplx* w, int step, int L, int sign, const Tw& tw) {
#if defined(PA_FFT_TWIDDLE_LOOKUPS) && PA_FFT_TWIDDLE_LOOKUPS == 1
  // one table lookup per butterfly, the powers by squaring / products (<= 4 roundings
  // more than a direct lookup: still far inside 8 eps log2(L))
  w[1] = tw(step & (L - 1));
  if (sign > 0) w[1].y = -w[1].y;
  if (R > 2) w[2] = cmul(w[1], w[1]);
  if (R > 4) w[4] = cmul(w[2], w[2]);
  if (R > 8) w[8] = cmul(w[4], w[4]);
#else
  w[1] = tw(step & (L - 1));
  if (R > 2) w[2] = tw((2 * step) & (L - 1));
  if (R > 4) w[4] = tw((4 * step) & (L - 1));
  if (R > 8) w[8] = tw((8 * step) & (L - 1));
  if (sign > 0) {
    w[1].y = -w[1].y;
    if (R > 2) w[2].y = -w[2].y;
    if (R > 4) w[4].y = -w[4].y;
    if (R > 8) w[8].y = -w[8].y;
  }
#endif
  if (R > 2) w[3] = cmul(w[1], w[2]);
  if (R > 4) {
    w[5] = cmul(w[4], w[1]);
    w[6] = cmul(w[4], w[2]);
    w[7] = cmul(w[4], w[3]);
  }
  if (R > 8) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int s = 1; s < 8; ++s) w[8 + s] = cmul(w[8], w[s]);
  }
}

// The R-point transform + twiddles of one butterfly on values already in registers
// (the fused kernel's first pass works on the rows it has just loaded from HBM).
template <int R, class Tw>
PA_HD void butterfly_regs(cplx* a, int j, int L, int M, int sign, const Tw& tw) {
  if (R == 2) dft2(a);
  if (R == 4) dft4(a, sign);
  if (R == 8) dft8(a, sign);
  if (R == 16) dft16(a, sign);
  const int step = j * (L / M);   // W_M^(j*s) = W_L^(j*s*L/M)
  if (step > 0) {
    cplx w[R > 1 ? R : 2];
    butterfly_twiddles<R>(w, step, L, sign, tw);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int s = 1; s < R; ++s) a[s] = cmul(a[s], w[s]);
  }
}

// One butterfly `u` (< L/R) of a pass with radix R over sub-length M on one line.
// x.get / x.put map a logical position to the (padded) storage; tw(i) = forward table
// entry W_L^i.  Only W^step, W^(2 step) and W^(4 step) are looked up: the other four
// twiddles of a radix-8 butterfly are one complex product away (<= 2 roundings more),
// which keeps the table traffic -- it shares the load/store path with shared memory --
// at 3/8 of the data traffic instead of 7/8.
template <int R, class Line, class Tw>
PA_HD void butterfly(Line& x, int u, int L, int M, int sign, const Tw& tw) {
  const int Q = M / R;            // stride between the R inputs
  const int g = u / Q, j = u % Q;
  const int base = g * M + j;
  cplx a[R];
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int q = 0; q < R; ++q) a[q] = x.get(base + q * Q);
  butterfly_regs<R>(a, j, L, M, sign, tw);
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int s = 0; s < R; ++s) x.put(base + s * Q, a[s]);
}

}  // namespace pa_fft
