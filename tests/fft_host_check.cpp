// CPU run of the shared FFT core (pencilarrays.jl_b200/csrc/fft_core.hpp): the same
// radix plan, butterflies, padded indexing and digit-reversed read-out the CUDA kernel
// executes.  TEST INFRASTRUCTURE -- not linked into libpa_b200, never used by the product.
//   fft_host_check L sign [logR1] < in.bin > out.bin      (L complex doubles each way;
//   logR1 > 0: first pass of radix 2^logR1 on register values, like the kernel's gather)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../pencilarrays.jl_b200/csrc/fft_core.hpp"

using namespace pa_fft;

struct Line {
  cplx* p;
  cplx get(int i) const { return p[pad_index(i)]; }
  void put(int i, cplx v) { p[pad_index(i)] = v; }
};

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const int L = atoi(argv[1]), sign = atoi(argv[2]);
  int logL = 0;
  while ((1 << logL) < L) ++logL;
  if ((1 << logL) != L) return 3;
  std::vector<cplx> in(L), tw(L), buf(padded_pitch(L));
  if (fread(in.data(), sizeof(cplx), L, stdin) != (size_t)L) return 4;
  for (int k = 0; k < L; ++k) {
    long double a = -2.0L * 3.141592653589793238462643383279502884L * k / L;
    tw[k] = cplx{(double)cosl(a), (double)sinl(a)};
  }
  Line x{buf.data()};
  for (int i = 0; i < L; ++i) x.put(i, in[i]);
  const int logR1 = argc > 3 ? atoi(argv[3]) : 0;
  const Radices R = logR1 > 0 ? radices_with_first(logL, logR1) : radices_of(logL);
  auto twf = [&](int i) { return tw[i]; };
  int M = L;
  for (int pss = 0; pss < R.n; ++pss) {
    const int r = R.r[pss];
    for (int u = 0; u < L / r; ++u) {
      if (r == 2) butterfly<2>(x, u, L, M, sign, twf);
      if (r == 4) butterfly<4>(x, u, L, M, sign, twf);
      if (r == 8) butterfly<8>(x, u, L, M, sign, twf);
      if (r == 16) butterfly<16>(x, u, L, M, sign, twf);
    }
    M /= r;
  }
  std::vector<cplx> out(L);
  for (int k = 0; k < L; ++k) out[k] = x.get(fft_position_of(k, L, R));
  fwrite(out.data(), sizeof(cplx), L, stdout);
  return 0;
}
