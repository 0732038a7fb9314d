// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the FFTW3 real-to-real interface the reference's Poisson
// solver calls (/root/reference/src/image_rec/laplace.cpp:640-760): fftw_plan_r2r_2d with FFTW_RODFT00 (DST-I) or
// FFTW_REDFT00 (DCT-I) in both dimensions, fftw_execute, fftw_destroy_plan and the thread / cleanup no-ops.
// The transforms follow FFTW's published definitions (fftw3 manual, "1d Real-odd DFTs (DSTs)" / "Real-even DFTs"):
//   RODFT00: Y_k = 2 sum_{j=0}^{n-1} X_j sin(pi (j+1)(k+1) / (n+1))
//   REDFT00: Y_k = X_0 + (-1)^k X_{n-1} + 2 sum_{j=1}^{n-2} X_j cos(pi j k / (n-1))
// evaluated directly (O(n^2) per line, long double accumulation, exact integer argument reduction) -- slower than
// an FFT but at least as accurate, which is what an oracle needs. FFTW is not installed in this image.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

typedef enum { FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2, FFTW_REDFT00 = 3, FFTW_REDFT01 = 4, FFTW_REDFT10 = 5,
               FFTW_REDFT11 = 6, FFTW_RODFT00 = 7, FFTW_RODFT01 = 8, FFTW_RODFT10 = 9, FFTW_RODFT11 = 10 } fftw_r2r_kind;
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

struct fftw_plan_shim { int n0, n1; double* in; double* out; fftw_r2r_kind k0, k1; };
typedef fftw_plan_shim* fftw_plan;

inline fftw_plan fftw_plan_r2r_2d(int n0, int n1, double* in, double* out, fftw_r2r_kind k0, fftw_r2r_kind k1, unsigned) {
  return new fftw_plan_shim{n0, n1, in, out, k0, k1};
}
inline void fftw_destroy_plan(fftw_plan p) { delete p; }
inline int fftw_init_threads() { return 1; }
inline void fftw_plan_with_nthreads(int) {}
inline void fftw_cleanup_threads() {}
inline void fftw_cleanup() {}

namespace fftw_shim {
// one r2r transform of `n` elements with stride `st`, in place via a temporary
inline void line(double* x, int n, std::ptrdiff_t st, fftw_r2r_kind kind, const std::vector<long double>& tab,
                 std::vector<long double>& tmp) {
  tmp.assign((size_t)n, 0.0L);
  if (kind == FFTW_RODFT00) {
    const long long period = 2LL * (n + 1);  // tab[m] = sin(pi m / (n+1)), m in [0, period)
    for (int k = 0; k < n; k++) {
      long double s = 0.0L;
      for (int j = 0; j < n; j++) s += (long double)x[j * st] * tab[(size_t)(((long long)(j + 1) * (k + 1)) % period)];
      tmp[(size_t)k] = 2.0L * s;
    }
  } else {  // FFTW_REDFT00
    const long long period = 2LL * (n - 1);  // tab[m] = cos(pi m / (n-1))
    for (int k = 0; k < n; k++) {
      long double s = 0.0L;
      for (int j = 1; j < n - 1; j++) s += (long double)x[j * st] * tab[(size_t)(((long long)j * k) % period)];
      tmp[(size_t)k] = (long double)x[0] + ((k & 1) ? -1.0L : 1.0L) * (long double)x[(n - 1) * st] + 2.0L * s;
    }
  }
  for (int k = 0; k < n; k++) x[k * st] = (double)tmp[(size_t)k];
}
inline std::vector<long double> table(int n, fftw_r2r_kind kind) {
  const long double pi = 3.14159265358979323846264338327950288L;
  const int den = (kind == FFTW_RODFT00) ? n + 1 : n - 1;
  std::vector<long double> t((size_t)(2 * den > 0 ? 2 * den : 1));
  for (int m = 0; m < 2 * den; m++) t[(size_t)m] = (kind == FFTW_RODFT00) ? sinl(pi * m / den) : cosl(pi * m / den);
  return t;
}
}  // namespace fftw_shim

inline void fftw_execute(const fftw_plan p) {
  if (p->in != p->out)
    for (long i = 0; i < (long)p->n0 * p->n1; i++) p->out[i] = p->in[i];
  std::vector<long double> tmp;
  const std::vector<long double> t1 = fftw_shim::table(p->n1, p->k1), t0 = fftw_shim::table(p->n0, p->k0);
#pragma omp parallel for private(tmp)
  for (int i = 0; i < p->n0; i++) fftw_shim::line(p->out + (size_t)i * p->n1, p->n1, 1, p->k1, t1, tmp);
#pragma omp parallel for private(tmp)
  for (int j = 0; j < p->n1; j++) fftw_shim::line(p->out + j, p->n0, p->n1, p->k0, t0, tmp);
}
