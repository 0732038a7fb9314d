// GPU event simulator for synthetic benchmark input (include/emba_synth.h). Not on the measured path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/emba_synth.h"

struct emba_synth_s {
  int device = 0;
  int64_t n = 0;
  uint32_t* d_pix = nullptr;  // sensor pixel per event, time-sorted
  int64_t* d_t = nullptr;
  uint8_t* d_pol = nullptr;
  int Ws = 0;
};

namespace {

__device__ __forceinline__ double sample_L(const double* __restrict__ L, int W, int H, const double* R, double bx,
                                           double by, double bz) {
  const double X = R[0] * bx + R[1] * by + R[2] * bz;
  const double Y = R[3] * bx + R[4] * by + R[5] * bz;
  const double Z = R[6] * bx + R[7] * by + R[8] * bz;
  const double fx = W / (2.0 * 3.14159265358979323846), fy = H / 3.14159265358979323846;
  const double px = W / 2.0 + fx * atan2(X, Z);
  const double py = H / 2.0 + fy * asin(Y / sqrt(X * X + Y * Y + Z * Z));
  const double x0 = floor(px), y0 = floor(py);
  const double ax = px - x0, ay = py - y0;
  int x0i = ((int)x0 % W + W) % W;
  int x1i = (x0i + 1) % W;
  int y0i = min(max((int)y0, 0), H - 1);
  int y1i = min(y0i + 1, H - 1);
  return (1 - ax) * (1 - ay) * L[(size_t)y0i * W + x0i] + ax * (1 - ay) * L[(size_t)y0i * W + x1i] +
         (1 - ax) * ay * L[(size_t)y1i * W + x0i] + ax * ay * L[(size_t)y1i * W + x1i];
}

// FILL = false: count events per pixel; FILL = true: write them at offs[pixel]
template <bool FILL>
__global__ void k_sim(int S, const double* __restrict__ lut, const double* __restrict__ L, int W, int H, double C,
                      int n_steps, const double* __restrict__ Rs, double t_start, double dt_sim,
                      int32_t* __restrict__ count, const int64_t* __restrict__ offs, int64_t* __restrict__ t_out,
                      uint8_t* __restrict__ pol_out, uint32_t* __restrict__ pix_out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= S) return;
  const double bx = lut[3 * p], by = lut[3 * p + 1], bz = lut[3 * p + 2];
  double Lprev = sample_L(L, W, H, Rs, bx, by, bz);
  double ref = Lprev;
  int64_t o = FILL ? offs[p] : 0;
  int32_t c = 0;
  for (int k = 1; k <= n_steps; k++) {
    const double Lk = sample_L(L, W, H, Rs + (size_t)k * 9, bx, by, bz);
    const double d = Lk - ref;
    const int nc = (int)floor(fabs(d) / C);
    if (nc > 0) {
      const double sgn = d > 0 ? 1.0 : -1.0;
      if (FILL) {
        const double dL = Lk - Lprev;
        for (int j = 1; j <= nc; j++) {
          const double lvl = ref + sgn * (j * C);
          double frac = (lvl - Lprev) / dL;
          frac = fmin(fmax(frac, 0.0), 1.0);
          const double t = t_start + dt_sim * (k - 1) + frac * dt_sim;
          t_out[o] = llrint(t * 1e9);
          pol_out[o] = sgn > 0 ? 1 : 0;
          pix_out[o] = (uint32_t)p;
          o++;
        }
      }
      c += nc;
      ref += sgn * nc * C;
    }
    Lprev = Lk;
  }
  if (!FILL) count[p] = c;
}

// exclusive scan of the per-pixel event counts (a few 10^4 entries): one block, chunks of 1024
__global__ void k_scan_counts(const int32_t* __restrict__ in, int64_t* __restrict__ out, int n) {
  __shared__ int64_t sh[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n + 1; base += 1024) {
    const int i = base + threadIdx.x;
    const int64_t v = i < n ? in[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int64_t u = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += u;
      __syncthreads();
    }
    if (i < n + 1) out[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
}
__global__ void k_xy(const uint32_t* __restrict__ pix, int64_t n, int Ws, uint16_t* __restrict__ x, uint16_t* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { x[i] = (uint16_t)(pix[i] % Ws); y[i] = (uint16_t)(pix[i] / Ws); }
}

}  // namespace

#define SY(call) do { if ((call) != cudaSuccess) { rc = -2; goto done; } } while (0)

extern "C" int emba_synth_simulate(int device, int Ws, int Hs, const double* lut, int Wp, int Hp, const double* L,
                                   double C_th, int n_steps, const double* R_steps, double t_start, double dt_sim,
                                   emba_synth_t* out, int64_t* n_events) {
  if (!lut || !L || !R_steps || !out || !n_events || n_steps < 1) return -1;
  int rc = 0;
  const int S = Ws * Hs;
  double *d_lut = nullptr, *d_L = nullptr, *d_R = nullptr;
  int32_t* d_cnt = nullptr;
  int64_t *d_off = nullptr, *d_t = nullptr;
  uint8_t* d_pol = nullptr;
  uint32_t* d_pix = nullptr;
  int64_t total = 0;
  emba_synth_s* s = nullptr;
  const int T = 128, G = (S + T - 1) / T;
  SY(cudaSetDevice(device));
  SY(cudaMalloc(&d_lut, sizeof(double) * 3 * S));
  SY(cudaMalloc(&d_L, sizeof(double) * (size_t)Wp * Hp));
  SY(cudaMalloc(&d_R, sizeof(double) * 9 * (size_t)(n_steps + 1)));
  SY(cudaMalloc(&d_cnt, sizeof(int32_t) * S));
  SY(cudaMalloc(&d_off, sizeof(int64_t) * (S + 1)));
  SY(cudaMemcpy(d_lut, lut, sizeof(double) * 3 * S, cudaMemcpyHostToDevice));
  SY(cudaMemcpy(d_L, L, sizeof(double) * (size_t)Wp * Hp, cudaMemcpyHostToDevice));
  SY(cudaMemcpy(d_R, R_steps, sizeof(double) * 9 * (size_t)(n_steps + 1), cudaMemcpyHostToDevice));
  k_sim<false><<<G, T>>>(S, d_lut, d_L, Wp, Hp, C_th, n_steps, d_R, t_start, dt_sim, d_cnt, nullptr, nullptr, nullptr, nullptr);
  k_scan_counts<<<1, 1024>>>(d_cnt, d_off, S);
  SY(cudaMemcpy(&total, d_off + S, sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (total >= ((int64_t)1 << 31)) { rc = -1; goto done; }
  {
    const int64_t n = total > 0 ? total : 1;
    SY(cudaMalloc(&d_t, sizeof(int64_t) * n));
    SY(cudaMalloc(&d_pol, n));
    SY(cudaMalloc(&d_pix, sizeof(uint32_t) * n));
  }
  // events come out sensor-pixel major (every pixel's own events in time order); the caller sorts them by time with
  // the library's own event-sequence sort (emba_events_sort_by_time), the step the reference does after parsing a bag
  k_sim<true><<<G, T>>>(S, d_lut, d_L, Wp, Hp, C_th, n_steps, d_R, t_start, dt_sim, nullptr, d_off, d_t, d_pol, d_pix);
  SY(cudaDeviceSynchronize());
  s = new emba_synth_s();
  s->device = device; s->n = total; s->Ws = Ws;
  s->d_t = d_t; d_t = nullptr;
  s->d_pol = d_pol; d_pol = nullptr;
  s->d_pix = d_pix; d_pix = nullptr;
  *out = s;
  *n_events = total;
done:
  cudaFree(d_lut); cudaFree(d_L); cudaFree(d_R); cudaFree(d_cnt); cudaFree(d_off); cudaFree(d_t); cudaFree(d_pol);
  cudaFree(d_pix);
  return rc;
}

extern "C" int emba_synth_fetch(emba_synth_t s, uint16_t* x, uint16_t* y, int64_t* t_ns, uint8_t* pol) {
  if (!s || !x || !y || !t_ns || !pol) return -1;
  if (s->n == 0) return 0;
  if (cudaSetDevice(s->device) != cudaSuccess) return -2;
  uint16_t *dx = nullptr, *dy = nullptr;
  int rc = 0;
  if (cudaMalloc(&dx, sizeof(uint16_t) * s->n) != cudaSuccess || cudaMalloc(&dy, sizeof(uint16_t) * s->n) != cudaSuccess) {
    cudaFree(dx); cudaFree(dy);
    return -2;
  }
  k_xy<<<(int)((s->n + 255) / 256), 256>>>(s->d_pix, s->n, s->Ws, dx, dy);
  if (cudaMemcpy(x, dx, sizeof(uint16_t) * s->n, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(y, dy, sizeof(uint16_t) * s->n, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(t_ns, s->d_t, sizeof(int64_t) * s->n, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(pol, s->d_pol, s->n, cudaMemcpyDeviceToHost) != cudaSuccess)
    rc = -2;
  cudaFree(dx); cudaFree(dy);
  return rc;
}

extern "C" int emba_synth_free(emba_synth_t s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  cudaFree(s->d_pix); cudaFree(s->d_t); cudaFree(s->d_pol);
  delete s;
  return 0;
}
