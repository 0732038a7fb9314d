// "Next" row N3: intensity map from the gradient map,
// poisson_reconstruction::reconstructFromGradient (reference src/image_rec/poisson_reconstruction.cpp:9-50) with
// pde::poisolve, zero Dirichlet boundary (src/image_rec/laplace.cpp:587-797, a1 = a2 = h1 = h2 = 1):
//   F[i][j] = Gx[i][j+1] - Gx[i][j] + Gy[i+1][j] - Gy[i][j]   (i < H-1, j < W-1; zero on the last row / column)
//   rhs = DST-I_2D(F) / (4 (H+1)(W+1));  U = rhs / (lambda1[i] + lambda2[j]);  M = DST-I_2D(U)
//   lambda[k] = -4 sin^2(pi (k+1) / (2 (n+1))),  DST-I: Y_k = 2 sum_j X_j sin(pi (j+1)(k+1) / (n+1))  (FFTW RODFT00)
//
// B200 design. The reference goes through FFTW; cuFFT has no DST and an odd-extension FFT quadruples the data. The
// DST-I is a multiplication with the symmetric sine matrix S_n[k][j] = 2 sin(pi (j+1)(k+1)/(n+1)), so the 2-D
// transform is S_H * X * S_W: two dense fp64 GEMMs -- genuinely dense contractions, run on the fp64 tensor cores
// (DMMA m8n8k4, cp.async double-buffered 64x64x16 tiles). At panorama sizes (1024x512 ... 4096x2048) that is
// 3 ... 200 GFLOP per reconstruction, milliseconds, with O(sqrt(n) eps) rounding. The sine matrices are built once
// per plan with exact integer argument reduction; the divergence and the eigenvalue division are fused elementwise
// kernels. Everything stays on the device; the map never leaves HBM when called through the optimiser handle.
#include <math_constants.h>

#include "emba_internal.cuh"

namespace emba {

struct PoissonPlan {
  int device = 0;
  int W = 0, H = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  double* SH = nullptr;   // [H][H] sine matrix
  double* SW = nullptr;   // [W][W]
  double* lamH = nullptr; // [H] eigenvalues
  double* lamW = nullptr; // [W]
  double* F = nullptr;    // [H][W] work
  double* T = nullptr;    // [H][W] work
  double* dGx = nullptr;  // staging for the host entry point
  double* dGy = nullptr;
  int64_t launches = 0;
  float last_ms = 0.f;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  std::string err;
};

// S[k][j] = 2 sin(pi (j+1)(k+1) / (n+1)); the argument is reduced exactly in integers before sinpi
__global__ void k_sine_matrix(int n, double* __restrict__ S, double* __restrict__ lam) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) lam[idx] = -4.0 * sinpi((double)(idx + 1) / (2.0 * (n + 1))) * sinpi((double)(idx + 1) / (2.0 * (n + 1)));
  if (idx >= (int64_t)n * n) return;
  const int k = (int)(idx / n), j = (int)(idx % n);
  const int64_t m = ((int64_t)(j + 1) * (k + 1)) % (2LL * (n + 1));
  S[idx] = 2.0 * sinpi((double)m / (double)(n + 1));
}

// divergence by forward differences (poisson_reconstruction.cpp:22-30)
__global__ void k_divergence(int H, int W, const double* __restrict__ Gx, const double* __restrict__ Gy,
                             double* __restrict__ F) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)H * W) return;
  const int i = (int)(idx / W), j = (int)(idx % W);
  double f = 0.0;
  if (i < H - 1 && j < W - 1) f = Gx[idx + 1] - Gx[idx] + Gy[idx + W] - Gy[idx];
  F[idx] = f;
}

// solve in eigen space (laplace.cpp:676-735): rhs * (1 / fft_norm) / (lambda1[i] + lambda2[j])
__global__ void k_eigen_divide(int H, int W, const double* __restrict__ lamH, const double* __restrict__ lamW,
                               double inv_norm, double* __restrict__ Y) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)H * W) return;
  const int i = (int)(idx / W), j = (int)(idx % W);
  const double rhs = Y[idx] * inv_norm;
  const double div = lamH[i] + lamW[j];
  Y[idx] = div == 0.0 ? 0.0 : rhs / div;
}

// ---------------------------------------------------------------------------------------------------
// C[M][N] = A[M][K] * B[K][N], all row-major fp64, leading dimensions = widths (even, so rows are 16-byte aligned).
// CTA tile 64 x 64, K step 16, 4 warps (2 x 2), each warp a 32 x 32 block = 4 x 4 DMMA m8n8k4 accumulators.
// ---------------------------------------------------------------------------------------------------
constexpr int kGT = 64, kGK = 16, kGThreads = 128;
constexpr int kApad = 20;  // A tile row stride (doubles): 4g + k mod 16 covers every bank pair twice
constexpr int kBpad = 72;  // B tile row stride: 8k + g mod 16 likewise

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16z(void* smem, const void* gmem, int bytes) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kGThreads)
k_gemm_f64(int M, int N, int K, const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C) {
  __shared__ __align__(16) double As[2][kGT][kApad];
  __shared__ __align__(16) double Bs[2][kGK][kBpad];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, kk = lane & 3;
  const int m0 = blockIdx.y * kGT, n0 = blockIdx.x * kGT;
  const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
  double acc[4][4][2];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) acc[r][c][0] = acc[r][c][1] = 0.0;
  auto issue = [&](int k0, int st) {
    // A tile: 64 rows x 16 doubles = 512 chunks of 16 bytes; B tile: 16 rows x 64 doubles = 512 chunks
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int e = tid + kGThreads * q;
      {
        const int r = e >> 3, c = (e & 7) * 2;
        const int gr = m0 + r, gc = k0 + c;
        const bool ok = gr < M && gc < K;
        cp16z(&As[st][r][c], ok ? A + (size_t)gr * K + gc : A, ok ? (gc + 1 < K ? 16 : 8) : 0);
      }
      {
        const int r = e >> 5, c = (e & 31) * 2;
        const int gr = k0 + r, gc = n0 + c;
        const bool ok = gr < K && gc < N;
        cp16z(&Bs[st][r][c], ok ? B + (size_t)gr * N + gc : B, ok ? (gc + 1 < N ? 16 : 8) : 0);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  const int nk = (K + kGK - 1) / kGK;
  issue(0, 0);
  for (int t = 0; t < nk; t++) {
    const int st = t & 1;
    if (t + 1 < nk) issue((t + 1) * kGK, st ^ 1);
    else asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < kGK; k0 += 4) {
      double a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; r++) a[r] = As[st][wr + 8 * r + g][k0 + kk];
#pragma unroll
      for (int c = 0; c < 4; c++) b[c] = Bs[st][k0 + kk][wc + 8 * c + g];
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) dmma884(acc[r][c][0], acc[r][c][1], a[r], b[c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int gr = m0 + wr + 8 * r + g, gc = n0 + wc + 8 * c + 2 * kk;
      if (gr < M) {
        if (gc < N) C[(size_t)gr * N + gc] = acc[r][c][0];
        if (gc + 1 < N) C[(size_t)gr * N + gc + 1] = acc[r][c][1];
      }
    }
}

static int gemm(PoissonPlan* p, int M, int N, int K, const double* A, const double* B, double* C) {
  dim3 grid((N + kGT - 1) / kGT, (M + kGT - 1) / kGT);
  k_gemm_f64<<<grid, kGThreads, 0, p->stream>>>(M, N, K, A, B, C);
  p->launches++;
  return cudaGetLastError() == cudaSuccess ? EMBA_OK : EMBA_E_CUDA;
}

int poisson_plan_create(int device, int W, int H, cudaStream_t stream, PoissonPlan** out) {
  if (!out || W < 2 || H < 2) return EMBA_E_ARG;
  if ((W & 1) || (H & 1)) return EMBA_E_SUPPORT;  // 16-byte aligned rows for the async tile copies
  if (cudaSetDevice(device) != cudaSuccess) return EMBA_E_CUDA;
  PoissonPlan* p = new PoissonPlan();
  p->device = device; p->W = W; p->H = H;
  if (stream) p->stream = stream;
  else {
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) { delete p; return EMBA_E_CUDA; }
    p->own_stream = true;
  }
  const size_t P = (size_t)W * H;
  bool ok = cudaMalloc((void**)&p->SH, sizeof(double) * H * H) == cudaSuccess &&
            cudaMalloc((void**)&p->SW, sizeof(double) * W * W) == cudaSuccess &&
            cudaMalloc((void**)&p->lamH, sizeof(double) * H) == cudaSuccess &&
            cudaMalloc((void**)&p->lamW, sizeof(double) * W) == cudaSuccess &&
            cudaMalloc((void**)&p->F, sizeof(double) * P) == cudaSuccess &&
            cudaMalloc((void**)&p->T, sizeof(double) * P) == cudaSuccess;
  ok = ok && cudaEventCreate(&p->e0) == cudaSuccess && cudaEventCreate(&p->e1) == cudaSuccess;
  if (!ok) {
    for (double* q : {p->SH, p->SW, p->lamH, p->lamW, p->F, p->T}) if (q) cudaFree(q);
    if (p->own_stream) cudaStreamDestroy(p->stream);
    delete p;
    return EMBA_E_CUDA;
  }
  k_sine_matrix<<<(unsigned)(((size_t)H * H + 255) / 256), 256, 0, p->stream>>>(H, p->SH, p->lamH);
  k_sine_matrix<<<(unsigned)(((size_t)W * W + 255) / 256), 256, 0, p->stream>>>(W, p->SW, p->lamW);
  p->launches += 2;
  if (cudaGetLastError() != cudaSuccess) return EMBA_E_CUDA;
  *out = p;
  return EMBA_OK;
}

void poisson_plan_destroy(PoissonPlan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  cudaStreamSynchronize(p->stream);
  for (double* q : {p->SH, p->SW, p->lamH, p->lamW, p->F, p->T, p->dGx, p->dGy}) if (q) cudaFree(q);
  if (p->e0) cudaEventDestroy(p->e0);
  if (p->e1) cudaEventDestroy(p->e1);
  if (p->own_stream) cudaStreamDestroy(p->stream);
  delete p;
}

// device pointers in, device pointer out. The inputs are read by the first kernel only, so dOut may alias one of them.
int poisson_solve_device(PoissonPlan* p, const double* dGx, const double* dGy, double* dOut) {
  const int W = p->W, H = p->H;
  const int64_t P = (int64_t)W * H;
  const int T = 256;
  cudaEventRecord(p->e0, p->stream);
  k_divergence<<<(unsigned)((P + T - 1) / T), T, 0, p->stream>>>(H, W, dGx, dGy, p->F);
  p->launches++;
  int rc;
  if ((rc = gemm(p, H, W, W, p->F, p->SW, p->T))) return rc;     // rows:    T = F S_W
  if ((rc = gemm(p, H, W, H, p->SH, p->T, p->F))) return rc;     // columns: F = S_H T
  k_eigen_divide<<<(unsigned)((P + T - 1) / T), T, 0, p->stream>>>(H, W, p->lamH, p->lamW,
                                                                 1.0 / (4.0 * (double)(H + 1) * (double)(W + 1)), p->F);
  p->launches++;
  if ((rc = gemm(p, H, W, W, p->F, p->SW, p->T))) return rc;
  if ((rc = gemm(p, H, W, H, p->SH, p->T, dOut))) return rc;
  cudaEventRecord(p->e1, p->stream);
  return cudaGetLastError() == cudaSuccess ? EMBA_OK : EMBA_E_CUDA;
}

}  // namespace emba

using namespace emba;

extern "C" {

int emba_poisson_create(int32_t device, int32_t pano_w, int32_t pano_h, emba_poisson_t* out) {
  PoissonPlan* p = nullptr;
  const int rc = poisson_plan_create(device, pano_w, pano_h, nullptr, &p);
  if (rc == EMBA_OK) *out = (emba_poisson_t)p;
  return rc;
}

int emba_poisson_destroy(emba_poisson_t pp) {
  poisson_plan_destroy((PoissonPlan*)pp);
  return EMBA_OK;
}

int emba_poisson_reconstruct(emba_poisson_t pp, const double* Gx, const double* Gy, double* img_out) {
  PoissonPlan* p = (PoissonPlan*)pp;
  if (!p || !Gx || !Gy || !img_out) return EMBA_E_ARG;
  if (cudaSetDevice(p->device) != cudaSuccess) return EMBA_E_CUDA;
  const size_t bytes = sizeof(double) * (size_t)p->W * p->H;
  if (!p->dGx && (cudaMalloc((void**)&p->dGx, bytes) != cudaSuccess || cudaMalloc((void**)&p->dGy, bytes) != cudaSuccess))
    return EMBA_E_CUDA;
  if (cudaMemcpyAsync(p->dGx, Gx, bytes, cudaMemcpyHostToDevice, p->stream) != cudaSuccess ||
      cudaMemcpyAsync(p->dGy, Gy, bytes, cudaMemcpyHostToDevice, p->stream) != cudaSuccess)
    return EMBA_E_CUDA;
  const int rc = poisson_solve_device(p, p->dGx, p->dGy, p->dGy);
  if (rc != EMBA_OK) return rc;
  if (cudaMemcpyAsync(img_out, p->dGy, bytes, cudaMemcpyDeviceToHost, p->stream) != cudaSuccess) return EMBA_E_CUDA;
  if (cudaStreamSynchronize(p->stream) != cudaSuccess) return EMBA_E_CUDA;
  cudaEventElapsedTime(&p->last_ms, p->e0, p->e1);
  return EMBA_OK;
}

int emba_poisson_last_ms(emba_poisson_t pp, double* ms_out, int64_t* launches_out) {
  PoissonPlan* p = (PoissonPlan*)pp;
  if (!p) return EMBA_E_ARG;
  if (ms_out) *ms_out = p->last_ms;
  if (launches_out) *launches_out = p->launches;
  return EMBA_OK;
}

// the optimiser's own map, device resident (solver.cpp:412-417 reconstructs from the evolving Gx, Gy)
int emba_reconstruct_map(emba_handle_t hh, int32_t which_state, double* img_out) {
  Handle* h = (Handle*)hh;
  if (!h || !img_out) return EMBA_E_ARG;
  if (which_state < 0 || which_state > 1) { h->err = "emba_reconstruct_map: which_state must be 0 or 1"; return EMBA_E_ARG; }
  EMBA_CUDA(cudaSetDevice(h->device));
  if (h->n <= 0) { h->err = "emba_reconstruct_map: set the state first"; return EMBA_E_ARG; }
  PoissonPlan* p = (PoissonPlan*)h->poisson;
  if (!p) {
    const int rc = poisson_plan_create(h->device, h->Wp, h->Hp, h->stream, &p);
    if (rc != EMBA_OK) { h->err = "emba_reconstruct_map: panorama width and height must be even"; return rc; }
    h->poisson = p;
  }
  const StateSlot& s = h->st[which_state ? 1 - h->cur : h->cur];
  const size_t bytes = sizeof(double) * (size_t)h->P;
  if (!p->dGx) EMBA_CUDA(cudaMalloc((void**)&p->dGx, bytes));
  EMBA_TRY(poisson_solve_device(p, s.Gx, s.Gy, p->dGx));
  h->launches += 7;
  EMBA_CUDA(cudaMemcpyAsync(img_out, p->dGx, bytes, cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&p->last_ms, p->e0, p->e1);
  return EMBA_OK;
}

}  // extern "C"
