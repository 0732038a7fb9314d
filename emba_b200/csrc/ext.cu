// EXTENSION MODE of the hot path (the variant BASELINE.json's north star words literally):
//   * cubic cumulative SO(3) B-spline evaluated PER EVENT, value and derivatives w.r.t. the 4 active control poses
//     re-derived from basalt's So3Spline<4>::evaluate (thirdparty/basalt-headers/include/basalt/spline/
//     so3_spline.h:218-274, wrapped by the reference's CubicTrajectory::evaluate, src/utils/trajectory.cpp:453-479);
//   * the bearing projected into the equirectangular panorama at t and at the time of the previous event of the
//     sensor pixel, BILINEAR gather of the gradient map (8 map Jacobian entries), linearised event generation model
//     with 2 x 12 rotation Jacobian entries (the reference's parity mode: 2 x 6 and nearest pixel, model.cpp:140-246);
//   * normal equations in Jacobian form: the rows are kept explicitly (sparse J), g = J^T e and the block diagonal of
//     H = J^T J (3x3 per control pose, 2x2 per active pixel) are assembled -- the pose side through warp-shuffle
//     reductions (time-sorted rows share their control poses), the rest through fp64 atomics -- and the damped
//     system (H + lambda diag H) x = g is solved by block-Jacobi PCG with the matrix-free product J^T (J v).
//     With bilinear sampling the map block of H is no longer block diagonal, so the reference's Schur complement
//     onto the control poses (model.cpp:721-792) has no cheap analogue here; PCG is its counterpart (model.cpp:794-840).
// There is no reference implementation of this mode: results are checked (tests/test_ext.py) against an fp64 CPU restatement of the same model
// on the reference's vendored basalt / Sophus, itself checked against finite differences, and are reported separately
// from the parity mode.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "emba_internal.cuh"

namespace emba {

PanoCam make_cam(const Handle* h);

constexpr int kExtKnot = 24;  // per knot i: R_i (9), delta_i = Log(R_i^-1 R_{i+1}) (3), Jl^-1(delta_i) (9), pad

struct ExtHandle {
  Handle* h = nullptr;
  const int64_t* d_t = nullptr;  // timestamps of the window's events (inside the device-resident sequence)
  int64_t Mt = 0;                // pairs of the window
  int64_t t0_ns = 0, dt_ns = 0;
  int n = 0;
  // state
  double *d_quat = nullptr, *d_Gx = nullptr, *d_Gy = nullptr, *d_knot = nullptr;
  // rows (SoA, one per pair, time order of the current event)
  uint32_t* d_mev = nullptr;   // current event of the pair
  double* d_e = nullptr;       // residual
  double2* d_dp = nullptr;     // displacement
  double2* d_pm = nullptr;     // warped current event
  double2* d_axy = nullptr;    // bilinear fractions
  int4* d_pix = nullptr;       // footprint (linear map indices)
  int2* d_cp = nullptr;        // first control pose of the current / previous event's spline segment
  double* d_J = nullptr;       // [Mt][24] Jc(12) Jp(12)
  int32_t* d_flag = nullptr;   // 0 outlier, 1 inlier, 3 inlier and used (all four footprint pixels active)
  int32_t* d_cnt = nullptr;    // [P] footprint hits
  int32_t* d_amap = nullptr;   // [P]
  int32_t* d_apix = nullptr;   // [P]
  int64_t Np = 0, M = 0, Mused = 0;
  // assembled parts: g = [b1 (3n) | b2 (2Np)], block diagonal of H: pose blocks [n][9], pixel blocks [Np][3]
  double *d_g = nullptr, *d_Bp = nullptr, *d_Bm = nullptr;
  double *d_vec = nullptr;     // PCG vectors
  int64_t vec_cap = 0;
  double* d_tmp = nullptr;     // [Mt] J v
  double* d_x = nullptr;       // last solution [3n + 2P]
  double* d_part = nullptr;
  void* d_scan = nullptr;
  int64_t* d_tot = nullptr;
  double t_ms[4] = {0, 0, 0, 0};
  bool evaluated = false, formed = false, solved = false;
};

// ---- small 3x3 helpers (row-major)
struct M3 { double m[9]; };
__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) c.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
  return c;
}
__device__ __forceinline__ M3 m3_mul_bt(const M3& a, const double* __restrict__ b) {  // a * b^T
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) c.m[3 * i + j] = a.m[3 * i] * b[3 * j] + a.m[3 * i + 1] * b[3 * j + 1] + a.m[3 * i + 2] * b[3 * j + 2];
  return c;
}
// I + a K + b K^2 for K = [v]x
__device__ __forceinline__ M3 m3_poly(const double v[3], double a, double b) {
  const double x = v[0], y = v[1], z = v[2];
  M3 r = {{1 - b * (y * y + z * z), -a * z + b * x * y, a * y + b * x * z,
           a * z + b * x * y, 1 - b * (x * x + z * z), -a * x + b * y * z,
           -a * y + b * x * z, a * x + b * y * z, 1 - b * (x * x + y * y)}};
  return r;
}
// Sophus SO3::exp as a matrix (so3.hpp:583-619 thresholds)
__device__ __forceinline__ M3 m3_exp(const double v[3]) {
  const double th2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  double a, b;
  if (th2 < kSophusEps * kSophusEps) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; }
  else { const double th = sqrt(th2); a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
  return m3_poly(v, a, b);
}
// leftJacobianSO3 / leftJacobianInvSO3 (sophus_utils.hpp:333-414)
__device__ __forceinline__ M3 m3_Jl(const double v[3]) {
  const double th2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  double a, b;
  if (th2 > kSophusEps) { const double th = sqrt(th2); a = (1.0 - cos(th)) / th2; b = (th - sin(th)) / (th2 * th); }
  else { a = 0.5; b = 1.0 / 6.0; }
  return m3_poly(v, a, b);
}
__device__ __forceinline__ M3 m3_JlInv(const double v[3]) {
  const double th2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  double c;
  if (th2 > kSophusEps) {
    const double th = sqrt(th2);
    if (th < 3.14159265358979323846 - 1e-5) c = 1.0 / th2 - (1.0 + cos(th)) / (2.0 * th * sin(th));
    else c = 1.0 / (3.14159265358979323846 * 3.14159265358979323846);
  } else c = 1.0 / 12.0;
  return m3_poly(v, -0.5, c);
}

__global__ void k_ext_knots(const double* __restrict__ quat, int n, double* __restrict__ knot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 q0 = reinterpret_cast<const double4*>(quat)[i];
  const Mat3 R = quat_to_R(q0);
  double* o = knot + (size_t)i * kExtKnot;
#pragma unroll
  for (int k = 0; k < 9; k++) o[k] = R.m[k];
  double d[3] = {0, 0, 0};
  if (i + 1 < n) {
    const double4 q1 = reinterpret_cast<const double4*>(quat)[i + 1];
    const Vec3 l = so3_log(quat_mul(make_double4(-q0.x, -q0.y, -q0.z, q0.w), q1));  // so3_spline.h:249-250
    d[0] = l.x; d[1] = l.y; d[2] = l.z;
  }
  const M3 Ji = m3_JlInv(d);
  o[9] = d[0]; o[10] = d[1]; o[11] = d[2];
#pragma unroll
  for (int k = 0; k < 9; k++) o[12 + k] = Ji.m[k];
}

// So3Spline<4>::evaluate (so3_spline.h:218-274): R = R_s Exp(l1 d0) Exp(l2 d1) Exp(l3 d2); D[k] = d R / d knot_{s+k}
// (left perturbations): D0 = I - H0, D1 = H0 - H1, D2 = H1 - H2, D3 = H2,
//   H_i = l_{i+1} (R_s prod_{j<i} Exp(l_{j+1} d_j)) Jl(l_{i+1} d_i) Jl^-1(d_i) R_{s+i}^T
__device__ __forceinline__ int spline4(const double* __restrict__ knot, int n, int64_t t, int64_t t0, int64_t dt, M3& R,
                                       M3 D[4], bool& ok) {
  const int64_t st = t - t0;
  int64_t s = st / dt;
  const double u = (double)(st % dt) / (double)dt;
  ok = st >= 0 && s + 4 <= (int64_t)n;
  if (!ok) s = 0;
  const double u2 = u * u, u3 = u2 * u;
  const double lam[3] = {(5.0 + 3.0 * u - 3.0 * u2 + u3) / 6.0, (1.0 + 3.0 * u + 3.0 * u2 - 2.0 * u3) / 6.0, u3 / 6.0};
  const double* k0 = knot + (size_t)s * kExtKnot;
#pragma unroll
  for (int k = 0; k < 9; k++) R.m[k] = k0[k];
  M3 Hprev;
#pragma unroll
  for (int k = 0; k < 9; k++) Hprev.m[k] = (k % 4 == 0) ? 1.0 : 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double* ki = knot + (size_t)(s + i) * kExtKnot;
    const double kd[3] = {lam[i] * ki[9], lam[i] * ki[10], lam[i] * ki[11]};
    M3 Jli;
#pragma unroll
    for (int k = 0; k < 9; k++) Jli.m[k] = ki[12 + k];
    M3 H = m3_mul_bt(m3_mul(m3_mul(R, m3_Jl(kd)), Jli), ki);
#pragma unroll
    for (int k = 0; k < 9; k++) { H.m[k] *= lam[i]; D[i].m[k] = Hprev.m[k] - H.m[k]; }
    Hprev = H;
    R = m3_mul(R, m3_exp(kd));
  }
  D[3] = Hprev;
  return (int)s;
}

constexpr int kExtThreads = 128;

// one thread per pair: both warps, residual, Jacobian row, footprint counts, cost partials
__global__ void __launch_bounds__(kExtThreads)
k_ext_eval(int64_t Mt, const uint32_t* __restrict__ mev, const int32_t* __restrict__ prev, const int64_t* __restrict__ t_ns,
           const uint32_t* __restrict__ spix, const uint8_t* __restrict__ pol, const double* __restrict__ lut,
           const double* __restrict__ knot, int n, int64_t t0, int64_t dt, const double* __restrict__ Gx,
           const double* __restrict__ Gy, PanoCam cam, int W, int H, double C_th, double* __restrict__ e_out,
           double2* __restrict__ dp_out, double2* __restrict__ pm_out, double2* __restrict__ axy_out,
           int4* __restrict__ pix_out, int2* __restrict__ cp_out, double* __restrict__ J_out, int32_t* __restrict__ flag_out,
           int32_t* __restrict__ cnt, double* __restrict__ part, int32_t* __restrict__ flags) {
  const int64_t m = (int64_t)blockIdx.x * kExtThreads + threadIdx.x;
  double cost = 0.0, num = 0.0;
  if (m < Mt) {
    const uint32_t ic = mev[m];
    const uint32_t ip = (uint32_t)prev[ic];
    const size_t sp = spix[ic];
    double bx = lut[3 * sp], by = lut[3 * sp + 1], bz = lut[3 * sp + 2];
    const double inv = 1.0 / sqrt(bx * bx + by * by + bz * bz);
    bx *= inv; by *= inv; bz *= inv;
    M3 Rc, Dc[4], Rp, Dp[4];
    bool okc, okp;
    const int sc = spline4(knot, n, t_ns[ic], t0, dt, Rc, Dc, okc);
    const int sq = spline4(knot, n, t_ns[ip], t0, dt, Rp, Dp, okp);
    if (!okc || !okp) atomicOr(flags, 2);
    const double Xc = Rc.m[0] * bx + Rc.m[1] * by + Rc.m[2] * bz, Yc = Rc.m[3] * bx + Rc.m[4] * by + Rc.m[5] * bz,
                 Zc = Rc.m[6] * bx + Rc.m[7] * by + Rc.m[8] * bz;
    const double Xp = Rp.m[0] * bx + Rp.m[1] * by + Rp.m[2] * bz, Yp = Rp.m[3] * bx + Rp.m[4] * by + Rp.m[5] * bz,
                 Zp = Rp.m[6] * bx + Rp.m[7] * by + Rp.m[8] * bz;
    double pcx, pcy, ppx, ppy;
    project_pm_unit(cam, Xc, Yc, Zc, pcx, pcy);
    project_pm_unit(cam, Xp, Yp, Zp, ppx, ppy);
    const double dx = pcx - ppx, dy = pcy - ppy;
    const double nrm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    int32_t fl = 0;
    double e = 0.0;
    int4 px4 = make_int4(0, 0, 0, 0);
    double ax = 0.0, ay = 0.0;
    double Jrow[24];
#pragma unroll
    for (int k = 0; k < 24; k++) Jrow[k] = 0.0;
    if (!(nrm > 10.0) && okc && okp) {  // outlier gate of model.cpp:199-205
      fl = 1;
      const double fx0 = floor(pcx), fy0 = floor(pcy);
      ax = pcx - fx0; ay = pcy - fy0;
      int x0 = (int)fx0 % W; if (x0 < 0) x0 += W;
      const int x1 = (x0 + 1) % W;
      const int y0 = min(max((int)fy0, 0), H - 1), y1 = min(max((int)fy0 + 1, 0), H - 1);
      px4 = make_int4(y0 * W + x0, y0 * W + x1, y1 * W + x0, y1 * W + x1);
      const double g00x = Gx[px4.x], g10x = Gx[px4.y], g01x = Gx[px4.z], g11x = Gx[px4.w];
      const double g00y = Gy[px4.x], g10y = Gy[px4.y], g01y = Gy[px4.z], g11y = Gy[px4.w];
      const double w00 = (1 - ax) * (1 - ay), w10 = ax * (1 - ay), w01 = (1 - ax) * ay, w11 = ax * ay;
      const double gx = w00 * g00x + w10 * g10x + w01 * g01x + w11 * g11x;
      const double gy = w00 * g00y + w10 * g10y + w01 * g01y + w11 * g11y;
      // derivative of the interpolant w.r.t. the sampling position
      const double dgx_dx = (1 - ay) * (g10x - g00x) + ay * (g11x - g01x), dgy_dx = (1 - ay) * (g10y - g00y) + ay * (g11y - g01y);
      const double dgx_dy = (1 - ax) * (g01x - g00x) + ax * (g11x - g10x), dgy_dy = (1 - ax) * (g01y - g00y) + ax * (g11y - g10y);
      const double C_pred = gx * dx + gy * dy;
      e = 2.0 * ((pol[ic] ? 1.0 : 0.0) - 0.5) * C_th - C_pred;  // model.cpp:217-221
      const double h0 = gx + dx * dgx_dx + dy * dgy_dx, h1 = gy + dx * dgx_dy + dy * dgy_dy;
      double Mj[6], v[3];
      project_jac(cam, Xc, Yc, Zc, Mj);
      v[0] = h0 * Mj[0] + h1 * Mj[3]; v[1] = h0 * Mj[1]; v[2] = h0 * Mj[2] + h1 * Mj[5];  // temp * dpm_drb * (-[rb]x)
#pragma unroll
      for (int k = 0; k < 4; k++)
#pragma unroll
        for (int j = 0; j < 3; j++) Jrow[3 * k + j] = v[0] * Dc[k].m[j] + v[1] * Dc[k].m[3 + j] + v[2] * Dc[k].m[6 + j];
      project_jac(cam, Xp, Yp, Zp, Mj);
      v[0] = -(gx * Mj[0] + gy * Mj[3]); v[1] = -(gx * Mj[1]); v[2] = -(gx * Mj[2] + gy * Mj[5]);
#pragma unroll
      for (int k = 0; k < 4; k++)
#pragma unroll
        for (int j = 0; j < 3; j++) Jrow[12 + 3 * k + j] = v[0] * Dp[k].m[j] + v[1] * Dp[k].m[3 + j] + v[2] * Dp[k].m[6 + j];
      atomicAdd(&cnt[px4.x], 1); atomicAdd(&cnt[px4.y], 1); atomicAdd(&cnt[px4.z], 1); atomicAdd(&cnt[px4.w], 1);
      cost = 0.5 * e * e;
      num = 1.0;
    }
    e_out[m] = e;
    dp_out[m] = make_double2(dx, dy);
    pm_out[m] = make_double2(pcx, pcy);
    axy_out[m] = make_double2(ax, ay);
    pix_out[m] = px4;
    cp_out[m] = make_int2(sc, sq);
    flag_out[m] = fl;
    double2* Jo = reinterpret_cast<double2*>(J_out + (size_t)m * 24);
#pragma unroll
    for (int k = 0; k < 12; k++) Jo[k] = make_double2(Jrow[2 * k], Jrow[2 * k + 1]);
  }
  __shared__ double s_c[kExtThreads / 32], s_n[kExtThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { cost += __shfl_down_sync(0xffffffffu, cost, o); num += __shfl_down_sync(0xffffffffu, num, o); }
  if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = cost; s_n[threadIdx.x >> 5] = num; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = 0, k = 0;
#pragma unroll
    for (int w = 0; w < kExtThreads / 32; w++) { c += s_c[w]; k += s_n[w]; }
    part[2 * blockIdx.x] = c; part[2 * blockIdx.x + 1] = k;
  }
}

__global__ void k_ext_pairflag(const int32_t* __restrict__ prev, int64_t N, int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) flag[i] = prev[i] >= 0 ? 1 : 0;
}
__global__ void k_ext_compact(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int64_t N, uint32_t* __restrict__ mev) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N && flag[i]) mev[pos[i]] = (uint32_t)i;
}
__global__ void k_ext_active(const int32_t* __restrict__ cnt, int64_t P, int thres, int32_t* __restrict__ flag) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) flag[p] = cnt[p] >= thres ? 1 : 0;
}
__global__ void k_ext_amap(const int32_t* __restrict__ flag, const int32_t* __restrict__ aidx, int64_t P, int32_t* __restrict__ amap,
                           int32_t* __restrict__ apix, int64_t* __restrict__ tot) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  if (flag[p]) { amap[p] = aidx[p]; apix[aidx[p]] = (int32_t)p; } else amap[p] = -1;
  if (p == P - 1) tot[0] = (int64_t)aidx[p] + flag[p];
}

// g = J^T e and the block diagonal of J^T J. The current-event side of a warp's 32 time-consecutive rows almost
// always touches the same 4 control poses: those sums go through warp shuffles and leave as ONE atomic per entry and
// warp; everything else (previous-event side, map side) is added with fp64 atomics per row.
__global__ void __launch_bounds__(kExtThreads)
k_ext_form(int64_t Mt, const int32_t* __restrict__ amap, const double* __restrict__ e_in, const double2* __restrict__ dp_in,
           const double2* __restrict__ axy_in, const int4* __restrict__ pix_in, const int2* __restrict__ cp_in,
           const double* __restrict__ J_in, int32_t* __restrict__ flag, int n, double* __restrict__ g, double* __restrict__ Bp,
           double* __restrict__ Bm, int4* __restrict__ apix4, unsigned long long* __restrict__ used_count) {
  const int64_t m = (int64_t)blockIdx.x * kExtThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool used = false;
  int2 cp = make_int2(-1, -1);
  int4 a4 = make_int4(-1, -1, -1, -1);
  double J[24];
  double e = 0.0;
  if (m < Mt && flag[m]) {
    const int4 p4 = pix_in[m];
    a4 = make_int4(amap[p4.x], amap[p4.y], amap[p4.z], amap[p4.w]);
    used = a4.x >= 0 && a4.y >= 0 && a4.z >= 0 && a4.w >= 0;
    flag[m] = used ? 3 : 1;
    apix4[m] = a4;
  }
  if (used) {
    cp = cp_in[m];
    e = e_in[m];
    const double2* Ji = reinterpret_cast<const double2*>(J_in + (size_t)m * 24);
#pragma unroll
    for (int k = 0; k < 12; k++) { const double2 v = Ji[k]; J[2 * k] = v.x; J[2 * k + 1] = v.y; }
  } else {
#pragma unroll
    for (int k = 0; k < 24; k++) J[k] = 0.0;
  }
  const unsigned um = __ballot_sync(0xffffffffu, used);
  if (lane == 0 && um) atomicAdd(used_count, (unsigned long long)__popc(um));
  // ---- current-event side: warp-uniform control poses?
  const int key = used ? cp.x : -1;
  const int lead = um ? __ffs(um) - 1 : 0;
  const int key0 = __shfl_sync(0xffffffffu, key, lead);
  const bool uniform = __all_sync(0xffffffffu, !used || key == key0);
  if (um && uniform) {
    // 12 gradient entries + 4 x 6 block entries, reduced over the warp (unused lanes hold zeros)
    double acc[36];
#pragma unroll
    for (int k = 0; k < 12; k++) acc[k] = J[k] * e;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const double a = J[3 * q], b = J[3 * q + 1], c = J[3 * q + 2];
      acc[12 + 6 * q] = a * a; acc[13 + 6 * q] = a * b; acc[14 + 6 * q] = a * c;
      acc[15 + 6 * q] = b * b; acc[16 + 6 * q] = b * c; acc[17 + 6 * q] = c * c;
    }
#pragma unroll
    for (int k = 0; k < 36; k++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    // lane k (< 36 over two rounds) adds entry k
#pragma unroll
    for (int k = 0; k < 36; k++) {
      if (lane == (k & 31)) {
        if (k < 12) atomicAdd(&g[3 * key0 + k], acc[k]);
        else {
          const int q = (k - 12) / 6, t = (k - 12) % 6;
          const int r = t < 3 ? 0 : (t < 5 ? 1 : 2), c = t < 3 ? t : (t < 5 ? t - 2 : 2);
          atomicAdd(&Bp[(size_t)(key0 + q) * 9 + 3 * r + c], acc[k]);
          if (r != c) atomicAdd(&Bp[(size_t)(key0 + q) * 9 + 3 * c + r], acc[k]);
        }
      }
    }
  } else if (used) {
#pragma unroll
    for (int k = 0; k < 12; k++) atomicAdd(&g[3 * cp.x + k], J[k] * e);
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) atomicAdd(&Bp[(size_t)(cp.x + q) * 9 + 3 * r + c], J[3 * q + r] * J[3 * q + c]);
  }
  if (!used) return;
  // ---- previous-event side. A control pose touched by BOTH sides gets the cross terms too: its diagonal block is
  // (Jc_k + Jp_k)^T (Jc_k + Jp_k).
#pragma unroll
  for (int k = 0; k < 12; k++) atomicAdd(&g[3 * cp.y + k], J[12 + k] * e);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int pose = cp.y + q;
    const int qc = pose - cp.x;  // index of the same pose on the current side, if 0 <= qc < 4
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        double v = J[12 + 3 * q + r] * J[12 + 3 * q + c];
        if (qc >= 0 && qc < 4) {
          double jr = 0, jc = 0;
#pragma unroll
          for (int z = 0; z < 4; z++) if (z == qc) { jr = J[3 * z + r]; jc = J[3 * z + c]; }
          v += jr * J[12 + 3 * q + c] + J[12 + 3 * q + r] * jc;
        }
        atomicAdd(&Bp[(size_t)pose * 9 + 3 * r + c], v);
      }
  }
  // ---- map side: J_m = w_k * dp on pixel k of the footprint (clamped rows may repeat a pixel: the atomics add up,
  // the diagonal block then misses the repeated pixel's cross term, which only weakens the preconditioner there)
  const double2 dp = dp_in[m];
  const double2 a = axy_in[m];
  const double w[4] = {(1 - a.x) * (1 - a.y), a.x * (1 - a.y), (1 - a.x) * a.y, a.x * a.y};
  const int ai[4] = {a4.x, a4.y, a4.z, a4.w};
  const size_t off = (size_t)3 * n;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double jx = w[k] * dp.x, jy = w[k] * dp.y;
    atomicAdd(&g[off + 2 * (size_t)ai[k]], jx * e);
    atomicAdd(&g[off + 2 * (size_t)ai[k] + 1], jy * e);
    atomicAdd(&Bm[3 * (size_t)ai[k]], jx * jx);
    atomicAdd(&Bm[3 * (size_t)ai[k] + 1], jx * jy);
    atomicAdd(&Bm[3 * (size_t)ai[k] + 2], jy * jy);
  }
}

// applyL2Reg analogue (model.cpp:689-719): map block += alpha I, g_map -= alpha G
__global__ void k_ext_reg(int64_t Np, int n, const int32_t* __restrict__ apix, const double* __restrict__ Gx,
                          const double* __restrict__ Gy, double alpha, double* __restrict__ g, double* __restrict__ Bm) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= Np) return;
  const int32_t p = apix[a];
  Bm[3 * a] += alpha; Bm[3 * a + 2] += alpha;
  g[3 * (size_t)n + 2 * a] -= alpha * Gx[p];
  g[3 * (size_t)n + 2 * a + 1] -= alpha * Gy[p];
}

// t = J v (one thread per used row)
__global__ void __launch_bounds__(256)
k_ext_Jv(int64_t Mt, const int32_t* __restrict__ flag, const int2* __restrict__ cp_in, const int4* __restrict__ apix4,
         const double2* __restrict__ dp_in, const double2* __restrict__ axy_in, const double* __restrict__ J_in, int n,
         const double* __restrict__ v, double* __restrict__ t) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mt) return;
  if (flag[m] != 3) { t[m] = 0.0; return; }
  const int2 cp = cp_in[m];
  const double2* Ji = reinterpret_cast<const double2*>(J_in + (size_t)m * 24);
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 6; k++) { const double2 j = Ji[k]; s += j.x * v[3 * cp.x + 2 * k] + j.y * v[3 * cp.x + 2 * k + 1]; }
#pragma unroll
  for (int k = 0; k < 6; k++) { const double2 j = Ji[6 + k]; s += j.x * v[3 * cp.y + 2 * k] + j.y * v[3 * cp.y + 2 * k + 1]; }
  const double2 dp = dp_in[m];
  const double2 a = axy_in[m];
  const int4 a4 = apix4[m];
  const double w[4] = {(1 - a.x) * (1 - a.y), a.x * (1 - a.y), (1 - a.x) * a.y, a.x * a.y};
  const int ai[4] = {a4.x, a4.y, a4.z, a4.w};
  const double* vm = v + (size_t)3 * n;
#pragma unroll
  for (int k = 0; k < 4; k++) s += w[k] * (dp.x * vm[2 * (size_t)ai[k]] + dp.y * vm[2 * (size_t)ai[k] + 1]);
  t[m] = s;
}
// y += J^T t
__global__ void __launch_bounds__(256)
k_ext_JTt(int64_t Mt, const int32_t* __restrict__ flag, const int2* __restrict__ cp_in, const int4* __restrict__ apix4,
          const double2* __restrict__ dp_in, const double2* __restrict__ axy_in, const double* __restrict__ J_in, int n,
          const double* __restrict__ t, double* __restrict__ y) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mt || flag[m] != 3) return;
  const double s = t[m];
  const int2 cp = cp_in[m];
  const double2* Ji = reinterpret_cast<const double2*>(J_in + (size_t)m * 24);
#pragma unroll
  for (int k = 0; k < 6; k++) { const double2 j = Ji[k]; atomicAdd(&y[3 * cp.x + 2 * k], j.x * s); atomicAdd(&y[3 * cp.x + 2 * k + 1], j.y * s); }
#pragma unroll
  for (int k = 0; k < 6; k++) { const double2 j = Ji[6 + k]; atomicAdd(&y[3 * cp.y + 2 * k], j.x * s); atomicAdd(&y[3 * cp.y + 2 * k + 1], j.y * s); }
  const double2 dp = dp_in[m];
  const double2 a = axy_in[m];
  const int4 a4 = apix4[m];
  const double w[4] = {(1 - a.x) * (1 - a.y), a.x * (1 - a.y), (1 - a.x) * a.y, a.x * a.y};
  const int ai[4] = {a4.x, a4.y, a4.z, a4.w};
  double* ym = y + (size_t)3 * n;
#pragma unroll
  for (int k = 0; k < 4; k++) { atomicAdd(&ym[2 * (size_t)ai[k]], w[k] * dp.x * s); atomicAdd(&ym[2 * (size_t)ai[k] + 1], w[k] * dp.y * s); }
}
// y = alpha_map v (map part) + lambda * diag(H) .* v ; diag from the assembled blocks (map blocks include alpha)
__global__ void k_ext_diag_part(int n, int64_t Np, const double* __restrict__ Bp, const double* __restrict__ Bm, double lambda,
                                double alpha, const double* __restrict__ v, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t d = 3 * (int64_t)n + 2 * Np;
  if (i >= d) return;
  double dg, extra = 0.0;
  if (i < 3 * n) dg = Bp[(size_t)(i / 3) * 9 + 4 * (i % 3)];
  else { const int64_t k = i - 3 * n; dg = Bm[3 * (k >> 1) + ((k & 1) ? 2 : 0)]; extra = alpha; }
  y[i] = (extra + lambda * dg) * v[i];
}
// z = M^-1 r with the damped diagonal blocks (3x3 per pose, 2x2 per pixel); a singular block (untouched pose) -> z = r
__global__ void k_ext_precond(int n, int64_t Np, const double* __restrict__ Bp, const double* __restrict__ Bm, double lambda,
                              const double* __restrict__ r, double* __restrict__ z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double a[9];
#pragma unroll
    for (int k = 0; k < 9; k++) a[k] = Bp[(size_t)i * 9 + k];
    a[0] += lambda * a[0]; a[4] += lambda * a[4]; a[8] += lambda * a[8];
    const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
    const double r0 = r[3 * i], r1 = r[3 * i + 1], r2 = r[3 * i + 2];
    if (fabs(det) > 1e-300) {
      const double id = 1.0 / det;
      z[3 * i] = id * (c0 * r0 + (a[2] * a[7] - a[1] * a[8]) * r1 + (a[1] * a[5] - a[2] * a[4]) * r2);
      z[3 * i + 1] = id * (c1 * r0 + (a[0] * a[8] - a[2] * a[6]) * r1 + (a[2] * a[3] - a[0] * a[5]) * r2);
      z[3 * i + 2] = id * (c2 * r0 + (a[1] * a[6] - a[0] * a[7]) * r1 + (a[0] * a[4] - a[1] * a[3]) * r2);
    } else { z[3 * i] = r0; z[3 * i + 1] = r1; z[3 * i + 2] = r2; }
  } else if (i < n + Np) {
    const int64_t a = i - n;
    const double xx = Bm[3 * a] * (1.0 + lambda), xy = Bm[3 * a + 1], yy = Bm[3 * a + 2] * (1.0 + lambda);
    const double det = xx * yy - xy * xy;
    const double r0 = r[3 * (int64_t)n + 2 * a], r1 = r[3 * (int64_t)n + 2 * a + 1];
    if (fabs(det) > 1e-300) { z[3 * (int64_t)n + 2 * a] = (yy * r0 - xy * r1) / det; z[3 * (int64_t)n + 2 * a + 1] = (xx * r1 - xy * r0) / det; }
    else { z[3 * (int64_t)n + 2 * a] = r0; z[3 * (int64_t)n + 2 * a + 1] = r1; }
  }
}
__global__ void __launch_bounds__(256) k_ext_dot(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ part) {
  double s = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) s += a[i] * b[i];
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void k_ext_axpy(int64_t n, double a, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += a * x[i];
}
__global__ void k_ext_xpby(int64_t n, const double* __restrict__ x, double b, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] + b * y[i];
}
// state update: R_i <- Exp(x1_i) R_i (left perturbation, trajectory.cpp:296-304); map: active += damping * x2,
// inactive <- 0 (model.cpp:863-903)
__global__ void k_ext_apply_q(int n, const double* __restrict__ x, double* __restrict__ quat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 q = reinterpret_cast<const double4*>(quat)[i];
  const Vec3 w = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
  double4 r = quat_mul(so3_exp(w), q);
  const double nrm = sqrt(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
  r.x /= nrm; r.y /= nrm; r.z /= nrm; r.w /= nrm;
  reinterpret_cast<double4*>(quat)[i] = r;
}
__global__ void k_ext_apply_m(int64_t P, int n, const int32_t* __restrict__ amap, const double* __restrict__ x, double damping,
                              double* __restrict__ Gx, double* __restrict__ Gy) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int32_t a = amap[p];
  if (a >= 0) { Gx[p] += damping * x[3 * (int64_t)n + 2 * a]; Gy[p] += damping * x[3 * (int64_t)n + 2 * a + 1]; }
  else { Gx[p] = 0.0; Gy[p] = 0.0; }
}
__global__ void k_ext_sumsq(int64_t P, const double* __restrict__ Gx, const double* __restrict__ Gy, double* __restrict__ part) {
  double s = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < P; i += (int64_t)gridDim.x * 256) s += Gx[i] * Gx[i] + Gy[i] * Gy[i];
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

static void ext_free(ExtHandle* x) {
  void* ptrs[] = {x->d_quat, x->d_Gx, x->d_Gy, x->d_knot, x->d_mev, x->d_e, x->d_dp, x->d_pm, x->d_axy, x->d_pix, x->d_cp,
                  x->d_J, x->d_flag, x->d_cnt, x->d_amap, x->d_apix, x->d_g, x->d_Bp, x->d_Bm, x->d_vec, x->d_tmp, x->d_x,
                  x->d_part, x->d_scan, x->d_tot};
  for (void* p : ptrs) if (p) cudaFree(p);
}

static int ext_dot(ExtHandle* x, int64_t n, const double* a, const double* b, double* out) {
  Handle* h = x->h;
  const int grid = 296;
  k_ext_dot<<<grid, 256, 0, h->stream>>>(n, a, b, x->d_part);
  EMBA_LAUNCH_CHECK();
  double hp[296];
  EMBA_CUDA(cudaMemcpyAsync(hp, x->d_part, sizeof(double) * grid, cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  double s = 0;
  for (int i = 0; i < grid; i++) s += hp[i];
  *out = s;
  return EMBA_OK;
}

// y = (J^T J + alpha I_map + lambda diag(H)) v
static int ext_matvec(ExtHandle* x, double lambda, double alpha, const double* v, double* y) {
  Handle* h = x->h;
  const int64_t d = 3 * (int64_t)x->n + 2 * x->Np;
  k_ext_diag_part<<<ceil_div64(d, 256), 256, 0, h->stream>>>(x->n, x->Np, x->d_Bp, x->d_Bm, lambda, alpha, v, y);
  EMBA_LAUNCH_CHECK();
  if (x->Mt > 0) {
    k_ext_Jv<<<ceil_div64(x->Mt, 256), 256, 0, h->stream>>>(x->Mt, x->d_flag, x->d_cp, reinterpret_cast<int4*>(x->d_pm + x->Mt),
                                                           x->d_dp, x->d_axy, x->d_J, x->n, v, x->d_tmp);
    EMBA_LAUNCH_CHECK();
    k_ext_JTt<<<ceil_div64(x->Mt, 256), 256, 0, h->stream>>>(x->Mt, x->d_flag, x->d_cp, reinterpret_cast<int4*>(x->d_pm + x->Mt),
                                                            x->d_dp, x->d_axy, x->d_J, x->n, x->d_tmp, y);
    EMBA_LAUNCH_CHECK();
  }
  return EMBA_OK;
}

}  // namespace emba

using namespace emba;

#define XH() ExtHandle* x = (ExtHandle*)xx; if (!x) return EMBA_E_ARG; Handle* h = x->h; EMBA_CUDA(cudaSetDevice(h->device))

extern "C" {

int emba_ext_create(emba_handle_t hh, emba_events_t ee, int64_t i0, int64_t i1, emba_ext_t* out) {
  Handle* h = (Handle*)hh;
  EventStore* ev = (EventStore*)ee;
  if (!h || !out) return EMBA_E_ARG;
  *out = nullptr;
  EMBA_TRY(emba_set_events_dev(hh, ee, i0, i1));  // pair links (EventMap), sensor pixel and polarity per event
  ExtHandle* x = new ExtHandle();
  x->h = h;
  x->d_t = ev->t + i0;
  x->Mt = h->Mc_total;
  const int64_t Mt = std::max<int64_t>(x->Mt, 1), P = h->P, Nu = std::max<int64_t>(h->Nuse, 1);
  bool ok = cudaMalloc((void**)&x->d_Gx, 8 * P) == cudaSuccess && cudaMalloc((void**)&x->d_Gy, 8 * P) == cudaSuccess &&
            cudaMalloc((void**)&x->d_mev, 4 * Mt) == cudaSuccess && cudaMalloc((void**)&x->d_e, 8 * Mt) == cudaSuccess &&
            cudaMalloc((void**)&x->d_dp, 16 * Mt) == cudaSuccess && cudaMalloc((void**)&x->d_pm, 32 * Mt) == cudaSuccess &&
            cudaMalloc((void**)&x->d_axy, 16 * Mt) == cudaSuccess && cudaMalloc((void**)&x->d_pix, 16 * Mt) == cudaSuccess &&
            cudaMalloc((void**)&x->d_cp, 8 * Mt) == cudaSuccess && cudaMalloc((void**)&x->d_J, 192 * Mt) == cudaSuccess &&
            cudaMalloc((void**)&x->d_flag, 4 * std::max(Mt, Nu)) == cudaSuccess && cudaMalloc((void**)&x->d_cnt, 4 * (P + 1)) == cudaSuccess &&
            cudaMalloc((void**)&x->d_amap, 4 * (P + 1)) == cudaSuccess && cudaMalloc((void**)&x->d_apix, 4 * (P + 1)) == cudaSuccess &&
            cudaMalloc((void**)&x->d_Bm, 24 * (P + 1)) == cudaSuccess && cudaMalloc((void**)&x->d_tmp, 8 * Mt) == cudaSuccess &&
            cudaMalloc((void**)&x->d_part, 8 * 2 * (size_t)(ceil_div64(Mt, kExtThreads) + 512)) == cudaSuccess &&
            cudaMalloc(&x->d_scan, scan_scratch_bytes(std::max(Nu, P) + 2)) == cudaSuccess &&
            cudaMalloc((void**)&x->d_tot, 64) == cudaSuccess;
  // (d_pm holds the warped positions [Mt] followed by the footprint's active indices [Mt] as int4)
  if (!ok) { cudaGetLastError(); ext_free(x); delete x; h->err = "emba_ext_create: out of device memory"; return EMBA_E_CUDA; }
  if (x->Mt > 0) {
    // pairs in time order of their current event
    int32_t* pos = x->d_amap;  // scratch (P + 1 >= ? no: sized by events) -> use a temporary
    int32_t* tmp_pos = nullptr;
    if (cudaMalloc((void**)&tmp_pos, 4 * Nu) != cudaSuccess) { cudaGetLastError(); ext_free(x); delete x; return EMBA_E_CUDA; }
    (void)pos;
    k_ext_pairflag<<<ceil_div64(h->Nuse, 256), 256, 0, h->stream>>>(h->d_prev, h->Nuse, x->d_flag);
    int rc = scan_exclusive<int32_t>(h, h->stream, x->d_flag, tmp_pos, h->Nuse, x->d_scan);
    if (rc == EMBA_OK) k_ext_compact<<<ceil_div64(h->Nuse, 256), 256, 0, h->stream>>>(x->d_flag, tmp_pos, h->Nuse, x->d_mev);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp_pos);
    if (rc != EMBA_OK || e != cudaSuccess) { ext_free(x); delete x; return rc != EMBA_OK ? rc : EMBA_E_CUDA; }
  }
  *out = (emba_ext_t)x;
  return EMBA_OK;
}

int emba_ext_destroy(emba_ext_t xx) {
  ExtHandle* x = (ExtHandle*)xx;
  if (!x) return EMBA_OK;
  cudaSetDevice(x->h->device);
  cudaStreamSynchronize(x->h->stream);
  ext_free(x);
  delete x;
  return EMBA_OK;
}

int emba_ext_num_pairs(emba_ext_t xx, int64_t* out) {
  ExtHandle* x = (ExtHandle*)xx;
  if (!x || !out) return EMBA_E_ARG;
  *out = x->Mt;
  return EMBA_OK;
}

int emba_ext_set_state(emba_ext_t xx, int64_t t0_ns, int64_t dt_ns, int32_t n, const double* quat, const double* Gx,
                       const double* Gy) {
  XH();
  if (!quat || !Gx || !Gy || n < 4 || n > 65535 || dt_ns <= 0) { h->err = "emba_ext_set_state: bad argument (a cubic spline needs >= 4 control poses)"; return EMBA_E_ARG; }
  if (n != x->n) {
    if (x->d_quat) cudaFree(x->d_quat);
    if (x->d_knot) cudaFree(x->d_knot);
    if (x->d_Bp) cudaFree(x->d_Bp);
    if (x->d_g) cudaFree(x->d_g);
    if (x->d_x) cudaFree(x->d_x);
    x->d_quat = x->d_knot = x->d_Bp = x->d_g = x->d_x = nullptr;
    const size_t d = 3 * (size_t)n + 2 * (size_t)(h->P + 1);
    if (cudaMalloc((void**)&x->d_quat, 32 * (size_t)n) != cudaSuccess || cudaMalloc((void**)&x->d_knot, 8 * kExtKnot * (size_t)n) != cudaSuccess ||
        cudaMalloc((void**)&x->d_Bp, 72 * (size_t)n) != cudaSuccess || cudaMalloc((void**)&x->d_g, 8 * d) != cudaSuccess ||
        cudaMalloc((void**)&x->d_x, 8 * d) != cudaSuccess) {
      cudaGetLastError(); x->n = 0; h->err = "emba_ext_set_state: out of device memory"; return EMBA_E_CUDA;
    }
    x->n = n;
  }
  x->t0_ns = t0_ns; x->dt_ns = dt_ns;
  EMBA_CUDA(upload_bytes(h->up, h->stream, x->d_quat, quat, 32 * (size_t)n));
  EMBA_CUDA(upload_bytes(h->up, h->stream, x->d_Gx, Gx, 8 * (size_t)h->P));
  EMBA_CUDA(upload_bytes(h->up, h->stream, x->d_Gy, Gy, 8 * (size_t)h->P));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  x->evaluated = x->formed = x->solved = false;
  return EMBA_OK;
}

int emba_ext_get_state(emba_ext_t xx, double* quat, double* Gx, double* Gy) {
  XH();
  if (x->n <= 0) return EMBA_E_ARG;
  if (quat) EMBA_CUDA(download_bytes(h->up, h->stream, quat, x->d_quat, 32 * (size_t)x->n));
  if (Gx) EMBA_CUDA(download_bytes(h->up, h->stream, Gx, x->d_Gx, 8 * (size_t)h->P));
  if (Gy) EMBA_CUDA(download_bytes(h->up, h->stream, Gy, x->d_Gy, 8 * (size_t)h->P));
  return EMBA_OK;
}

int emba_ext_evaluate(emba_ext_t xx, double alpha, double* cost_data, double* cost_reg, int64_t* M) {
  XH();
  if (x->n <= 0) { h->err = "emba_ext_evaluate: set the state first"; return EMBA_E_ARG; }
  const PanoCam cam = make_cam(h);
  EMBA_CUDA(cudaEventRecord(h->ev[0], h->stream));
  k_ext_knots<<<ceil_div64(x->n, 64), 64, 0, h->stream>>>(x->d_quat, x->n, x->d_knot);
  EMBA_LAUNCH_CHECK();
  EMBA_CUDA(cudaMemsetAsync(x->d_cnt, 0, 4 * (size_t)h->P, h->stream));
  EMBA_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(int32_t) * 16, h->stream));
  const int grid = ceil_div64(std::max<int64_t>(x->Mt, 1), kExtThreads);
  k_ext_eval<<<grid, kExtThreads, 0, h->stream>>>(x->Mt, x->d_mev, h->d_prev, x->d_t, h->d_spix_ev, h->d_pol, h->d_lut, x->d_knot,
                                                  x->n, x->t0_ns, x->dt_ns, x->d_Gx, x->d_Gy, cam, h->Wp, h->Hp, h->C_th, x->d_e,
                                                  x->d_dp, x->d_pm, x->d_axy, x->d_pix, x->d_cp, x->d_J, x->d_flag, x->d_cnt,
                                                  x->d_part, h->d_flags);
  EMBA_LAUNCH_CHECK();
  const int rg = 256;
  k_ext_sumsq<<<rg, 256, 0, h->stream>>>(h->P, x->d_Gx, x->d_Gy, x->d_part + 2 * (size_t)grid);
  EMBA_LAUNCH_CHECK();
  EMBA_CUDA(cudaEventRecord(h->ev[1], h->stream));
  std::vector<double> hp(2 * (size_t)grid + rg);
  int32_t fl = 0;
  EMBA_CUDA(cudaMemcpyAsync(hp.data(), x->d_part, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaMemcpyAsync(&fl, h->d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  x->t_ms[0] = ms;
  if (fl & 2) { h->err = "event time outside the cubic spline's support (basalt asserts here: so3_spline.h:221-230)"; return EMBA_E_SUPPORT; }
  double c = 0, k = 0, r = 0;
  for (int b = 0; b < grid; b++) { c += hp[2 * b]; k += hp[2 * b + 1]; }
  for (int b = 0; b < rg; b++) r += hp[2 * (size_t)grid + b];
  x->M = (int64_t)llround(k);
  if (cost_data) *cost_data = c;
  if (cost_reg) *cost_reg = 0.5 * alpha * r;
  if (M) *M = x->M;
  x->evaluated = true;
  x->formed = x->solved = false;
  return EMBA_OK;
}

int emba_ext_form(emba_ext_t xx, int32_t thres, double alpha, int64_t* Np_out, int64_t* Mused_out) {
  XH();
  if (!x->evaluated) { h->err = "emba_ext_form: evaluate first"; return EMBA_E_ARG; }
  const int64_t P = h->P;
  const int n = x->n;
  EMBA_CUDA(cudaEventRecord(h->ev[0], h->stream));
  int32_t* d_flag = h->d_pflag;
  int32_t* d_aidx = h->d_paidx;
  k_ext_active<<<ceil_div64(P, 256), 256, 0, h->stream>>>(x->d_cnt, P, thres, d_flag);
  EMBA_LAUNCH_CHECK();
  EMBA_TRY(scan_exclusive<int32_t>(h, h->stream, d_flag, d_aidx, P, x->d_scan));
  k_ext_amap<<<ceil_div64(P, 256), 256, 0, h->stream>>>(d_flag, d_aidx, P, x->d_amap, x->d_apix, x->d_tot);
  EMBA_LAUNCH_CHECK();
  EMBA_CUDA(cudaMemcpyAsync(h->h_pin + 40, x->d_tot, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  x->Np = h->h_pin[40];
  const int64_t d = 3 * (int64_t)n + 2 * x->Np;
  EMBA_CUDA(cudaMemsetAsync(x->d_g, 0, 8 * (size_t)d, h->stream));
  EMBA_CUDA(cudaMemsetAsync(x->d_Bp, 0, 72 * (size_t)n, h->stream));
  EMBA_CUDA(cudaMemsetAsync(x->d_Bm, 0, 24 * (size_t)(x->Np + 1), h->stream));
  EMBA_CUDA(cudaMemsetAsync(x->d_tot + 1, 0, sizeof(int64_t), h->stream));
  if (x->Mt > 0) {
    k_ext_form<<<ceil_div64(x->Mt, kExtThreads), kExtThreads, 0, h->stream>>>(
        x->Mt, x->d_amap, x->d_e, x->d_dp, x->d_axy, x->d_pix, x->d_cp, x->d_J, x->d_flag, n, x->d_g, x->d_Bp, x->d_Bm,
        reinterpret_cast<int4*>(x->d_pm + x->Mt), reinterpret_cast<unsigned long long*>(x->d_tot + 1));
    EMBA_LAUNCH_CHECK();
  }
  if (x->Np > 0) {
    k_ext_reg<<<ceil_div64(x->Np, 256), 256, 0, h->stream>>>(x->Np, n, x->d_apix, x->d_Gx, x->d_Gy, alpha, x->d_g, x->d_Bm);
    EMBA_LAUNCH_CHECK();
  }
  EMBA_CUDA(cudaEventRecord(h->ev[1], h->stream));
  EMBA_CUDA(cudaMemcpyAsync(h->h_pin + 41, x->d_tot + 1, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  x->Mused = h->h_pin[41];
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  x->t_ms[1] = ms;
  if (Np_out) *Np_out = x->Np;
  if (Mused_out) *Mused_out = x->Mused;
  x->formed = true;
  x->solved = false;
  return EMBA_OK;
}

int emba_ext_get_rows(emba_ext_t xx, int64_t cap, double* e, double* dp, double* pm, double* J24, double* axy, int32_t* cp,
                      int32_t* pix, int32_t* ev, int32_t* flag, int64_t* n_rows) {
  XH();
  if (!x->evaluated || !n_rows) return EMBA_E_ARG;
  *n_rows = x->Mt;
  if (cap < x->Mt) { h->err = "emba_ext_get_rows: capacity too small"; return EMBA_E_ARG; }
  const size_t M = (size_t)x->Mt;
  if (M == 0) return EMBA_OK;
  if (e) EMBA_CUDA(download_bytes(h->up, h->stream, e, x->d_e, 8 * M));
  if (dp) EMBA_CUDA(download_bytes(h->up, h->stream, dp, x->d_dp, 16 * M));
  if (pm) EMBA_CUDA(download_bytes(h->up, h->stream, pm, x->d_pm, 16 * M));
  if (J24) EMBA_CUDA(download_bytes(h->up, h->stream, J24, x->d_J, 192 * M));
  if (axy) EMBA_CUDA(download_bytes(h->up, h->stream, axy, x->d_axy, 16 * M));
  if (cp) EMBA_CUDA(download_bytes(h->up, h->stream, cp, x->d_cp, 8 * M));
  if (pix) EMBA_CUDA(download_bytes(h->up, h->stream, pix, x->d_pix, 16 * M));
  if (ev) EMBA_CUDA(download_bytes(h->up, h->stream, ev, x->d_mev, 4 * M));
  if (flag) EMBA_CUDA(download_bytes(h->up, h->stream, flag, x->d_flag, 4 * M));
  return EMBA_OK;
}

int emba_ext_get_normal_eq(emba_ext_t xx, double* g, double* pose_blocks, double* pixel_blocks, int32_t* active) {
  XH();
  if (!x->formed) { h->err = "emba_ext_get_normal_eq: form first"; return EMBA_E_ARG; }
  const size_t d = 3 * (size_t)x->n + 2 * (size_t)x->Np;
  if (g) EMBA_CUDA(download_bytes(h->up, h->stream, g, x->d_g, 8 * d));
  if (pose_blocks) EMBA_CUDA(download_bytes(h->up, h->stream, pose_blocks, x->d_Bp, 72 * (size_t)x->n));
  if (pixel_blocks && x->Np) EMBA_CUDA(download_bytes(h->up, h->stream, pixel_blocks, x->d_Bm, 24 * (size_t)x->Np));
  if (active && x->Np) EMBA_CUDA(download_bytes(h->up, h->stream, active, x->d_apix, 4 * (size_t)x->Np));
  return EMBA_OK;
}

int emba_ext_matvec(emba_ext_t xx, double lambda, double alpha, const double* v, double* y) {
  XH();
  if (!x->formed || !v || !y) return EMBA_E_ARG;
  const int64_t d = 3 * (int64_t)x->n + 2 * x->Np;
  EMBA_TRY(dev_reserve(h, &x->d_vec, &x->vec_cap, 6 * d + 16));
  EMBA_CUDA(upload_bytes(h->up, h->stream, x->d_vec, v, 8 * (size_t)d));
  EMBA_TRY(ext_matvec(x, lambda, alpha, x->d_vec, x->d_vec + d));
  EMBA_CUDA(download_bytes(h->up, h->stream, y, x->d_vec + d, 8 * (size_t)d));
  return EMBA_OK;
}

// block-Jacobi PCG on (J^T J + alpha I_map + lambda diag H) x = g  (the counterpart of model.cpp:794-840; Eigen's
// loop, ConjugateGradient.h:26-96, with the block preconditioner the north star names)
int emba_ext_solve(emba_ext_t xx, double lambda, double alpha, int32_t max_iter, double tol, double* x_out, int32_t* iters,
                   double* err) {
  XH();
  if (!x->formed) { h->err = "emba_ext_solve: form first"; return EMBA_E_ARG; }
  const int64_t d = 3 * (int64_t)x->n + 2 * x->Np;
  const int T = 256, G = ceil_div64(d, T);
  const int nb = x->n + (int)std::min<int64_t>(x->Np, INT32_MAX - x->n);
  EMBA_TRY(dev_reserve(h, &x->d_vec, &x->vec_cap, 6 * d + 16));
  EMBA_CUDA(cudaEventRecord(h->ev[0], h->stream));
  double *sol = x->d_x, *r = x->d_vec, *p = r + d, *z = p + d, *tmp = z + d;
  EMBA_CUDA(cudaMemsetAsync(sol, 0, 8 * (size_t)d, h->stream));
  EMBA_CUDA(cudaMemcpyAsync(r, x->d_g, 8 * (size_t)d, cudaMemcpyDeviceToDevice, h->stream));
  double rhs2 = 0, rn2 = 0, absNew = 0;
  EMBA_TRY(ext_dot(x, d, r, r, &rhs2));
  int it = 0;
  rn2 = rhs2;
  const double thr = std::max(tol * tol * rhs2, 2.2250738585072014e-308);
  if (rhs2 > 0 && rn2 >= thr) {
    k_ext_precond<<<ceil_div64(nb, T), T, 0, h->stream>>>(x->n, x->Np, x->d_Bp, x->d_Bm, lambda, r, p);
    EMBA_LAUNCH_CHECK();
    EMBA_TRY(ext_dot(x, d, r, p, &absNew));
    while (it < max_iter) {
      EMBA_TRY(ext_matvec(x, lambda, alpha, p, tmp));
      double ptmp = 0;
      EMBA_TRY(ext_dot(x, d, p, tmp, &ptmp));
      const double a = absNew / ptmp;
      k_ext_axpy<<<G, T, 0, h->stream>>>(d, a, p, sol);
      k_ext_axpy<<<G, T, 0, h->stream>>>(d, -a, tmp, r);
      h->launches += 2;
      EMBA_TRY(ext_dot(x, d, r, r, &rn2));
      if (rn2 < thr) break;
      k_ext_precond<<<ceil_div64(nb, T), T, 0, h->stream>>>(x->n, x->Np, x->d_Bp, x->d_Bm, lambda, r, z);
      EMBA_LAUNCH_CHECK();
      const double absOld = absNew;
      EMBA_TRY(ext_dot(x, d, r, z, &absNew));
      k_ext_xpby<<<G, T, 0, h->stream>>>(d, z, absNew / absOld, p);
      EMBA_LAUNCH_CHECK();
      it++;
    }
  }
  EMBA_CUDA(cudaEventRecord(h->ev[1], h->stream));
  if (x_out) EMBA_CUDA(download_bytes(h->up, h->stream, x_out, sol, 8 * (size_t)d));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  x->t_ms[2] = ms;
  if (iters) *iters = it;
  if (err) *err = rhs2 > 0 ? sqrt(rn2 / rhs2) : 0.0;
  x->solved = true;
  return EMBA_OK;
}

int emba_ext_apply(emba_ext_t xx, double damping) {
  XH();
  if (!x->solved) { h->err = "emba_ext_apply: solve first"; return EMBA_E_ARG; }
  k_ext_apply_q<<<ceil_div64(x->n, 128), 128, 0, h->stream>>>(x->n, x->d_x, x->d_quat);
  EMBA_LAUNCH_CHECK();
  k_ext_apply_m<<<ceil_div64(h->P, 256), 256, 0, h->stream>>>(h->P, x->n, x->d_amap, x->d_x, damping, x->d_Gx, x->d_Gy);
  EMBA_LAUNCH_CHECK();
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  x->evaluated = x->formed = x->solved = false;
  return EMBA_OK;
}

int emba_ext_last_ms(emba_ext_t xx, double* out3) {
  ExtHandle* x = (ExtHandle*)xx;
  if (!x || !out3) return EMBA_E_ARG;
  out3[0] = x->t_ms[0]; out3[1] = x->t_ms[1]; out3[2] = x->t_ms[2];
  return EMBA_OK;
}

}  // extern "C"
