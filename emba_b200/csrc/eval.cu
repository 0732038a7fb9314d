// LEGM::evaluateDataError on the device (reference src/emba/model.cpp:72-258): second-order maps, per-batch
// spline poses, and the per-measurement residual kernel.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "emba_internal.cuh"

namespace emba {

int rebuild_static(Handle* h);
int comm_allreduce(Handle* h, void* buf, int64_t count, int dtype);  // dtype: 0 = int32, 1 = fp64
int comm_allreduce_eval(Handle* h, const int32_t* hist_loc, int32_t* hist, int64_t P, double* scal2, int32_t* flags);

// ---------------------------------------------------------------------------------------------------
// Second-order gradient maps (model.cpp:87-97): 0.125 * 3x3 Sobel (cv::Sobel defaults: scale 1,
// BORDER_REFLECT_101), Gxy = 0.5 * (dGx/dy + dGy/dx). Also interleaves (Gx, Gy) for single 16-byte gathers.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  if (p < 0) p = -p;
  if (p >= len) p = 2 * (len - 1) - p;
  return p;
}

__global__ void k_map_prepare(const double* __restrict__ Gx, const double* __restrict__ Gy, int W, int H,
                              double2* __restrict__ G2, double4* __restrict__ H3) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
  const int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H);
  const int64_t r0 = (int64_t)ym * W, r1 = (int64_t)y * W, r2 = (int64_t)yp * W;
  // d/dx: derivative [-1 0 1] along x, smoothing [1 2 1] along y; d/dy the transpose
  auto ddx = [&](const double* A) {
    return (A[r0 + xp] - A[r0 + xm]) + 2.0 * (A[r1 + xp] - A[r1 + xm]) + (A[r2 + xp] - A[r2 + xm]);
  };
  auto ddy = [&](const double* A) {
    return (A[r2 + xm] + 2.0 * A[r2 + x] + A[r2 + xp]) - (A[r0 + xm] + 2.0 * A[r0 + x] + A[r0 + xp]);
  };
  const double gxx = 0.125 * ddx(Gx);
  const double gxy = 0.125 * ddy(Gx);
  const double gyx = 0.125 * ddx(Gy);
  const double gyy = 0.125 * ddy(Gy);
  const int64_t i = r1 + x;
  G2[i] = make_double2(Gx[i], Gy[i]);
  H3[i] = make_double4(gxx, 0.5 * (gxy + gyx), gyy, 0.0);
}

// ---------------------------------------------------------------------------------------------------
// Per-batch pose and Jacobian factor (model.cpp:115-136 -> src/utils/trajectory.cpp:122-147 ->
// basalt So3Spline<2>::evaluate, so3_spline.h:218-274). For the linear spline
//   R = R_s * Exp(u * Log(R_s^-1 R_{s+1})),   d_val_d_knot = [I - A | A],
//   A = u * R_s * Jl(u delta) * Jl^-1(delta) * R_s^T          (so3_spline.h:254-265)
// ---------------------------------------------------------------------------------------------------
__global__ void k_knot_table(const double* __restrict__ quat, int n, double* __restrict__ Ktab) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n - 1) return;
  const double4 q0 = reinterpret_cast<const double4*>(quat)[s];
  const double4 q1 = reinterpret_cast<const double4*>(quat)[s + 1];
  const double4 q0inv = make_double4(-q0.x, -q0.y, -q0.z, q0.w);
  const Vec3 db = so3_log(quat_mul(q0inv, q1));  // delta = Log(R_s^-1 R_{s+1}) (so3_spline.h:249-250)
  const Mat3 R0 = quat_to_R(q0);
  double* o = Ktab + (size_t)s * kKnotStride;
#pragma unroll
  for (int i = 0; i < 9; i++) o[i] = R0.m[i];
  // world-frame increment delta_w = R_s delta: R_s Exp(u delta) = Exp(u delta_w) R_s
  o[9] = R0.m[0] * db.x + R0.m[1] * db.y + R0.m[2] * db.z;
  o[10] = R0.m[3] * db.x + R0.m[4] * db.y + R0.m[5] * db.z;
  o[11] = R0.m[6] * db.x + R0.m[7] * db.y + R0.m[8] * db.z;
}

// Per batch: Rodrigues scalars of Exp(u delta_w) and the three coefficients of
//   A = u Jl(u delta_w) Jl^-1(delta_w) = alpha I + beta K + gamma K^2
// (both Jacobians are polynomials in K = [delta_w]x, K^3 = -th^2 K), with the small-angle branches of
// leftJacobianSO3 / leftJacobianInvSO3 (sophus_utils.hpp:333-414).
__global__ void k_batch_table(const double* __restrict__ Ktab, const int32_t* __restrict__ bs,
                              const double* __restrict__ bu, int64_t B, double4* __restrict__ RotTab,
                              double4* __restrict__ JacTab) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int s = bs[b];
  const double u = bu[b];
  const double* kt = Ktab + (size_t)s * kKnotStride;
  const double th2 = kt[9] * kt[9] + kt[10] * kt[10] + kt[11] * kt[11];
  const double th = sqrt(th2);
  const double ut = u * th, ut2 = ut * ut;
  double s1, s2;
  if (th2 < kSophusEps * kSophusEps) {
    s1 = u * (1.0 - ut2 / 6.0);
    s2 = 0.5 * u * u * (1.0 - ut2 / 12.0);
  } else {
    s1 = sin(ut) / th;
    s2 = (1.0 - cos(ut)) / th2;
  }
  // Jl(u delta) = I + p K + q K^2
  double p, q;
  if (ut2 > kSophusEps) {
    p = u * (1.0 - cos(ut)) / ut2;
    q = u * u * (ut - sin(ut)) / (ut2 * ut);
  } else {
    p = 0.5 * u;
    q = u * u / 6.0;
  }
  // Jl^-1(delta) = I - K/2 + c K^2
  double c;
  if (th2 > kSophusEps) {
    if (th < 3.14159265358979323846 - 1e-5) c = 1.0 / th2 - (1.0 + cos(th)) / (2.0 * th * sin(th));
    else c = 1.0 / (3.14159265358979323846 * 3.14159265358979323846);
  } else {
    c = 1.0 / 12.0;
  }
  const double beta = p - 0.5 - th2 * (p * c - 0.5 * q);
  const double gamma = q + c - 0.5 * p - th2 * q * c;
  RotTab[b] = make_double4(s1, s2, __longlong_as_double((long long)s), 0.0);
  JacTab[b] = make_double4(u, u * beta, u * gamma, 0.0);
}

// ---------------------------------------------------------------------------------------------------
// The per-measurement residual kernel (model.cpp:140-246). One thread per event pair:
//   rb_c = R(batch_c) b, rb_p = R(batch_p) b      (same sensor pixel -> same bearing b)
//   pm = equirectangular(rb); dp = pm_c - pm_p; outlier iff |dp| > 10
//   pix = round(pm_c); e = +-C_th - G(pix).dp; num_ev_map[pix]++
// Outputs dp, e, pix per measurement; cost and inlier count as per-block partial sums (fixed grid ->
// deterministic summation order).
// ---------------------------------------------------------------------------------------------------
template <int COST>
__device__ __forceinline__ double rho_of(double e, double a) {
  if (COST == EMBA_COST_QUADRATIC) return 0.5 * e * e;
  if (COST == EMBA_COST_CAUCHY) return (0.5 / a) * log1p(a * e * e);  // model.cpp:283-290
  const double ab = fabs(e);                                           // model.cpp:294-312
  return ab < a ? 0.5 * ab * ab : a * ab - 0.5 * a * a;
}

// (Projecting every EVENT once into a 16-byte array and forming the pairs from it -- half the atan2 / asin -- was
// built and measured on C4: 1.47 ms projection + 2.81 ms pair kernel against 3.7 ms for this fused kernel. The pair
// kernel is bound by its gathers and the counting atomic, not by the fp64 pipe, so the split only added traffic.)
constexpr int kEvalThreads = 256;

// PF (software prefetch across the grid-stride loop; the kernel is bound by the latency of its dependent loads --
// record -> batch entries -> knot entries -> map gather, "long scoreboard" 7 of 12.8 stall cycles per issue):
//   bit 0: the record two iterations ahead and the two batch-table entries of the NEXT iteration's record are
//          prefetched into L1, bit 1: into L2 instead; bit 2: both batch-table entries are requested before the first
//          one is consumed (the 256-bit loads are volatile asm, which otherwise keeps them in source order, the second
//          behind the first's dependent knot loads).
// Measured on C4 / C2 (ms): PF 0: 3.58 / 0.276, 4: 3.21 / 0.262 (default), 5: 3.21 / 0.285, 6: 3.40 / 0.285,
// 1: 3.62 / 0.299. Two more steps on top of PF 4 were built and measured without effect (3.18 - 3.22 ms): the map
// gather requested right after the current event's projection (before the second projection), and the slot store
// (which waits for the counting atomic's return) held back into the next iteration behind its record load.
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int COST, int MINB, int PF>
__global__ void __launch_bounds__(kEvalThreads, MINB)
k_eval(const MeasRec* __restrict__ rec, int64_t Mc, const double* __restrict__ Ktab,
       const double4* __restrict__ RotTab, const double2* __restrict__ G2, PanoCam cam, int W, int H, double C_th, double eta,
       double2* __restrict__ dp_out, double* __restrict__ e_out, int32_t* __restrict__ pix_out,
       int32_t* __restrict__ slot_out, int32_t* __restrict__ hist, double* __restrict__ part,
       int32_t* __restrict__ flags) {
  double cost = 0.0;
  double cnt = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kEvalThreads;
  for (int64_t m = (int64_t)blockIdx.x * kEvalThreads + threadIdx.x; m < Mc; m += stride) {
    const double4 r0 = ldg256(rec + m);
    unsigned long long wn = 0ull;
    bool have_next = false;
    if (PF & 3) {
      if (m + 2 * stride < Mc) { if (PF & 1) prefetch_l1(rec + m + 2 * stride); else prefetch_l2(rec + m + 2 * stride); }
      have_next = m + stride < Mc;
      if (have_next) wn = __ldg(reinterpret_cast<const unsigned long long*>(rec + m + stride) + 3);
    }
    const double bx = r0.x, by = r0.y, bz = r0.z;
    const unsigned long long w = (unsigned long long)__double_as_longlong(r0.w);
    const uint32_t bcp = (uint32_t)w, bp = (uint32_t)(w >> 32);
    const uint32_t bc = bcp & 0x7FFFFFFFu;
    const double pol = (bcp >> 31) ? 1.0 : 0.0;
    double pcx, pcy, ppx, ppy;
    if (PF & 4) {
      const double4 rtc = ldg256(RotTab + bc);
      const double4 rtp = ldg256(RotTab + bp);
      double X, Y, Z;
      // ptxas schedules loads as it likes, whatever the source order: the (always zero) high bits of the second
      // entry's knot index enter the first one's knot address, so both entries are in flight before either is used
      const int skc = (int)__double_as_longlong(rtc.z) + (int)(__double_as_longlong(rtp.z) >> 40);
      rotate_bearing_v(Ktab + (size_t)skc * kKnotStride, rtc.x, rtc.y, bx, by, bz, X, Y, Z);
      project_pm_unit(cam, X, Y, Z, pcx, pcy);
      rotate_bearing_v(Ktab + (size_t)((int)__double_as_longlong(rtp.z)) * kKnotStride, rtp.x, rtp.y, bx, by, bz, X, Y, Z);
      project_pm_unit(cam, X, Y, Z, ppx, ppy);
    } else {
      {
        const double4 rt = ldg256(RotTab + bc);
        const int sk = (int)__double_as_longlong(rt.z);
        double X, Y, Z;
        rotate_bearing_v(Ktab + (size_t)sk * kKnotStride, rt.x, rt.y, bx, by, bz, X, Y, Z);
        project_pm_unit(cam, X, Y, Z, pcx, pcy);
      }
      {
        const double4 rt = ldg256(RotTab + bp);
        const int sk = (int)__double_as_longlong(rt.z);
        double X, Y, Z;
        rotate_bearing_v(Ktab + (size_t)sk * kKnotStride, rt.x, rt.y, bx, by, bz, X, Y, Z);
        project_pm_unit(cam, X, Y, Z, ppx, ppy);
      }
    }
    if ((PF & 3) && have_next) {
      const uint32_t nbc = (uint32_t)wn & 0x7FFFFFFFu, nbp = (uint32_t)(wn >> 32);
      if (PF & 1) { prefetch_l1(RotTab + nbc); prefetch_l1(RotTab + nbp); }
      else { prefetch_l2(RotTab + nbc); prefetch_l2(RotTab + nbp); }
    }
    const double dx = pcx - ppx, dy = pcy - ppy;
    // dp.norm() > 10 (model.cpp:199-200), evaluated without FMA contraction like the CPU build
    const double nrm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    int32_t pix = -1, slot = -1;
    double e = 0.0;
    if (!(nrm > 10.0)) {
      const int px = (int)round(pcx), py = (int)round(pcy);  // std::round, model.cpp:209-210
      // cv::Mat::at(py, px) of a continuous row-major matrix is element py*W + px: column W (a pixel within half a
      // pixel of the phi = +-pi seam) is element (py + 1, 0), which is what the reference's release build reads
      // and counts (model.cpp:213-227). Only an index past the last element is out of bounds there: such a pair is
      // dropped as an outlier (and reported, see emba_set_strict_range).
      const long long lin = (long long)py * W + px;
      if (px < 0 || px > W || py < 0 || lin >= (long long)W * H) {
        atomicOr(flags, 4);
      } else {
        pix = (int32_t)lin;
        const double2 g = G2[pix];
        const double C_pred = g.x * dx + g.y * dy;        // model.cpp:217
        const double C_meas = 2.0 * (pol - 0.5) * C_th;   // model.cpp:219
        e = C_meas - C_pred;
        slot = atomicAdd(&hist[pix], 1);                  // model.cpp:227; the old count is this row's slot in the pixel
        cost += rho_of<COST>(e, eta);
        cnt += 1.0;
      }
    }
    dp_out[m] = make_double2(dx, dy);
    e_out[m] = e;
    pix_out[m] = pix;
    slot_out[m] = slot;
  }
  // block reduction, fixed tree
  __shared__ double s_cost[kEvalThreads / 32], s_cnt[kEvalThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cost += __shfl_down_sync(0xffffffffu, cost, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { s_cost[threadIdx.x >> 5] = cost; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = 0, k = 0;
#pragma unroll
    for (int w = 0; w < kEvalThreads / 32; w++) { c += s_cost[w]; k += s_cnt[w]; }
    part[2 * blockIdx.x] = c;
    part[2 * blockIdx.x + 1] = k;
  }
}

// sum of squares of both maps over all pixels (evaluateRegError, model.cpp:260-277 + solver.cpp:90)
__global__ void __launch_bounds__(256) k_reg_partial(const double* __restrict__ Gx, const double* __restrict__ Gy,
                                                     int64_t P, double* __restrict__ part) {
  double s = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < P; i += (int64_t)gridDim.x * 256) {
    const double a = Gx[i], b = Gy[i];
    s += a * a + b * b;
  }
  __shared__ double sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += sh[w];
    part[blockIdx.x] = t;
  }
}

// final fixed-order sum of `nblk` partial rows of width `stride`; out[j] = sum_b part[b*stride + j]
__global__ void k_sum_partials(const double* __restrict__ part, int nblk, int stride, double* __restrict__ out) {
  const int j = blockIdx.x;
  __shared__ double sh[256];
  double s = 0;
  for (int b = threadIdx.x; b < nblk; b += 256) s += part[(size_t)b * stride + j];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[j] = sh[0];
}

// scatter residuals to the reference's measurement order for emba_get_evaluation
__global__ void k_scatter_ref(const uint32_t* __restrict__ refpos, const int32_t* __restrict__ pix,
                              const double* __restrict__ e, int64_t Mc, double* __restrict__ tmp_e,
                              int32_t* __restrict__ tmp_flag) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mc) return;
  if (pix[m] >= 0) {
    const uint32_t r = refpos[m];
    tmp_e[r] = e[m];
    tmp_flag[r] = 1;
  }
}
__global__ void k_compact_ref(const double* __restrict__ tmp_e, const int32_t* __restrict__ flag,
                              const int32_t* __restrict__ pos, int64_t Mt, double* __restrict__ out,
                              int64_t* __restrict__ total) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Mt) return;
  if (flag[r]) out[pos[r]] = tmp_e[r];
  if (r == Mt - 1) *total = (int64_t)pos[r] + flag[r];
}
// the same for the 128-byte Jacobian rows (emba_get_jacobian_rows): 8 threads per row, 16 bytes each; the meta
// slot (last double) is cleared
__global__ void k_scatter_rows(const uint32_t* __restrict__ refpos, const int32_t* __restrict__ pix,
                               const double* __restrict__ jrec, int64_t Mc, double* __restrict__ tmp,
                               int32_t* __restrict__ tmp_flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = i >> 3;
  const int c = (int)(i & 7);
  if (m >= Mc || pix[m] < 0) return;
  const uint32_t r = refpos[m];
  double2 v = reinterpret_cast<const double2*>(jrec)[m * 8 + c];
  if (c == 7) v.y = 0.0;
  reinterpret_cast<double2*>(tmp)[(int64_t)r * 8 + c] = v;
  if (c == 0) tmp_flag[r] = 1;
}
__global__ void k_compact_rows(const double* __restrict__ tmp, const int32_t* __restrict__ flag,
                               const int32_t* __restrict__ pos, int64_t Mt, double* __restrict__ out,
                               int64_t* __restrict__ total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = i >> 3;
  const int c = (int)(i & 7);
  if (r >= Mt) return;
  if (flag[r]) reinterpret_cast<double2*>(out)[(int64_t)pos[r] * 8 + c] = reinterpret_cast<const double2*>(tmp)[r * 8 + c];
  if (r == Mt - 1 && c == 0) *total = (int64_t)pos[r] + flag[r];
}

PanoCam make_cam(const Handle* h) {
  PanoCam c;
  // focalFromFOV(imageSize, 360, 180) (include/utils/equirectangular_camera.h:64-67)
  c.fx = ((double)h->Wp / 360.0) * 180.0 / 3.1415926535897932384626433832795;
  c.fy = ((double)h->Hp / 180.0) * 180.0 / 3.1415926535897932384626433832795;
  c.cx = (double)h->Wp / 2.0;
  c.cy = (double)h->Hp / 2.0;
  return c;
}

int prepare_state_tables(Handle* h, StateSlot& s) {
  dim3 blk(32, 8), grd((h->Wp + 31) / 32, (h->Hp + 7) / 8);
  k_map_prepare<<<grd, blk, 0, h->stream>>>(s.Gx, s.Gy, h->Wp, h->Hp, s.G2, s.H3);
  EMBA_LAUNCH_CHECK();
  k_knot_table<<<ceil_div64(h->n, 64), 64, 0, h->stream>>>(s.quat, h->n, s.Ktab);
  EMBA_LAUNCH_CHECK();
  if (h->B) {
    k_batch_table<<<ceil_div64(h->B, 128), 128, 0, h->stream>>>(s.Ktab, h->d_bs, h->d_bu, h->B, s.RotTab, s.JacTab);
    EMBA_LAUNCH_CHECK();
  }
  return EMBA_OK;
}

int evaluate_slot(Handle* h, int slot, int cost_type, double eta, double alpha) {
  StateSlot& s = h->st[slot];
  // CTAs per SM of the grid-stride kernel (4 are resident). Measured on C4 (ms): 4: 3.27, 8: 3.20, 12: 3.21, 16: 3.27,
  // 24: 3.05, 32: 2.95 - 3.0, 48: 2.99, 64: 3.06, 128: 3.08, 256: 2.94, 1024: 3.10; C2 is flat (0.26) up to 48.
  const int grid_mult = getenv("EMBA_EVAL_GRID") ? std::max(1, atoi(getenv("EMBA_EVAL_GRID"))) : 32;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((h->Mc + kEvalThreads - 1) / kEvalThreads, (int64_t)h->sm_count * grid_mult));
  const int rgrid = h->sm_count * 4;
  EMBA_TRY(dev_reserve(h, &h->d_part, &h->part_cap, (int64_t)2 * grid + rgrid + 16));
  EMBA_CUDA(cudaEventRecord(h->ev[0], h->stream));
  EMBA_TRY(prepare_state_tables(h, s));
  EMBA_CUDA(cudaMemsetAsync(s.hist_loc, 0, sizeof(int32_t) * h->P, h->stream));
  EMBA_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(int32_t) * 16, h->stream));
  const PanoCam cam = make_cam(h);
  EMBA_CUDA(cudaEventRecord(h->ev[1], h->stream));
  if (h->Mc > 0) {
    // resident CTAs per SM the kernel is compiled for: 4 (64 registers, no spills; the default), 5 (48 registers),
    // 6 (40). Measured on C4: 3.58 / 3.77 / 4.32 ms -- the spilled values cost more than the extra warps hide.
    static const int occ = getenv("EMBA_EVAL_OCC") ? atoi(getenv("EMBA_EVAL_OCC")) : 4;
    // default 4; measured on C4 / C2 (ms): PF 0: 3.58 / 0.276, 4: 3.21 / 0.262, 5: 3.21 / 0.285, 6: 3.40 / 0.285, 1: 3.62 / 0.299
    const int pf = getenv("EMBA_EVAL_PF") ? atoi(getenv("EMBA_EVAL_PF")) : 4;  // read per call: tools/eval_variants.py switches it
#define EMBA_EVAL_LAUNCH2(C, B, F)                                                                                   \
  k_eval<C, B, F><<<grid, kEvalThreads, 0, h->stream>>>(h->d_rec, h->Mc, s.Ktab, s.RotTab, s.G2, cam, h->Wp, h->Hp,  \
                                                        h->C_th, eta, s.dp, s.e, s.pix, s.slot, s.hist_loc,           \
                                                        h->d_part, h->d_flags)
#define EMBA_EVAL_LAUNCH(C)                           \
  do {                                                \
    if (occ >= 6) EMBA_EVAL_LAUNCH2(C, 6, 0);         \
    else if (occ == 5) EMBA_EVAL_LAUNCH2(C, 5, 0);    \
    else if (pf == 1) EMBA_EVAL_LAUNCH2(C, 4, 1);     \
    else if (pf == 0) EMBA_EVAL_LAUNCH2(C, 4, 0);     \
    else if (pf == 5) EMBA_EVAL_LAUNCH2(C, 4, 5);     \
    else if (pf == 6) EMBA_EVAL_LAUNCH2(C, 4, 6);     \
    else EMBA_EVAL_LAUNCH2(C, 4, 4);                  \
  } while (0)
    if (cost_type == EMBA_COST_QUADRATIC) EMBA_EVAL_LAUNCH(EMBA_COST_QUADRATIC);
    else if (cost_type == EMBA_COST_CAUCHY) EMBA_EVAL_LAUNCH(EMBA_COST_CAUCHY);
    else EMBA_EVAL_LAUNCH(EMBA_COST_HUBER);
#undef EMBA_EVAL_LAUNCH2
#undef EMBA_EVAL_LAUNCH
    EMBA_LAUNCH_CHECK();
  }
  EMBA_CUDA(cudaEventRecord(h->ev[2], h->stream));
  if (h->Mc > 0) {
    k_sum_partials<<<2, 256, 0, h->stream>>>(h->d_part, grid, 2, h->d_scal);
    EMBA_LAUNCH_CHECK();
  } else {
    EMBA_CUDA(cudaMemsetAsync(h->d_scal, 0, sizeof(double) * 2, h->stream));
  }
  k_reg_partial<<<rgrid, 256, 0, h->stream>>>(s.Gx, s.Gy, h->P, h->d_part + 2 * grid);
  EMBA_LAUNCH_CHECK();
  k_sum_partials<<<1, 256, 0, h->stream>>>(h->d_part + 2 * grid, rgrid, 1, h->d_scal + 2);
  EMBA_LAUNCH_CHECK();
  if (h->world > 1) {
    // the active-pixel decision is global: combine the histogram and the scalars
    EMBA_CUDA(cudaEventRecord(h->ev_x[0], h->stream));
    EMBA_TRY(comm_allreduce_eval(h, s.hist_loc, s.hist, h->P, h->d_scal, h->d_flags));
    EMBA_CUDA(cudaEventRecord(h->ev_x[1], h->stream));
  }
  EMBA_CUDA(cudaEventRecord(h->ev[3], h->stream));
  double sc[3];
  int32_t fl = 0;
  EMBA_CUDA(cudaMemcpyAsync(sc, h->d_scal, sizeof(double) * 3, cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaMemcpyAsync(&fl, h->d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[3]); h->t_ms[0] = ms;
  cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->t_ms[1] = ms;
  if (h->world > 1) { cudaEventElapsedTime(&ms, h->ev_x[0], h->ev_x[1]); h->t_comm_ms[0] = ms; }
  s.cost_data = sc[0];
  s.M = (int64_t)llround(sc[1]);
  s.cost_reg = 0.5 * alpha * sc[2];
  s.evaluated = true;
  if ((fl & 4) && h->strict_range) {  // with several ranks the flag was max-reduced: every rank takes this branch
    h->err = "a warped event rounds past the last panorama element (out-of-bounds access in the reference, model.cpp:213)";
    return EMBA_E_RANGE;
  }
  return EMBA_OK;
}

}  // namespace emba

using namespace emba;

extern "C" {

int emba_set_state(emba_handle_t hh, int32_t which, int64_t t0_ns, int64_t dt_ns, int32_t n_poses,
                   const double* quat, const double* Gx, const double* Gy) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if ((which != 0 && which != 1) || !quat || !Gx || !Gy || n_poses < 2 || dt_ns <= 0) {
    h->err = "emba_set_state: bad argument";
    return EMBA_E_ARG;
  }
  if (n_poses > 65535) {  // control-pose indices travel as 16-bit fields of the Jacobian rows and as 32-bit pair keys
    h->err = "emba_set_state: more than 65535 control poses per window";
    return EMBA_E_ARG;
  }
  EMBA_CUDA(cudaSetDevice(h->device));
  if (t0_ns != h->t0_ns || dt_ns != h->dt_ns || n_poses != h->n) {
    h->t0_ns = t0_ns; h->dt_ns = dt_ns; h->n = n_poses;
    int rc = rebuild_static(h);
    if (rc != EMBA_OK) { h->t0_ns = -1; return rc; }
  }
  StateSlot& s = h->st[which ? 1 - h->cur : h->cur];
  EMBA_CUDA(upload_bytes(h->up, h->stream, s.quat, quat, sizeof(double) * 4 * n_poses));
  EMBA_CUDA(upload_bytes(h->up, h->stream, s.Gx, Gx, sizeof(double) * h->P));
  EMBA_CUDA(upload_bytes(h->up, h->stream, s.Gy, Gy, sizeof(double) * h->P));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  s.evaluated = false;
  if (which == EMBA_STATE_CURRENT) h->formed = h->solved = false;
  return EMBA_OK;
}

int emba_get_state(emba_handle_t hh, int32_t which, double* quat, double* Gx, double* Gy) {
  Handle* h = (Handle*)hh;
  if (!h || (which != 0 && which != 1) || h->n <= 0) return EMBA_E_ARG;
  EMBA_CUDA(cudaSetDevice(h->device));
  StateSlot& s = h->st[which ? 1 - h->cur : h->cur];
  if (quat) EMBA_CUDA(download_bytes(h->up, h->stream, quat, s.quat, sizeof(double) * 4 * h->n));
  if (Gx) EMBA_CUDA(download_bytes(h->up, h->stream, Gx, s.Gx, sizeof(double) * h->P));
  if (Gy) EMBA_CUDA(download_bytes(h->up, h->stream, Gy, s.Gy, sizeof(double) * h->P));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  return EMBA_OK;
}

int emba_evaluate(emba_handle_t hh, int32_t which, int32_t cost_type, double eta, double alpha, double* cost_data,
                  double* cost_reg, int64_t* M) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if ((which != 0 && which != 1) || h->t0_ns < 0 || cost_type < 0 || cost_type > 2) {
    h->err = "emba_evaluate: set events and state first";
    return EMBA_E_ARG;
  }
  EMBA_CUDA(cudaSetDevice(h->device));
  const int slot = which ? 1 - h->cur : h->cur;
  int rc = evaluate_slot(h, slot, cost_type, eta, alpha);
  if (which == EMBA_STATE_CURRENT) h->formed = h->solved = false;
  if (cost_data) *cost_data = h->st[slot].cost_data;
  if (cost_reg) *cost_reg = h->st[slot].cost_reg;
  if (M) *M = h->st[slot].M;
  return rc;
}

int emba_get_evaluation(emba_handle_t hh, int32_t which, double* ep_out, int32_t* num_out) {
  Handle* h = (Handle*)hh;
  if (!h || (which != 0 && which != 1)) return EMBA_E_ARG;
  EMBA_CUDA(cudaSetDevice(h->device));
  StateSlot& s = h->st[which ? 1 - h->cur : h->cur];
  if (!s.evaluated) { h->err = "emba_get_evaluation: state not evaluated"; return EMBA_E_ARG; }
  if (num_out) EMBA_CUDA(download_bytes(h->up, h->stream, num_out, s.hist, sizeof(int32_t) * h->P));
  if (ep_out && h->world > 1) {
    h->err = "emba_get_evaluation: the residual vector in the reference's order is assembled on one GPU only";
    return EMBA_E_SUPPORT;
  }
  if (ep_out && h->Mc_total > 0) {
    // residuals in the reference's order (sensor pixel row-major, then time): scatter by the static reference rank,
    // compact the inliers. No allocation: the scratch arena of the pre-pass is idle between windows.
    const int64_t Mt = h->Mc_total;
    h->ar_tmp.reset();
    double* tmp_e = h->ar_tmp.take<double>(Mt);
    double* d_out = h->ar_tmp.take<double>(Mt);
    int32_t* flag = h->ar_tmp.take<int32_t>(Mt);
    int32_t* pos = h->ar_tmp.take<int32_t>(Mt);
    void* scr = h->ar_tmp.take<char>((int64_t)scan_scratch_bytes(Mt));
    int64_t* d_total = h->ar_tmp.take<int64_t>(2);
    if (!tmp_e || !d_out || !flag || !pos || !scr || !d_total) { h->err = "emba_get_evaluation: scratch arena too small"; return EMBA_E_CUDA; }
    EMBA_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t) * Mt, h->stream));
    if (h->Mc) {
      k_scatter_ref<<<ceil_div64(h->Mc, 256), 256, 0, h->stream>>>(h->d_refpos, s.pix, s.e, h->Mc, tmp_e, flag);
      EMBA_LAUNCH_CHECK();
    }
    EMBA_TRY(scan_exclusive<int32_t>(h, h->stream, flag, pos, Mt, scr));
    k_compact_ref<<<ceil_div64(Mt, 256), 256, 0, h->stream>>>(tmp_e, flag, pos, Mt, d_out, d_total);
    EMBA_LAUNCH_CHECK();
    EMBA_CUDA(cudaMemcpyAsync(h->h_pin + 12, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    EMBA_CUDA(cudaStreamSynchronize(h->stream));
    const int64_t cnt = h->h_pin[12];
    if (cnt > 0) EMBA_CUDA(download_bytes(h->up, h->stream, ep_out, d_out, sizeof(double) * (size_t)cnt));
  }
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  return EMBA_OK;
}

int emba_get_jacobian_rows(emba_handle_t hh, double* rows_out, int64_t cap_rows, int64_t* n_rows) {
  Handle* h = (Handle*)hh;
  if (!h || !rows_out || !n_rows) return EMBA_E_ARG;
  if (!h->formed || !h->jrec_valid) { h->err = "emba_get_jacobian_rows: form the normal equations first"; return EMBA_E_ARG; }
  if (h->world > 1 || h->Mc_total > ((int64_t)1 << 24)) { h->err = "emba_get_jacobian_rows: single GPU, at most 2^24 pairs"; return EMBA_E_SUPPORT; }
  EMBA_CUDA(cudaSetDevice(h->device));
  StateSlot& s = h->st[h->cur];
  const int64_t Mt = h->Mc_total;
  *n_rows = 0;
  if (Mt == 0) return EMBA_OK;
  double *tmp = nullptr, *d_out = nullptr;
  int32_t *flag = nullptr, *pos = nullptr;
  void* scr = nullptr;
  int64_t* d_total = nullptr;
  int rc = EMBA_OK;
  do {  // parity-only download: plain allocations, freed below
    if (cudaMalloc((void**)&tmp, 128 * (size_t)Mt) != cudaSuccess || cudaMalloc((void**)&d_out, 128 * (size_t)Mt) != cudaSuccess ||
        cudaMalloc((void**)&flag, 4 * (size_t)Mt) != cudaSuccess || cudaMalloc((void**)&pos, 4 * (size_t)Mt) != cudaSuccess ||
        cudaMalloc(&scr, scan_scratch_bytes(Mt)) != cudaSuccess || cudaMalloc((void**)&d_total, 16) != cudaSuccess) {
      cudaGetLastError(); h->err = "emba_get_jacobian_rows: out of device memory"; rc = EMBA_E_CUDA; break;
    }
    cudaMemsetAsync(flag, 0, 4 * (size_t)Mt, h->stream);
    k_scatter_rows<<<ceil_div64(h->Mc * 8, 256), 256, 0, h->stream>>>(h->d_refpos, s.pix, h->d_jrec, h->Mc, tmp, flag);
    h->launches++;
    if ((rc = scan_exclusive<int32_t>(h, h->stream, flag, pos, Mt, scr))) break;
    k_compact_rows<<<ceil_div64(Mt * 8, 256), 256, 0, h->stream>>>(tmp, flag, pos, Mt, d_out, d_total);
    h->launches++;
    cudaMemcpyAsync(h->h_pin + 12, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { h->err = std::string("emba_get_jacobian_rows: ") + cudaGetErrorString(e); rc = EMBA_E_CUDA; break; }
    const int64_t cnt = h->h_pin[12];
    *n_rows = cnt;
    if (cnt > cap_rows) { h->err = "emba_get_jacobian_rows: output capacity too small"; rc = EMBA_E_ARG; break; }
    if (cnt > 0 && download_bytes(h->up, h->stream, rows_out, d_out, 128 * (size_t)cnt) != cudaSuccess) {
      cudaGetLastError(); h->err = "emba_get_jacobian_rows: copy failed"; rc = EMBA_E_CUDA; break;
    }
  } while (0);
  cudaFree(tmp); cudaFree(d_out); cudaFree(flag); cudaFree(pos); cudaFree(scr); cudaFree(d_total);
  return rc;
}

int emba_last_timings_ms(emba_handle_t hh, double* out8) {
  Handle* h = (Handle*)hh;
  if (!h || !out8) return EMBA_E_ARG;
  for (int i = 0; i < 8; i++) out8[i] = h->t_ms[i];
  return EMBA_OK;
}
int emba_launch_count(emba_handle_t hh, int64_t* out) {
  Handle* h = (Handle*)hh;
  if (!h || !out) return EMBA_E_ARG;
  *out = h->launches;
  return EMBA_OK;
}
int emba_synchronize(emba_handle_t hh) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  EMBA_CUDA(cudaSetDevice(h->device));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  return EMBA_OK;
}

}  // extern "C"
