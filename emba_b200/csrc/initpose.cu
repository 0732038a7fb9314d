// Control-pose initialisation from a dense front-end trajectory (SURVEY section 8(f) "next" row N2):
// LinearTrajectory::generateCtrlPosesLong (reference src/utils/trajectory.cpp:258-294) with generateCtrlPoses
// (:245-256) and fitCtrlPoses (:149-229), as EMBA::Run calls it with sub-interval = dt_knots
// (src/emba/emba.cpp:416): per knot interval, lift the front-end poses inside the interval to the tangent space at
// the first of them, least-squares fit of the two linear-spline control increments, retract.
// One thread per knot interval; the 2-unknown least-squares problem is solved through its 2x2 normal equations
// (the reference uses a full-pivoting Householder QR on the same tiny, well-conditioned system).
#include <cmath>
#include <vector>

#include "emba_internal.cuh"

namespace emba {

__device__ __forceinline__ double ros_to_sec(int64_t ns) {
  const int64_t sec = ns / 1000000000LL;
  const int64_t nsec = ns - sec * 1000000000LL;
  return __dadd_rn((double)sec, __dmul_rn(1e-9, (double)nsec));  // ros::Time::toSec, no FMA contraction
}

__global__ void k_fit_ctrl(int n_int, const int64_t* __restrict__ beg_ns, const int64_t* __restrict__ end_ns,
                           int64_t n_poses, const int64_t* __restrict__ t_ns, const double* __restrict__ quat,
                           double dt_knots, double* __restrict__ out, int32_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_int) return;
  const int64_t b = beg_ns[i], e = end_ns[i];
  // upper_bound(t_subset_beg) / lower_bound(t_subset_end) on the time-sorted poses (trajectory.cpp:273-274)
  int64_t lo = 0, hi = n_poses;
  while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (t_ns[m] <= b) lo = m + 1; else hi = m; }
  const int64_t i0 = lo;
  lo = i0; hi = n_poses;
  while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (t_ns[m] < e) lo = m + 1; else hi = m; }
  const int64_t i1 = lo;
  if (i1 - i0 < 2) { atomicOr(flags, 1); return; }  // CHECK_GE(poses.size(), num_cps), trajectory.cpp:153
  const double4 off = reinterpret_cast<const double4*>(quat)[i0];
  const double4 offinv = make_double4(-off.x, -off.y, -off.z, off.w);
  const double tb = ros_to_sec(b);
  double s00 = 0, s01 = 0, s11 = 0;
  double r0[3] = {0, 0, 0}, r1[3] = {0, 0, 0};
  for (int64_t k = i0; k < i1; k++) {
    const Vec3 D = so3_log(quat_mul(offinv, reinterpret_cast<const double4*>(quat)[k]));  // lift, :160-167
    const double t = ros_to_sec(t_ns[k]);
    const double ti = floor((t - tb) / dt_knots);  // :197
    if (ti != 0.0) atomicOr(flags, 2);             // the reference would index N out of bounds
    const double u = (t - (ti * dt_knots + tb)) / dt_knots;  // :199
    const double a = 1.0 - u;  // [1, u] * M2 = [1 - u, u], :201-205
    s00 += a * a; s01 += a * u; s11 += u * u;
    r0[0] += a * D.x; r0[1] += a * D.y; r0[2] += a * D.z;
    r1[0] += u * D.x; r1[1] += u * D.y; r1[2] += u * D.z;
  }
  const double det = s00 * s11 - s01 * s01;
  if (!(fabs(det) > 0.0)) { atomicOr(flags, 4); return; }
  Vec3 p0, p1;
  p0.x = (s11 * r0[0] - s01 * r1[0]) / det; p0.y = (s11 * r0[1] - s01 * r1[1]) / det; p0.z = (s11 * r0[2] - s01 * r1[2]) / det;
  p1.x = (s00 * r1[0] - s01 * r0[0]) / det; p1.y = (s00 * r1[1] - s01 * r0[1]) / det; p1.z = (s00 * r1[2] - s01 * r0[2]) / det;
  auto retract = [&](const Vec3& p) {  // offset * exp(drotv), :223
    double4 q = quat_mul(off, so3_exp(p));
    const double nrm = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= nrm; q.y /= nrm; q.z /= nrm; q.w /= nrm;
    return q;
  };
  // the first control pose of every sub-interval but the first is dropped (:280-283)
  if (i == 0) reinterpret_cast<double4*>(out)[0] = retract(p0);
  reinterpret_cast<double4*>(out)[i + 1] = retract(p1);
}

// ros::Duration(double): sec = floor(d), nsec = round((d - sec) * 1e9)
static int64_t ros_duration_ns(double d) {
  const double s = std::floor(d);
  volatile double frac = (d - s) * 1e9;
  return (int64_t)s * 1000000000LL + (int64_t)std::round(frac);
}
static double ros_ns_to_sec(int64_t ns) {
  const int64_t sec = ns / 1000000000LL, nsec = ns - sec * 1000000000LL;
  volatile double a = 1e-9 * (double)nsec;
  return (double)sec + a;
}

}  // namespace emba

using namespace emba;

extern "C" int emba_fit_control_poses(int32_t device, int64_t n_poses, const int64_t* t_ns, const double* quat_xyzw,
                                      int64_t tb_ns, int64_t te_ns, double dt_knots, double* ctrl_quat_out,
                                      int32_t cap, int32_t* n_ctrl_out) {
  if (!t_ns || !quat_xyzw || !ctrl_quat_out || !n_ctrl_out || n_poses < 2 || !(dt_knots > 0) || !(te_ns > tb_ns))
    return EMBA_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return EMBA_E_CUDA;
  // host: sub-interval boundaries in ros::Time / ros::Duration arithmetic (trajectory.cpp:263-270)
  // t_beg / t_end arrive as the ns-exact ros::Time::toNSec() of the interval (emba.cpp:416)
  const double span = ros_ns_to_sec(te_ns - tb_ns);
  const int n_int = (int)std::floor(span / dt_knots + 1e-6);
  if (n_int < 1) return EMBA_E_ARG;
  if (n_int + 1 > cap) return EMBA_E_ARG;
  const int64_t sub_ns = ros_duration_ns(dt_knots);
  const double sub_sec = ros_ns_to_sec(sub_ns);
  std::vector<int64_t> beg(n_int), end(n_int);
  for (int i = 0; i < n_int; i++) {
    volatile double d = sub_sec * (double)i;  // Duration * i -> Duration(toSec() * i)
    beg[i] = tb_ns + ros_duration_ns(d);
    end[i] = beg[i] + sub_ns;
  }
  int64_t *d_beg = nullptr, *d_end = nullptr, *d_t = nullptr;
  double *d_q = nullptr, *d_out = nullptr;
  int32_t* d_flags = nullptr;
  int rc = EMBA_OK;
  int32_t fl = 0;
  do {
    if (cudaMalloc(&d_beg, sizeof(int64_t) * n_int) != cudaSuccess || cudaMalloc(&d_end, sizeof(int64_t) * n_int) != cudaSuccess ||
        cudaMalloc(&d_t, sizeof(int64_t) * n_poses) != cudaSuccess || cudaMalloc(&d_q, sizeof(double) * 4 * n_poses) != cudaSuccess ||
        cudaMalloc(&d_out, sizeof(double) * 4 * (n_int + 1)) != cudaSuccess || cudaMalloc(&d_flags, sizeof(int32_t)) != cudaSuccess) {
      rc = EMBA_E_CUDA; break;
    }
    cudaMemcpy(d_beg, beg.data(), sizeof(int64_t) * n_int, cudaMemcpyHostToDevice);
    cudaMemcpy(d_end, end.data(), sizeof(int64_t) * n_int, cudaMemcpyHostToDevice);
    cudaMemcpy(d_t, t_ns, sizeof(int64_t) * n_poses, cudaMemcpyHostToDevice);
    cudaMemcpy(d_q, quat_xyzw, sizeof(double) * 4 * n_poses, cudaMemcpyHostToDevice);
    cudaMemset(d_flags, 0, sizeof(int32_t));
    k_fit_ctrl<<<(n_int + 127) / 128, 128>>>(n_int, d_beg, d_end, n_poses, d_t, d_q, dt_knots, d_out, d_flags);
    if (cudaMemcpy(&fl, d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(ctrl_quat_out, d_out, sizeof(double) * 4 * (n_int + 1), cudaMemcpyDeviceToHost) != cudaSuccess) {
      rc = EMBA_E_CUDA; break;
    }
    if (fl) rc = EMBA_E_SUPPORT;  // too few front-end poses in a knot interval (the reference aborts: CHECK_GE)
    *n_ctrl_out = n_int + 1;
  } while (0);
  cudaFree(d_beg); cudaFree(d_end); cudaFree(d_t); cudaFree(d_q); cudaFree(d_out); cudaFree(d_flags);
  return rc;
}
