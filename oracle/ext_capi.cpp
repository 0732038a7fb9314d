// TEST INFRASTRUCTURE ONLY (oracle/). Never linked into, imported by or called from the product (emba_b200/).
//
// CPU restatement (fp64) of the EXTENSION MODE of the hot path: what the north star text
// describes beyond the reference's parity mode --
//   * cubic cumulative SO(3) B-spline, pose evaluated PER EVENT (no batches of 100), value and Jacobian w.r.t. the 4
//     active control poses taken straight from the reference's vendored basalt::So3Spline<4>::evaluate
//     (thirdparty/basalt-headers/include/basalt/spline/so3_spline.h:218-274), exactly as the reference's
//     CubicTrajectory::evaluate wraps it (src/utils/trajectory.cpp:453-479: 3 x 12 Jacobian);
//   * projection and its Jacobian from the reference's own EquirectangularCamera::projectToImage
//     (include/utils/equirectangular_camera.h:18-45) and drb/ddrot = -[rb]x (src/utils/event_pano_warper.cpp:62);
//   * BILINEAR sampling of the gradient map at the warped current event (the parity mode rounds to the nearest
//     pixel, model.cpp:209-214): 4 neighbours -> 8 map Jacobian entries; the map's spatial derivative is the exact
//     derivative of the bilinear interpolant (the parity mode uses Sobel Hessians, model.cpp:87-97).
// The event-generation model, the pairing (consecutive events of one sensor pixel), the outlier gate |dp| > 10
// (model.cpp:199-205) and the sign conventions (J = d C_pred / d unknown, model.cpp:418-487) are the reference's.
// There is no reference implementation of this mode to compare with ("parity unpinned" for N4): this file is the
// oracle of the CUDA extension kernels and is itself checked against central finite differences (tests/test_ext.py).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include <Eigen/Dense>
#include <basalt/spline/so3_spline.h>
#include "utils/equirectangular_camera.h"

namespace {
typedef basalt::So3Spline<4, double> Spline4;

Eigen::Matrix3d skew(const Eigen::Vector3d& v) {
  Eigen::Matrix3d m;
  m << 0, -v.z(), v.y(), v.z(), 0, -v.x(), -v.y(), v.x(), 0;
  return m;
}

struct Warp {
  Eigen::Vector2d pm;
  Eigen::Matrix<double, 2, 12> E;  // d pm / d (4 control-pose left perturbations)
  int s;
};

Warp warp(const Spline4& spl, const dvs::EquirectangularCamera& cam, int64_t t_ns, const Eigen::Vector3d& b, bool jac) {
  Warp w;
  Spline4::JacobianStruct J;
  const Sophus::SO3d R = spl.evaluate(t_ns, jac ? &J : nullptr);
  const Eigen::Vector3d rb = R * b;
  cv::Matx23d jp;
  w.pm = cam.projectToImage(rb, jac ? &jp : nullptr);
  w.s = jac ? (int)J.start_idx : 0;
  if (jac) {
    Eigen::Matrix<double, 2, 3> Jp;
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 3; j++) Jp(i, j) = jp(i, j);
    const Eigen::Matrix<double, 2, 3> M = Jp * (-skew(rb));  // event_pano_warper.cpp:62-65
    for (int k = 0; k < 4; k++) w.E.block<2, 3>(0, 3 * k) = M * J.d_val_d_knot[k];
  }
  return w;
}
}  // namespace

extern "C" {

// Per-measurement rows of the extension mode, measurements in time order of their CURRENT event.
// Inputs: time-sorted events, bearing LUT [S*3], spline (t0_ns, dt_ns, n knots xyzw), maps row-major [H*W].
// Outputs (capacity N rows each): e, dp[2], pm_c[2], Jc[12], Jp[12], w[4], cp[2] = (s_c, s_p), pix[4] (linear map
// indices of the bilinear footprint: (x0,y0), (x0+1,y0), (x0,y0+1), (x0+1,y0+1); x wraps, y clamps),
// ev[1] = index of the current event. Outliers (|dp| > 10) are skipped. Returns the number of rows, or -1 when a
// timestamp leaves the spline support (basalt asserts there).
long embaext_rows(int sensor_w, int sensor_h, const double* lut, int pano_w, int pano_h, double C_th, long N,
                  const uint16_t* x, const uint16_t* y, const int64_t* t_ns, const uint8_t* pol, int n_knots,
                  int64_t t0_ns, int64_t dt_ns, const double* quat_xyzw, const double* Gx, const double* Gy,
                  double* e_out, double* dp_out, double* pm_out, double* Jc_out, double* Jp_out, double* w_out,
                  int32_t* cp_out, int32_t* pix_out, int32_t* ev_out) {
  Spline4 spl(dt_ns, t0_ns);
  for (int i = 0; i < n_knots; i++) {
    Eigen::Quaterniond q(quat_xyzw[4 * i + 3], quat_xyzw[4 * i], quat_xyzw[4 * i + 1], quat_xyzw[4 * i + 2]);
    spl.knotsPushBack(Sophus::SO3d(q));
  }
  const dvs::EquirectangularCamera cam(cv::Size(pano_w, pano_h), 360.0, 180.0);
  const int64_t t_max = t0_ns + (int64_t)(n_knots - 3) * dt_ns;
  std::vector<long> last((size_t)sensor_w * sensor_h, -1);
  long m = 0;
  for (long i = 0; i < N; i++) {
    const size_t sp = (size_t)y[i] * sensor_w + x[i];
    const long p = last[sp];
    last[sp] = i;
    if (p < 0) continue;
    if (t_ns[i] < t0_ns || t_ns[i] >= t_max || t_ns[p] < t0_ns) return -1;
    const Eigen::Vector3d b(lut[3 * sp], lut[3 * sp + 1], lut[3 * sp + 2]);
    const Warp wc = warp(spl, cam, t_ns[i], b, true), wp = warp(spl, cam, t_ns[p], b, true);
    const Eigen::Vector2d dp = wc.pm - wp.pm;
    if (dp.norm() > 10) continue;  // model.cpp:199-205
    const double fx0 = std::floor(wc.pm.x()), fy0 = std::floor(wc.pm.y());
    const double ax = wc.pm.x() - fx0, ay = wc.pm.y() - fy0;
    const int x0 = (((int)fx0 % pano_w) + pano_w) % pano_w, x1 = (x0 + 1) % pano_w;
    const int y0 = std::min(std::max((int)fy0, 0), pano_h - 1), y1 = std::min(std::max((int)fy0 + 1, 0), pano_h - 1);
    const int idx[4] = {y0 * pano_w + x0, y0 * pano_w + x1, y1 * pano_w + x0, y1 * pano_w + x1};
    const double w[4] = {(1 - ax) * (1 - ay), ax * (1 - ay), (1 - ax) * ay, ax * ay};
    Eigen::Vector2d G(0, 0), g4[4];
    for (int k = 0; k < 4; k++) { g4[k] = Eigen::Vector2d(Gx[idx[k]], Gy[idx[k]]); G += w[k] * g4[k]; }
    // derivative of the bilinear interpolant w.r.t. the sampling position
    const Eigen::Vector2d dG_dx = (1 - ay) * (g4[1] - g4[0]) + ay * (g4[3] - g4[2]);
    const Eigen::Vector2d dG_dy = (1 - ax) * (g4[2] - g4[0]) + ax * (g4[3] - g4[1]);
    const double C_pred = G.dot(dp);
    const double C_meas = 2 * ((double)pol[i] - 0.5) * C_th;  // model.cpp:219
    e_out[m] = C_meas - C_pred;
    dp_out[2 * m] = dp.x(); dp_out[2 * m + 1] = dp.y();
    pm_out[2 * m] = wc.pm.x(); pm_out[2 * m + 1] = wc.pm.y();
    // temp = Gpm + dp^T dG/dpm (the parity mode's model.cpp:233-238 with the interpolant's own derivative)
    const Eigen::RowVector2d h(G.x() + dp.dot(dG_dx), G.y() + dp.dot(dG_dy));
    const Eigen::Matrix<double, 1, 12> Jc = h * wc.E;                 // model.cpp:449
    const Eigen::Matrix<double, 1, 12> Jp = -G.transpose() * wp.E;    // model.cpp:459
    for (int k = 0; k < 12; k++) { Jc_out[12 * m + k] = Jc(k); Jp_out[12 * m + k] = Jp(k); }
    for (int k = 0; k < 4; k++) { w_out[4 * m + k] = w[k]; pix_out[4 * m + k] = idx[k]; }
    cp_out[2 * m] = wc.s; cp_out[2 * m + 1] = wp.s;
    ev_out[m] = (int32_t)i;
    m++;
  }
  return m;
}

// C_pred of one pair as a function of the knots and the maps, for finite-difference checks of the rows above:
// events i (current) and p (previous) of sensor pixel sp.
double embaext_cpred(const double* lut, int sensor_pix, int pano_w, int pano_h, int64_t t_c, int64_t t_p, int n_knots,
                     int64_t t0_ns, int64_t dt_ns, const double* quat_xyzw, const double* Gx, const double* Gy) {
  Spline4 spl(dt_ns, t0_ns);
  for (int i = 0; i < n_knots; i++) {
    Eigen::Quaterniond q(quat_xyzw[4 * i + 3], quat_xyzw[4 * i], quat_xyzw[4 * i + 1], quat_xyzw[4 * i + 2]);
    spl.knotsPushBack(Sophus::SO3d(q));
  }
  const dvs::EquirectangularCamera cam(cv::Size(pano_w, pano_h), 360.0, 180.0);
  const Eigen::Vector3d b(lut[3 * sensor_pix], lut[3 * sensor_pix + 1], lut[3 * sensor_pix + 2]);
  const Warp wc = warp(spl, cam, t_c, b, false), wp = warp(spl, cam, t_p, b, false);
  const Eigen::Vector2d dp = wc.pm - wp.pm;
  const double fx0 = std::floor(wc.pm.x()), fy0 = std::floor(wc.pm.y());
  const double ax = wc.pm.x() - fx0, ay = wc.pm.y() - fy0;
  const int x0 = (((int)fx0 % pano_w) + pano_w) % pano_w, x1 = (x0 + 1) % pano_w;
  const int y0 = std::min(std::max((int)fy0, 0), pano_h - 1), y1 = std::min(std::max((int)fy0 + 1, 0), pano_h - 1);
  const int idx[4] = {y0 * pano_w + x0, y0 * pano_w + x1, y1 * pano_w + x0, y1 * pano_w + x1};
  const double w[4] = {(1 - ax) * (1 - ay), ax * (1 - ay), (1 - ax) * ay, ax * ay};
  Eigen::Vector2d G(0, 0);
  for (int k = 0; k < 4; k++) G += w[k] * Eigen::Vector2d(Gx[idx[k]], Gy[idx[k]]);
  return G.dot(dp);
}

}  // extern "C"
