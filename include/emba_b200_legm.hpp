// Reference-side binding of the emba_b200 C ABI: a header-only C++ class with the method names, argument lists and
// call semantics of the reference's `EMBA::LEGM` (reference include/emba/model.h:72-133), so that
// `EMBA::solveTimeWindow` (src/emba/solver.cpp:11-368), `EMBA::EMBA` (src/emba/emba.cpp), the launch files, the
// calibration YAMLs and image_rec keep working unchanged. It is compiled INSIDE the reference's catkin package
// (it includes the reference's own headers) and links against libemba_b200.so; see INTEGRATION.md.
//
// This file is not part of the library build; tests/test_adapter_compiles.py syntax-checks it against the
// reference headers where /root/reference is available.
#pragma once

#include <cmath>
#include <cstdint>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "emba/model.h"  // the reference's Model base class, typedefs, Trajectory, EventWarper
#include "emba_b200.h"

namespace EMBA {

class LEGM_B200 : public Model {
public:
  // LEGM::LEGM (model.cpp:56-70). The bearing LUT is computed on the host exactly as
  // EventWarper::precomputeBearingVectors does (event_pano_warper.cpp:27-41), so image_geometry (and any lens
  // distortion it handles) stays on the CPU side.
  LEGM_B200(const sensor_msgs::CameraInfo& camera_info_msg, double C_th, int pano_width, int pano_height,
            int device = 0) {
    C_th_ = C_th;
    event_warper_ptr_ = new dvs::EventWarper();
    event_warper_ptr_->initialize(camera_info_msg, pano_width, pano_height);
    image_geometry::PinholeCameraModel cam;
    cam.fromCameraInfo(camera_info_msg);
    const int W = camera_info_msg.width, H = camera_info_msg.height;
    std::vector<double> lut(3 * (size_t)W * H);
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        const cv::Point2d r = cam.rectifyPoint(cv::Point2d(x, y));
        const cv::Point3d b = cam.projectPixelTo3dRay(r);
        double* o = &lut[3 * ((size_t)y * W + x)];
        o[0] = b.x; o[1] = b.y; o[2] = b.z;
      }
    emba_config_t cfg;
    cfg.sensor_w = W; cfg.sensor_h = H; cfg.pano_w = pano_width; cfg.pano_h = pano_height;
    cfg.C_th = C_th; cfg.bearing_lut = lut.data(); cfg.device = device;
    check(emba_create(&cfg, &h_), "emba_create");
    pano_w_ = pano_width; pano_h_ = pano_height;
  }
  ~LEGM_B200() {
    emba_destroy(h_);
    delete event_warper_ptr_;
  }
  LEGM_B200(const LEGM_B200&) = delete;
  LEGM_B200& operator=(const LEGM_B200&) = delete;

  // LEGM::evaluateDataError (model.cpp:72-258). Like the reference, the model remembers the per-measurement state
  // of the LAST evaluated point; formNormalEq uses it.
  VecXd evaluateDataError(Trajectory* traj_ptr, const cv::Mat& Gx, const cv::Mat& Gy, const EventPacket& events,
                          bool /*eval_deriv*/, cv::Mat& num_ev_map) {
    syncEvents(events);
    const int n = (int)traj_ptr->size();
    std::vector<double> q(4 * (size_t)n);
    for (int i = 0; i < n; i++) {
      const Eigen::Quaterniond u = traj_ptr->getControlPose(i).unit_quaternion();
      q[4 * i] = u.x(); q[4 * i + 1] = u.y(); q[4 * i + 2] = u.z(); q[4 * i + 3] = u.w();
    }
    // the integer time base the spline itself uses (trajectory.cpp:61-70): protected members, read through a
    // pointer to member named via a derived class (legal access to a protected base member)
    const int64_t t0 = traj_ptr->*(&Peek::t_beg_ns_);
    const int64_t dt = traj_ptr->*(&Peek::dt_knots_ns_);
    check(emba_set_state(h_, EMBA_STATE_CANDIDATE, t0, dt, n, q.data(), Gx.ptr<double>(), Gy.ptr<double>()),
          "emba_set_state");
    double cd = 0, cr = 0;
    int64_t M = 0;
    check(emba_evaluate(h_, EMBA_STATE_CANDIDATE, cost_type_, eta_, 0.0, &cd, &cr, &M), "emba_evaluate");
    pending_ = true;
    n_poses_ = n;
    last_data_cost_ = cd;
    VecXd ep(M);
    if (num_ev_map.empty()) num_ev_map = cv::Mat::zeros(pano_h_, pano_w_, CV_32SC1);
    check(emba_get_evaluation(h_, EMBA_STATE_CANDIDATE, ep.data(), num_ev_map.ptr<int32_t>()), "emba_get_evaluation");
    return ep;
  }

  // model.cpp:260-277
  VecXd evaluateRegError(const cv::Mat& Gx, const cv::Mat& Gy) {
    const size_t P = (size_t)Gx.rows * Gx.cols;
    VecXd ep(2 * P);
    const double* gx = Gx.ptr<double>();
    const double* gy = Gy.ptr<double>();
    for (size_t i = 0; i < P; i++) { ep(2 * i) = gx[i]; ep(2 * i + 1) = gy[i]; }
    return ep;
  }

  // model.cpp:279-314 (already reduced on the device by the last evaluateDataError when the same cost is configured)
  double evaluateRobustDataCost(const VecXd& ep, const std::string cost_type, const double a) {
    setRobustCost(cost_type, a);
    double c = 0;
    if (cost_type == "cauchy") {
      for (Eigen::Index k = 0; k < ep.size(); k++) c += std::log1p(a * ep(k) * ep(k));
      return (0.5 / a) * c;
    }
    for (Eigen::Index k = 0; k < ep.size(); k++) {
      const double e = std::abs(ep(k));
      c += e < a ? 0.5 * e * e : a * e - 0.5 * a * a;
    }
    return c;
  }

  // model.cpp:316-491
  void formNormalEq(MatXd& A11, MatXd& A12, std::vector<Mat2d>& A22_blocks, VecXd& b1, VecXd& b2, const VecXd& /*ep*/,
                    const int num_ctrl_poses, const cv::Mat& /*num_ev_map*/, const int thres_valid_pixel,
                    std::set<size_t>& active_pix_idxes, std::set<size_t>& inactive_pix_idxes) {
    form(A11, A12, A22_blocks, b1, b2, num_ctrl_poses, thres_valid_pixel, active_pix_idxes, inactive_pix_idxes,
         EMBA_COST_QUADRATIC, 1.0);
  }
  // model.cpp:493-687
  void formNormalEqIRLS(MatXd& A11, MatXd& A12, std::vector<Mat2d>& A22_blocks, VecXd& b1, VecXd& b2,
                        const VecXd& /*ep*/, const int num_ctrl_poses, const cv::Mat& /*num_ev_map*/,
                        const int thres_valid_pixel, std::set<size_t>& active_pix_idxes,
                        std::set<size_t>& inactive_pix_idxes, const std::string cost_type, const double a) {
    setRobustCost(cost_type, a);
    form(A11, A12, A22_blocks, b1, b2, num_ctrl_poses, thres_valid_pixel, active_pix_idxes, inactive_pix_idxes,
         cost_type_, eta_);
  }

  // model.cpp:689-719: device-resident A22/b2 and the caller's host copies are updated alike
  void applyL2Reg(std::vector<Mat2d>& A22_blocks, VecXd& b2, const std::set<size_t>& active_pix_idxes,
                  const double alpha, const cv::Mat& Gx, const cv::Mat& Gy) {
    check(emba_apply_l2_reg(h_, alpha), "emba_apply_l2_reg");
    for (auto& B : A22_blocks) { B(0, 0) += alpha; B(1, 1) += alpha; }
    size_t j = 0;
    const double* gx = Gx.ptr<double>();
    const double* gy = Gy.ptr<double>();
    for (auto p : active_pix_idxes) { b2(2 * j) -= alpha * gx[p]; b2(2 * j + 1) -= alpha * gy[p]; j++; }
  }

  // model.cpp:721-792. The matrices stay on the device; the host arguments only tell whether solver.cpp has
  // removed the first control pose (solver.cpp:156-165).
  void solveNormalEq(const MatXd& /*A11*/, const MatXd& /*A12*/, const std::vector<Mat2d>& A22_blocks, const VecXd& b1,
                     const VecXd& /*b2*/, const double lambda, VecXd& x1, VecXd& x2) {
    const int fix = (b1.size() == 3 * (n_poses_ - 1)) ? 1 : 0;
    x1.resize(b1.size());
    x2.resize(2 * (Eigen::Index)A22_blocks.size());
    check(emba_solve(h_, lambda, 0, fix, x1.data(), x2.data(), nullptr, nullptr), "emba_solve");
  }
  // model.cpp:794-840
  std::pair<int, double> solveNormalEqCG(const MatXd& /*A11*/, const MatXd& /*A12*/,
                                         const std::vector<Mat2d>& A22_blocks, const VecXd& b1, const VecXd& /*b2*/,
                                         const double lambda, VecXd& x1, VecXd& x2) {
    const int fix = (b1.size() == 3 * (n_poses_ - 1)) ? 1 : 0;
    x1.resize(b1.size());
    x2.resize(2 * (Eigen::Index)A22_blocks.size());
    int32_t it = 0;
    double err = 0;
    check(emba_solve(h_, lambda, 1, fix, x1.data(), x2.data(), &it, &err), "emba_solve");
    return std::pair<int, double>(it, err);
  }

  // model.cpp:863-903 (O(P) on the caller's cv::Mat clones, like the reference; the next evaluateDataError
  // uploads them)
  void updateMap(cv::Mat& Gx_new, cv::Mat& Gy_new, const VecXd& x2, const double damping_factor,
                 const std::set<size_t>& active_pix_idxes, const std::set<size_t>& inactive_pix_idxes) {
    double* gx = Gx_new.ptr<double>();
    double* gy = Gy_new.ptr<double>();
    size_t i = 0;
    for (auto p : active_pix_idxes) { gx[p] += damping_factor * x2(2 * i); gy[p] += damping_factor * x2(2 * i + 1); i++; }
    for (auto p : inactive_pix_idxes) { gx[p] = 0.0; gy[p] = 0.0; }
  }

  // model.cpp:842-861
  SpMat recoverA22FromBlocks(const std::vector<Mat2d>& A22_blocks) {
    std::vector<Triplet> t;
    t.reserve(4 * A22_blocks.size());
    for (size_t i = 0; i < A22_blocks.size(); i++) {
      t.push_back(Triplet(2 * i, 2 * i, A22_blocks[i](0, 0)));
      t.push_back(Triplet(2 * i + 1, 2 * i + 1, A22_blocks[i](1, 1)));
      t.push_back(Triplet(2 * i + 1, 2 * i, A22_blocks[i](1, 0)));
      t.push_back(Triplet(2 * i, 2 * i + 1, A22_blocks[i](0, 1)));
    }
    SpMat A22(2 * A22_blocks.size(), 2 * A22_blocks.size());
    A22.setFromTriplets(t.begin(), t.end());
    A22.makeCompressed();
    return A22;
  }

  // Whole LM loop on the device (replaces the body of EMBA::solveTimeWindow, solver.cpp:11-368): one upload, one
  // download. Optional; the per-call interface above already keeps solver.cpp unchanged.
  double solveTimeWindowOnDevice(Trajectory*& traj_ptr, cv::Mat& Gx, cv::Mat& Gy, const EventPacket& events,
                                 const emba_lm_settings_t& s, std::vector<emba_lm_log_t>* log = nullptr) {
    cv::Mat num;
    cost_type_ = s.cost_type; eta_ = s.eta;
    evaluateDataError(traj_ptr, Gx, Gy, events, true, num);  // uploads events + state into the candidate slot
    check(emba_accept_candidate(h_), "emba_accept_candidate");
    pending_ = false;
    std::vector<emba_lm_log_t> rows(s.max_num_iter + 8);
    int32_t nlog = 0;
    double fcost = 0;
    check(emba_solve_time_window(h_, &s, rows.data(), (int32_t)rows.size(), &nlog, &fcost), "emba_solve_time_window");
    if (log) log->assign(rows.begin(), rows.begin() + std::min<int32_t>(nlog, (int32_t)rows.size()));
    const int n = (int)traj_ptr->size();
    std::vector<double> q(4 * (size_t)n);
    check(emba_get_state(h_, EMBA_STATE_CURRENT, q.data(), Gx.ptr<double>(), Gy.ptr<double>()), "emba_get_state");
    std::vector<Sophus::SO3d> cps;
    for (int i = 0; i < n; i++) cps.emplace_back(Eigen::Quaterniond(q[4 * i + 3], q[4 * i], q[4 * i + 1], q[4 * i + 2]));
    // replace the control poses in place through the public API (trajectory.h:67-68)
    LinearTrajectory refined(traj_ptr->begTime().toSec(), traj_ptr->getDtCtrlPoses(), cps);
    traj_ptr->replaceWith(&refined, n, 0, 0);
    return fcost;
  }

  double lastDataCost() const { return last_data_cost_; }
  void setDenseA12Download(bool on) { dense_a12_ = on; }

private:
  struct Peek : public Trajectory {
    using Trajectory::dt_knots_ns_;
    using Trajectory::t_beg_ns_;
  };

  void check(int rc, const char* what) {
    if (rc != EMBA_OK) throw std::runtime_error(std::string(what) + ": " + (h_ ? emba_last_error(h_) : "no handle"));
  }
  void setRobustCost(const std::string& cost_type, double a) {
    cost_type_ = cost_type == "cauchy" ? EMBA_COST_CAUCHY : cost_type == "huber" ? EMBA_COST_HUBER : EMBA_COST_QUADRATIC;
    eta_ = a;
  }
  void syncEvents(const EventPacket& events) {
    if (events.data() == ev_ptr_ && events.size() == ev_size_) return;  // same window: already on the device
    const size_t N = events.size();
    std::vector<uint16_t> x(N), y(N);
    std::vector<int64_t> t(N);
    std::vector<uint8_t> p(N);
    for (size_t i = 0; i < N; i++) {
      x[i] = events[i].x; y[i] = events[i].y; p[i] = events[i].polarity;
      t[i] = (int64_t)events[i].ts.toNSec();
    }
    check(emba_set_events(h_, (int64_t)N, x.data(), y.data(), t.data(), p.data()), "emba_set_events");
    ev_ptr_ = events.data();
    ev_size_ = N;
  }
  void form(MatXd& A11, MatXd& A12, std::vector<Mat2d>& A22_blocks, VecXd& b1, VecXd& b2, int n, int thres,
            std::set<size_t>& active, std::set<size_t>& inactive, int cost_type, double eta) {
    if (pending_) {  // the last evaluated point becomes the linearisation point (solver.cpp:96-130)
      check(emba_accept_candidate(h_), "emba_accept_candidate");
      pending_ = false;
    }
    int64_t Np = 0;
    check(emba_form_normal_eq(h_, thres, cost_type, eta, 0.0, &Np), "emba_form_normal_eq");
    const int d = 3 * n;
    Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> A11r(d, d);
    std::vector<double> a22(4 * (size_t)Np);
    std::vector<int64_t> act((size_t)Np);
    b1.resize(d);
    b2.resize(2 * Np);
    Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> A12r;
    if (dense_a12_) A12r.resize(d, 2 * Np);
    check(emba_get_normal_eq(h_, A11r.data(), b1.data(), a22.data(), b2.data(), act.data(),
                             dense_a12_ ? A12r.data() : nullptr), "emba_get_normal_eq");
    A11 = A11r;
    if (dense_a12_) A12 = A12r;
    else A12.resize(d, 0);  // A12 stays on the device (per-pixel strips); solveNormalEq does not read the host copy
    A22_blocks.resize((size_t)Np);
    for (int64_t i = 0; i < Np; i++) A22_blocks[i] << a22[4 * i], a22[4 * i + 1], a22[4 * i + 2], a22[4 * i + 3];
    active.clear();
    inactive.clear();
    size_t k = 0;
    const size_t P = (size_t)pano_w_ * pano_h_;
    for (size_t p = 0; p < P; p++) {
      if (k < (size_t)Np && (size_t)act[k] == p) { active.insert(active.end(), p); k++; }
      else inactive.insert(inactive.end(), p);
    }
  }

  emba_handle_t h_ = nullptr;
  int pano_w_ = 0, pano_h_ = 0, n_poses_ = 0;
  const dvs_msgs::Event* ev_ptr_ = nullptr;
  size_t ev_size_ = 0;
  bool pending_ = false, dense_a12_ = false;
  int cost_type_ = EMBA_COST_QUADRATIC;
  double eta_ = 1.0, last_data_cost_ = 0.0;
};

}  // namespace EMBA

// Drop-in for poisson_reconstruction::reconstructFromGradient (src/image_rec/poisson_reconstruction.cpp:9-50;
// declared in include/image_rec/poisson_reconstruction.h:8), same argument and return types: the CV_64FC2 gradient
// map cv::merge({Gx, Gy}) in, the CV_64F intensity map out. The plan (sine-transform matrices of this panorama size)
// is kept between calls, as solver.cpp calls it once per recorded iteration (:417, :471).
namespace poisson_reconstruction_b200 {
inline cv::Mat reconstructFromGradient(const cv::Mat& gradients, int device = 0) {
  static emba_poisson_t plan = nullptr;
  static int plan_w = 0, plan_h = 0;
  const int H = gradients.rows, W = gradients.cols;
  if (!plan || plan_w != W || plan_h != H) {
    if (plan) emba_poisson_destroy(plan);
    plan = nullptr;
    if (emba_poisson_create(device, W, H, &plan) != EMBA_OK) throw std::runtime_error("emba_poisson_create failed");
    plan_w = W;
    plan_h = H;
  }
  std::vector<double> gx((size_t)H * W), gy((size_t)H * W);
  for (int i = 0; i < H; i++)
    for (int j = 0; j < W; j++) {
      const cv::Vec2d& g = gradients.at<cv::Vec2d>(i, j);
      gx[(size_t)i * W + j] = g[0];
      gy[(size_t)i * W + j] = g[1];
    }
  cv::Mat img(H, W, CV_64F);
  if (emba_poisson_reconstruct(plan, gx.data(), gy.data(), img.ptr<double>(0)) != EMBA_OK)
    throw std::runtime_error("emba_poisson_reconstruct failed");
  return img;
}
}  // namespace poisson_reconstruction_b200
