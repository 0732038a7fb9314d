// TEST INFRASTRUCTURE ONLY (oracle/). Never linked into, imported by or called
// from the product (emba_b200/). Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library this
// file builds.
//
// C API over the UNMODIFIED reference hot-path translation units
//   /root/reference/src/emba/model.cpp
//   /root/reference/src/utils/trajectory.cpp
//   /root/reference/src/utils/event_pano_warper.cpp
//   /root/reference/src/utils/eigen_utils.cpp
// compiled where they lie (see oracle/Makefile) against oracle/shim/ and the
// reference's vendored Eigen/Sophus/basalt headers. The result,
// oracle/_ref/libemba_ref.so, is the ground truth every parity test is pinned
// to, and the "reference" CPU baseline of bench.py.
//
// The LM control flow of EMBA::solveTimeWindow (/root/reference/src/emba/
// solver.cpp:11-368) cannot be compiled here (cv_bridge, image_transport,
// rosbag, filesystem dumps), so embaref_solve_time_window() below restates it,
// citing the lines it follows, and calls the reference's own LEGM methods for
// every numerical step.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <set>
#include <string>
#include <vector>

// The per-event Jacobian pieces live in LEGM's private event_map_
// (include/emba/model.h:131-132). The parity tests need them, so this TU (and
// only this TU; the reference TUs are compiled untouched) opens the class up.
// Everything model.h pulls in is included first (all #pragma once), so the
// access override only touches emba/model.h and utils/event_pano_warper.h.
#include <array>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <unordered_map>
#include <Eigen/Dense>
#include <Eigen/Sparse>
#include "emba/event_map.h"
#include "utils/trajectory.h"
#include "utils/equirectangular_camera.h"
#include "image_rec/poisson_reconstruction.h"
#include <sensor_msgs/CameraInfo.h>
#include <image_geometry/pinhole_camera_model.h>
#define private public
#define protected public
#include "emba/model.h"
#undef private
#undef protected

using EMBA::LEGM;
using EMBA::MatXd;
using EMBA::Mat2d;
using EMBA::VecXd;
using EMBA::EventPacket;

namespace {

struct RefHandle {
  LEGM* model = nullptr;
  EventPacket events;
  int sensor_w = 0, sensor_h = 0, pano_w = 0, pano_h = 0;
  // last evaluation
  VecXd ep;
  cv::Mat num_ev_map;
  // last normal equations
  MatXd A11, A12;
  std::vector<Mat2d> A22;
  VecXd b1, b2;
  std::set<size_t> active, inactive;
  int n_poses = 0;
  // timers (seconds) around the three regions the reference instruments
  // (solver.cpp:105-151 form, :181-222 solve, :242-294 objective)
  double t_form = 0, t_solve = 0, t_obj = 0;
  long c_form = 0, c_solve = 0, c_obj = 0;
};

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::high_resolution_clock::now().time_since_epoch()).count();
}

cv::Mat wrapCopy(const double* src, int rows, int cols) {
  cv::Mat m(rows, cols, CV_64FC1);
  std::memcpy(m.data, src, sizeof(double) * (size_t)rows * cols);
  return m;
}

LinearTrajectory* makeTraj(double t_beg, double dt_knots, int n, const double* quat_xyzw) {
  std::vector<Sophus::SO3d> cps;
  cps.reserve(n);
  for (int i = 0; i < n; i++) {
    Eigen::Quaterniond q(quat_xyzw[4 * i + 3], quat_xyzw[4 * i + 0], quat_xyzw[4 * i + 1], quat_xyzw[4 * i + 2]);
    cps.emplace_back(q);  // Sophus normalises
  }
  // the constructor the LM's clone()/cloneSegment() use (trajectory.cpp:61-74)
  return new LinearTrajectory(t_beg, dt_knots, cps);
}

void readTraj(Trajectory* t, double* quat_xyzw) {
  const int n = (int)t->size();
  for (int i = 0; i < n; i++) {
    const Eigen::Quaterniond q = t->getControlPose(i).unit_quaternion();
    quat_xyzw[4 * i + 0] = q.x();
    quat_xyzw[4 * i + 1] = q.y();
    quat_xyzw[4 * i + 2] = q.z();
    quat_xyzw[4 * i + 3] = q.w();
  }
}

}  // namespace

extern "C" {

void* embaref_create(int sensor_w, int sensor_h, double fx, double fy, double cx, double cy, double C_th,
                     int pano_w, int pano_h) {
  sensor_msgs::CameraInfo ci;
  ci.width = sensor_w;
  ci.height = sensor_h;
  ci.distortion_model = "plumb_bob";
  ci.D = {0, 0, 0, 0, 0};
  ci.K = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  ci.R = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  ci.P = {fx, 0, cx, 0, 0, fy, cy, 0, 0, 0, 1, 0};
  RefHandle* h = new RefHandle();
  h->model = new LEGM(ci, C_th, pano_w, pano_h);  // model.cpp:56-70
  h->sensor_w = sensor_w; h->sensor_h = sensor_h; h->pano_w = pano_w; h->pano_h = pano_h;
  h->num_ev_map = cv::Mat::zeros(pano_h, pano_w, CV_32SC1);
  return h;
}

void embaref_destroy(void* hv) {
  RefHandle* h = (RefHandle*)hv;
  if (!h) return;
  delete h->model;
  delete h;
}

// bearing LUT the reference precomputes (event_pano_warper.cpp:27-41); out: [H_s*W_s*3]
void embaref_get_bearing_lut(void* hv, double* out) {
  RefHandle* h = (RefHandle*)hv;
  const auto& v = h->model->event_warper_ptr_->precomputed_bearing_vectors_;
  for (size_t i = 0; i < v.size(); i++) { out[3 * i] = v[i].x; out[3 * i + 1] = v[i].y; out[3 * i + 2] = v[i].z; }
}

void embaref_set_events(void* hv, long N, const uint16_t* x, const uint16_t* y, const int64_t* t_ns,
                        const uint8_t* pol) {
  RefHandle* h = (RefHandle*)hv;
  h->events.resize(N);
  for (long i = 0; i < N; i++) {
    dvs_msgs::Event& e = h->events[i];
    e.x = x[i]; e.y = y[i]; e.polarity = pol[i];
    e.ts.fromNSec((uint64_t)t_ns[i]);
  }
}

void* embaref_traj_create(double t_beg, double dt_knots, int n, const double* quat_xyzw) {
  return makeTraj(t_beg, dt_knots, n, quat_xyzw);
}
void embaref_traj_destroy(void* t) { delete (Trajectory*)t; }
int embaref_traj_size(void* t) { return (int)((Trajectory*)t)->size(); }
void embaref_traj_get(void* t, double* quat_xyzw) { readTraj((Trajectory*)t, quat_xyzw); }

// Trajectory::evaluate (trajectory.cpp:122-147): R row-major [9], jac row-major 3x6 [18]
int embaref_traj_evaluate(void* tv, int64_t t_ns, double* R_out, double* jac_out) {
  Trajectory* t = (Trajectory*)tv;
  ros::Time tt; tt.fromNSec((uint64_t)t_ns);
  int idx = -1;
  cv::Mat J;
  Sophus::SO3d R = t->evaluate(tt, &idx, &J);
  Eigen::Matrix3d Rm = R.matrix();
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R_out[3 * i + j] = Rm(i, j);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 6; j++) jac_out[6 * i + j] = J.at<double>(i, j);
  return idx;
}

// Model::updateTraj (model.cpp:22-53)
void embaref_update_traj(void* hv, void* tv, const double* x1, int len, int fix_first) {
  RefHandle* h = (RefHandle*)hv;
  VecXd v = Eigen::Map<const VecXd>(x1, len);
  if (fix_first) h->model->updateTraj((Trajectory*)tv, v, 1);
  else h->model->updateTraj((Trajectory*)tv, v);
}

// LEGM::evaluateDataError (model.cpp:72-258). Returns the number of inlier
// measurements M; ep_out (capacity >= N) and num_ev_map_out ([H_p*W_p] int32)
// may be NULL.
long embaref_evaluate(void* hv, void* tv, const double* Gx, const double* Gy, int eval_deriv, double* ep_out,
                      int32_t* num_ev_map_out) {
  RefHandle* h = (RefHandle*)hv;
  cv::Mat mGx = wrapCopy(Gx, h->pano_h, h->pano_w), mGy = wrapCopy(Gy, h->pano_h, h->pano_w);
  h->ep = h->model->evaluateDataError((Trajectory*)tv, mGx, mGy, h->events, eval_deriv != 0, h->num_ev_map);
  if (ep_out) std::memcpy(ep_out, h->ep.data(), sizeof(double) * h->ep.size());
  if (num_ev_map_out) std::memcpy(num_ev_map_out, h->num_ev_map.data, sizeof(int32_t) * h->num_ev_map.total());
  return (long)h->ep.size();
}

// Per-measurement dump of the state left in event_map_ by the last
// evaluateDataError(eval_deriv=true), in the reference's own measurement order
// (sensor pixel row-major, then time: model.cpp:179-246). One row per INLIER
// measurement:
//   rec[0]=e  rec[1..2]=dp  rec[3..4]=pm_curr  rec[5..6]=Gpm  rec[7..8]=temp
//   rec[9..14]=Jc = temp*dpm_ddrot_cp(curr)  (model.cpp:449)
//   rec[15..20]=Jp = -Gpm*dpm_ddrot_cp(prev) (model.cpp:459)
//   rec[21]=cp_idx(curr) rec[22]=cp_idx(prev) rec[23]=sensor pixel index
//   rec[24]=rank of the current event among the events of its sensor pixel (k)
// Also counts outliers. Returns number of rows written (== M).
long embaref_dump_measurements(void* hv, double* rec, long cap, long* n_outliers) {
  RefHandle* h = (RefHandle*)hv;
  auto& em = h->model->event_map_;
  long m = 0, outl = 0;
  for (int y = 0; y < em.height(); y++)
    for (int x = 0; x < em.width(); x++) {
      std::vector<EMBA::State_LEGM>& v = em.at(x, y);
      for (size_t k = 1; k < v.size(); k++) {
        EMBA::State_LEGM& c = v[k];
        EMBA::State_LEGM& p = v[k - 1];
        if (c.inlier_idx < 0) { outl++; continue; }
        if (m >= cap) return -1;
        double* r = rec + 25 * m;
        r[0] = h->ep(c.inlier_idx);
        r[1] = c.dp(0); r[2] = c.dp(1);
        r[3] = c.pm.x; r[4] = c.pm.y;
        r[5] = c.Gpm(0); r[6] = c.Gpm(1);
        r[7] = c.temp(0); r[8] = c.temp(1);
        EMBA::RowVector6d Jc = c.temp * c.dpm_ddrot_cp;
        EMBA::RowVector6d Jp = -c.Gpm * p.dpm_ddrot_cp;
        for (int i = 0; i < 6; i++) { r[9 + i] = Jc(i); r[15 + i] = Jp(i); }
        r[21] = c.cp_idx; r[22] = p.cp_idx;
        r[23] = (double)(y * em.width() + x);
        r[24] = (double)k;
        m++;
      }
    }
  if (n_outliers) *n_outliers = outl;
  return m;
}

// evaluateRegError + the cost of solver.cpp:90: 0.5*alpha*||ep_reg||^2
double embaref_reg_cost(void* hv, const double* Gx, const double* Gy, double alpha) {
  RefHandle* h = (RefHandle*)hv;
  cv::Mat mGx = wrapCopy(Gx, h->pano_h, h->pano_w), mGy = wrapCopy(Gy, h->pano_h, h->pano_w);
  VecXd r = h->model->evaluateRegError(mGx, mGy);
  return alpha * 0.5 * r.dot(r);
}

// data cost of the last evaluation: solver.cpp:82-89. irls_type: 0 none, 1 cauchy, 2 huber
double embaref_data_cost(void* hv, int irls_type, double a) {
  RefHandle* h = (RefHandle*)hv;
  if (irls_type == 0) return 0.5 * h->ep.dot(h->ep);
  return h->model->evaluateRobustDataCost(h->ep, irls_type == 1 ? "cauchy" : "huber", a);
}

// formNormalEq / formNormalEqIRLS (model.cpp:316-687) + applyL2Reg (:689-719)
// on the state left by the last evaluation. Returns Np.
long embaref_form(void* hv, int n_poses, int thres, int irls_type, double a, double alpha, int apply_l2,
                  const double* Gx, const double* Gy) {
  RefHandle* h = (RefHandle*)hv;
  h->n_poses = n_poses;
  if (irls_type == 0)
    h->model->formNormalEq(h->A11, h->A12, h->A22, h->b1, h->b2, h->ep, n_poses, h->num_ev_map, thres, h->active,
                           h->inactive);
  else
    h->model->formNormalEqIRLS(h->A11, h->A12, h->A22, h->b1, h->b2, h->ep, n_poses, h->num_ev_map, thres,
                               h->active, h->inactive, irls_type == 1 ? "cauchy" : "huber", a);
  if (apply_l2) {
    cv::Mat mGx = wrapCopy(Gx, h->pano_h, h->pano_w), mGy = wrapCopy(Gy, h->pano_h, h->pano_w);
    h->model->applyL2Reg(h->A22, h->b2, h->active, alpha, mGx, mGy);
  }
  return (long)h->active.size();
}

// copies (row-major): A11 [3n*3n], A12 [3n*2Np] (may be NULL), A22 [Np*4], b1 [3n], b2 [2Np], active [Np]
void embaref_get_normal_eq(void* hv, double* A11, double* A12, double* A22, double* b1, double* b2,
                           int64_t* active) {
  RefHandle* h = (RefHandle*)hv;
  const long d = h->A11.rows();
  const long c2 = h->A12.cols();
  if (A11) for (long i = 0; i < d; i++) for (long j = 0; j < d; j++) A11[i * d + j] = h->A11(i, j);
  if (A12) for (long i = 0; i < d; i++) for (long j = 0; j < c2; j++) A12[i * c2 + j] = h->A12(i, j);
  if (A22) for (size_t i = 0; i < h->A22.size(); i++) {
    A22[4 * i + 0] = h->A22[i](0, 0); A22[4 * i + 1] = h->A22[i](0, 1);
    A22[4 * i + 2] = h->A22[i](1, 0); A22[4 * i + 3] = h->A22[i](1, 1);
  }
  if (b1) std::memcpy(b1, h->b1.data(), sizeof(double) * h->b1.size());
  if (b2) std::memcpy(b2, h->b2.data(), sizeof(double) * h->b2.size());
  if (active) { size_t j = 0; for (auto p : h->active) active[j++] = (int64_t)p; }
}

// nnz of the reference's dense A12 (for the algorithmic-bytes accounting)
long embaref_a12_nnz(void* hv) {
  RefHandle* h = (RefHandle*)hv;
  long nnz = 0;
  const double* p = h->A12.data();
  const long tot = h->A12.size();
  for (long i = 0; i < tot; i++) nnz += (p[i] != 0.0);
  return nnz;
}

// solveNormalEq (model.cpp:721-792) or solveNormalEqCG (:794-840) on the stored
// system; fix_first applies the gauge shrink of solver.cpp:156-165 first.
// x1_out has 3n (or 3(n-1)) entries, x2_out 2Np.
void embaref_solve(void* hv, double lambda, int use_cg, int fix_first, double* x1_out, double* x2_out,
                   int* cg_iters, double* cg_err) {
  RefHandle* h = (RefHandle*)hv;
  MatXd A11 = h->A11, A12 = h->A12;
  VecXd b1 = h->b1;
  if (fix_first) {
    const long dl = 3 * (h->n_poses - 1);
    MatXd A11s = A11.block(3, 3, dl, dl);
    MatXd A12s = A12.block(3, 0, dl, A12.cols());
    VecXd b1s = b1.tail(dl);
    A11 = A11s; A12 = A12s; b1 = b1s;
  }
  VecXd x1, x2;
  if (!use_cg) {
    h->model->solveNormalEq(A11, A12, h->A22, b1, h->b2, lambda, x1, x2);
  } else {
    std::pair<int, double> r = h->model->solveNormalEqCG(A11, A12, h->A22, b1, h->b2, lambda, x1, x2);
    if (cg_iters) *cg_iters = r.first;
    if (cg_err) *cg_err = r.second;
  }
  if (x1_out) std::memcpy(x1_out, x1.data(), sizeof(double) * x1.size());
  if (x2_out) std::memcpy(x2_out, x2.data(), sizeof(double) * x2.size());
}

// LEGM::updateMap (model.cpp:863-903) with the stored active/inactive sets; in place
void embaref_update_map(void* hv, double* Gx, double* Gy, const double* x2, double damping) {
  RefHandle* h = (RefHandle*)hv;
  cv::Mat mGx = wrapCopy(Gx, h->pano_h, h->pano_w), mGy = wrapCopy(Gy, h->pano_h, h->pano_w);
  VecXd v = Eigen::Map<const VecXd>(x2, 2 * (long)h->active.size());
  h->model->updateMap(mGx, mGy, v, damping, h->active, h->inactive);
  std::memcpy(Gx, mGx.data, mGx.bytes());
  std::memcpy(Gy, mGy.data, mGy.bytes());
}

// Restatement of EMBA::solveTimeWindow (solver.cpp:11-368): same state
// variables, same order of calls into the reference's LEGM, same accept/reject
// and termination rules. I/O (saveEvoData/saveOptData, runtime_*.txt) is
// dropped. The trajectory handle is updated in place (its pointee is replaced
// on acceptance exactly like `Trajectory*& traj_ptr`), Gx/Gy are in/out.
// log rows (cap rows of 6): iter, lambda, cost_min, cost_new(after evaluation), accepted, Np
// Returns the number of log rows (= number of solves).
int embaref_solve_time_window(void* hv, void** traj_io, double* Gx_io, double* Gy_io, int max_num_iter,
                              double tol_fun, int num_times_tol_fun_sat, int use_cg, int irls_type, double eta,
                              int thres_valid_pixel, double damping_factor, double alpha, int first_time_window,
                              double* log, int log_cap, double* final_cost) {
  RefHandle* h = (RefHandle*)hv;
  LEGM* model = h->model;
  Trajectory* traj_ptr = (Trajectory*)(*traj_io);
  cv::Mat Gx = wrapCopy(Gx_io, h->pano_h, h->pano_w), Gy = wrapCopy(Gy_io, h->pano_h, h->pano_w);
  const EventPacket& event_subset = h->events;
  const std::string cost_type = irls_type == 1 ? "cauchy" : (irls_type == 2 ? "huber" : "quadratic");
  const bool use_IRLS = irls_type != 0;
  h->t_form = h->t_solve = h->t_obj = 0; h->c_form = h->c_solve = h->c_obj = 0;

  // solver.cpp:15-25
  double lambda = 1e-3, lambda_max = 1e3, lambda_min = 1e-300;
  double cost_min_old = 1e99, cost_new = cost_min_old, cost_min = cost_min_old;
  int iter = 0, count_tol_fun_sat = 0;
  bool cost_has_decreased = true;
  // solver.cpp:28-48
  VecXd ep_data, ep_data_new, ep_reg, ep_reg_new;
  double cost_data = 0, cost_reg = 0, cost_data_new = 0, cost_reg_new = 0;
  MatXd A11, A12;
  std::vector<Mat2d> A22_blocks;
  VecXd b1, b2, x1, x2;
  const int num_ctrl_poses = (int)traj_ptr->size();
  cv::Mat num_ev_map = cv::Mat::zeros(Gx.rows, Gx.cols, CV_32SC1);
  cv::Mat num_ev_map_new = cv::Mat::zeros(Gx.rows, Gx.cols, CV_32SC1);
  std::set<size_t> active_pix_idxes, inactive_pix_idxes;
  int nlog = 0;
  bool converged = false;

  // solver.cpp:63-64
  while (iter <= max_num_iter && cost_min > 1e-16 && lambda <= lambda_max && lambda >= lambda_min) {
    if (cost_has_decreased) {
      if (iter == 0) {
        // solver.cpp:75-91
        ep_data = model->evaluateDataError(traj_ptr, Gx, Gy, event_subset, true, num_ev_map);
        ep_reg = model->evaluateRegError(Gx, Gy);
        if (use_IRLS) cost_data = model->evaluateRobustDataCost(ep_data, cost_type, eta);
        else cost_data = 0.5 * ep_data.dot(ep_data);
        cost_reg = alpha * 0.5 * ep_reg.dot(ep_reg);
        cost_min = cost_data + cost_reg;
      } else {
        // solver.cpp:96-102
        ep_data = ep_data_new;
        ep_reg = ep_reg_new;
        num_ev_map_new.copyTo(num_ev_map);
      }
      const double t1 = now_s();
      // solver.cpp:114-130
      if (use_IRLS)
        model->formNormalEqIRLS(A11, A12, A22_blocks, b1, b2, ep_data, num_ctrl_poses, num_ev_map,
                                thres_valid_pixel, active_pix_idxes, inactive_pix_idxes, cost_type, eta);
      else
        model->formNormalEq(A11, A12, A22_blocks, b1, b2, ep_data, num_ctrl_poses, num_ev_map, thres_valid_pixel,
                            active_pix_idxes, inactive_pix_idxes);
      model->applyL2Reg(A22_blocks, b2, active_pix_idxes, alpha, Gx, Gy);
      h->t_form += now_s() - t1; h->c_form++;
      // solver.cpp:156-165
      if (first_time_window) {
        const size_t dim_pose_left = 3 * (num_ctrl_poses - 1);
        MatXd A11_1st = A11.block(3, 3, dim_pose_left, dim_pose_left);
        MatXd A12_1st = A12.block(3, 0, dim_pose_left, A12.cols());
        VecXd b1_1st = b1.tail(dim_pose_left);
        A11 = A11_1st; A12 = A12_1st; b1 = b1_1st;
      }
    }
    const double lambda_used = lambda, cost_min_before = cost_min;
    // solver.cpp:190-202
    const double t2 = now_s();
    if (!use_cg) model->solveNormalEq(A11, A12, A22_blocks, b1, b2, lambda, x1, x2);
    else model->solveNormalEqCG(A11, A12, A22_blocks, b1, b2, lambda, x1, x2);
    h->t_solve += now_s() - t2; h->c_solve++;
    // solver.cpp:226-240
    Trajectory* traj_new_ptr = traj_ptr->clone();
    if (first_time_window) model->updateTraj(traj_new_ptr, x1, 1);
    else model->updateTraj(traj_new_ptr, x1);
    cv::Mat Gx_new = Gx.clone(), Gy_new = Gy.clone();
    model->updateMap(Gx_new, Gy_new, x2, damping_factor, active_pix_idxes, inactive_pix_idxes);
    // solver.cpp:251-271
    const double t3 = now_s();
    ep_data_new = model->evaluateDataError(traj_new_ptr, Gx_new, Gy_new, event_subset, true, num_ev_map_new);
    ep_reg_new = model->evaluateRegError(Gx_new, Gy_new);
    if (use_IRLS) cost_data_new = model->evaluateRobustDataCost(ep_data_new, cost_type, eta);
    else cost_data_new = 0.5 * ep_data_new.dot(ep_data_new);
    cost_reg_new = alpha * 0.5 * ep_reg_new.dot(ep_reg_new);
    cost_new = cost_data_new + cost_reg_new;
    iter += 1;
    h->t_obj += now_s() - t3; h->c_obj++;

    const bool accepted = cost_new < cost_min;
    if (nlog < log_cap) {
      double* r = log + 6 * nlog;
      r[0] = iter - 1; r[1] = lambda_used; r[2] = cost_min_before; r[3] = cost_new; r[4] = accepted ? 1 : 0;
      r[5] = (double)active_pix_idxes.size();
    }
    nlog++;
    // solver.cpp:299-352
    if (accepted) {
      cost_has_decreased = true;
      delete traj_ptr;
      traj_ptr = traj_new_ptr;
      Gx_new.copyTo(Gx);
      Gy_new.copyTo(Gy);
      lambda = lambda / 10;
      cost_min_old = cost_min;
      cost_min = cost_new;
      cost_data = cost_data_new;
      cost_reg = cost_reg_new;
      if (std::abs(1 - cost_min / (cost_min_old + 1e-10)) < tol_fun) {
        count_tol_fun_sat = count_tol_fun_sat + 1;
        if (count_tol_fun_sat >= num_times_tol_fun_sat) { converged = true; break; }
      }
    } else {
      delete traj_new_ptr;  // the reference leaks this clone (solver.cpp:342-352)
      cost_has_decreased = false;
      lambda *= 10;
      count_tol_fun_sat = 0;
    }
  }
  (void)converged; (void)cost_data; (void)cost_reg;
  *traj_io = traj_ptr;
  std::memcpy(Gx_io, Gx.data, Gx.bytes());
  std::memcpy(Gy_io, Gy.data, Gy.bytes());
  if (final_cost) *final_cost = cost_min;
  return nlog;
}

// LinearTrajectory::generateCtrlPosesLong (src/utils/trajectory.cpp:258-294 -> generateCtrlPoses :245-256 ->
// fitCtrlPoses :149-229): control-pose initialisation from a dense front-end trajectory, as EMBA::Run calls it
// (src/emba/emba.cpp:416, sub-interval length = dt_knots). poses: time-sorted (t_ns, quat xyzw).
// Returns the number of control poses written to out_quat (capacity cap poses), or -1.
int embaref_generate_ctrl_poses_long(long n_poses, const int64_t* t_ns, const double* quat_xyzw, double t_beg,
                                     double t_end, double dt_knots, double sub_interval, double* out_quat, int cap) {
  PoseMap poses;
  for (long i = 0; i < n_poses; i++) {
    ros::Time t; t.fromNSec((uint64_t)t_ns[i]);
    Eigen::Quaterniond q(quat_xyzw[4 * i + 3], quat_xyzw[4 * i], quat_xyzw[4 * i + 1], quat_xyzw[4 * i + 2]);
    poses.insert(poses.end(), std::make_pair(t, Sophus::SO3d(q)));
  }
  TrajectorySettings cfg;
  cfg.t_beg = ros::Time(t_beg);
  cfg.t_end = ros::Time(t_end);
  cfg.dt_knots = dt_knots;
  LinearTrajectory traj(cfg);
  std::vector<Sophus::SO3d> cps = traj.generateCtrlPosesLong(poses, ros::Time(t_beg), ros::Time(t_end), sub_interval);
  if ((int)cps.size() > cap) return -1;
  for (size_t i = 0; i < cps.size(); i++) {
    const Eigen::Quaterniond q = cps[i].unit_quaternion();
    out_quat[4 * i] = q.x(); out_quat[4 * i + 1] = q.y(); out_quat[4 * i + 2] = q.z(); out_quat[4 * i + 3] = q.w();
  }
  return (int)cps.size();
}

// poisson_reconstruction::reconstructFromGradient (src/image_rec/poisson_reconstruction.cpp:9-50) on the two-channel
// gradient map the solver builds with cv::merge({Gx, Gy}) (src/emba/solver.cpp:412-417). Gx, Gy, out: H x W row-major.
int embaref_poisson_reconstruct(const double* Gx, const double* Gy, int H, int W, double* out) {
  cv::Mat g(H, W, CV_64FC2);
  for (int i = 0; i < H; i++)
    for (int j = 0; j < W; j++) {
      g.at<cv::Vec2d>(i, j)[0] = Gx[(size_t)i * W + j];
      g.at<cv::Vec2d>(i, j)[1] = Gy[(size_t)i * W + j];
    }
  const cv::Mat m = poisson_reconstruction::reconstructFromGradient(g);
  if (m.rows != H || m.cols != W) return -1;
  std::memcpy(out, m.data, sizeof(double) * (size_t)H * W);
  return 0;
}

// timers of the last embaref_solve_time_window: seconds and call counts
void embaref_get_timers(void* hv, double* t3, long* c3) {
  RefHandle* h = (RefHandle*)hv;
  t3[0] = h->t_form; t3[1] = h->t_solve; t3[2] = h->t_obj;
  c3[0] = h->c_form; c3[1] = h->c_solve; c3[2] = h->c_obj;
}

}  // extern "C"
