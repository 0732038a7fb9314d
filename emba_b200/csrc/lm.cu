// EMBA::solveTimeWindow (reference src/emba/solver.cpp:11-368): the Levenberg-Marquardt loop, with the trajectory,
// the maps, the residuals and the normal equations resident on the device. Only two scalars per iteration
// (data cost, regulariser cost) come back to the host for the accept/reject decision.
#include <cmath>

#include "emba_internal.cuh"

namespace emba {
int evaluate_slot(Handle* h, int slot, int cost_type, double eta, double alpha);
int form_normal_eq(Handle* h, int thres, int cost_type, double eta, double alpha);
int solve_schur(Handle* h, double lambda, int fix);
int solve_pcg(Handle* h, double lambda, int fix, int* iters_out, double* err_out);
int make_candidate(Handle* h, double damping, int fix);
}  // namespace emba

using namespace emba;

extern "C" int emba_solve_time_window_cb(emba_handle_t hh, const emba_lm_settings_t* s, emba_lm_log_t* log,
                                         int32_t log_cap, int32_t* n_log, double* final_cost, emba_lm_callback_t cb,
                                         void* user) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if (!s || h->t0_ns < 0) { h->err = "emba_solve_time_window: set events and the current state first"; return EMBA_E_ARG; }
  EMBA_CUDA(cudaSetDevice(h->device));
  const int cost_type = s->cost_type;
  const int fix = s->first_time_window ? 1 : 0;
  // solver.cpp:15-25
  double lambda = 1e-3;
  const double lambda_max = 1e3, lambda_min = 1e-300;
  double cost_min_old = 1e99, cost_new = cost_min_old, cost_min = cost_min_old;
  int iter = 0, count_tol_fun_sat = 0;
  bool cost_has_decreased = true;
  int nlog = 0;
  while (iter <= s->max_num_iter && cost_min > 1e-16 && lambda <= lambda_max && lambda >= lambda_min) {  // :63-64
    double ms_form = 0.0;
    if (cost_has_decreased) {
      if (iter == 0) {  // :70-92
        EMBA_TRY(evaluate_slot(h, h->cur, cost_type, s->eta, s->alpha));
        cost_min = h->st[h->cur].cost_data + h->st[h->cur].cost_reg;
      }
      // :96-102 is a no-op here: accepting a candidate swaps the device slots, so the current slot already holds
      // ep_data_new / num_ev_map_new
      EMBA_TRY(form_normal_eq(h, s->thres_valid_pixel, cost_type, s->eta, s->alpha));  // :114-130
      ms_form = h->t_ms[2];
    }
    const double lambda_used = lambda, cost_min_before = cost_min;
    int cg_it = 0;
    double cg_err = 0.0;
    EMBA_CUDA(cudaEventRecord(h->ev[10], h->stream));
    if (s->use_cg) EMBA_TRY(solve_pcg(h, lambda, fix, &cg_it, &cg_err));  // :190-202
    else EMBA_TRY(solve_schur(h, lambda, fix));
    h->solved = true;
    h->solved_fix = fix;
    EMBA_TRY(make_candidate(h, s->damping_factor, fix));  // :226-240
    EMBA_CUDA(cudaEventRecord(h->ev[11], h->stream));
    const int cand = 1 - h->cur;
    // :251-268. In strict range mode a candidate that pushes an event past the last panorama element is a rejected
    // step (infinite cost), not a failure of the window: the current state is still valid.
    bool cand_out_of_range = false;
    {
      const int rc = evaluate_slot(h, cand, cost_type, s->eta, s->alpha);
      if (rc == EMBA_E_RANGE) cand_out_of_range = true;
      else if (rc != EMBA_OK) return rc;
    }
    cost_new = cand_out_of_range ? INFINITY : h->st[cand].cost_data + h->st[cand].cost_reg;
    iter += 1;
    const bool accepted = cost_new < cost_min;  // :299
    emba_lm_log_t row;
    row.iter = iter - 1; row.lambda = lambda_used; row.cost_min = cost_min_before; row.cost_new = cost_new;
    row.accepted = accepted ? 1 : 0; row.num_active_pixels = h->Np; row.num_measurements = h->st[cand].M;
    row.cg_iters = cg_it; row.cg_error = cg_err;
    {
      float ms = 0;
      cudaEventElapsedTime(&ms, h->ev[10], h->ev[11]);
      row.ms_solve = ms;
    }
    row.ms_form = ms_form;
    row.ms_evaluate = h->t_ms[0];
    if (log && nlog < log_cap) log[nlog] = row;
    nlog++;
    const int stop = cb ? cb(&row, user) : 0;  // solver.cpp:170-179: log line / evolution images
    if (accepted) {  // :299-341
      cost_has_decreased = true;
      h->cur = cand;
      h->formed = false;
      h->solved = false;
      lambda = lambda / 10;
      cost_min_old = cost_min;
      cost_min = cost_new;
      if (fabs(1 - cost_min / (cost_min_old + 1e-10)) < s->tol_fun) {
        count_tol_fun_sat += 1;
        if (count_tol_fun_sat >= s->num_times_tol_fun_sat) break;
      }
    } else {  // :342-352
      cost_has_decreased = false;
      lambda *= 10;
      count_tol_fun_sat = 0;
    }
    if (stop) break;
  }
  if (n_log) *n_log = nlog;
  if (final_cost) *final_cost = cost_min;
  return EMBA_OK;
}

extern "C" int emba_solve_time_window(emba_handle_t hh, const emba_lm_settings_t* s, emba_lm_log_t* log,
                                      int32_t log_cap, int32_t* n_log, double* final_cost) {
  return emba_solve_time_window_cb(hh, s, log, log_cap, n_log, final_cost, nullptr, nullptr);
}
