// EMBA::solveTimeWindow (reference src/emba/solver.cpp:11-368): the Levenberg-Marquardt loop, with the trajectory,
// the maps, the residuals and the normal equations resident on the device. Only two scalars per iteration
// (data cost, regulariser cost) come back to the host for the accept/reject decision.
#include "emba_internal.cuh"

namespace emba {
int evaluate_slot(Handle* h, int slot, int cost_type, double eta, double alpha);
int form_normal_eq(Handle* h, int thres, int cost_type, double eta, double alpha);
int solve_schur(Handle* h, double lambda, int fix);
int solve_pcg(Handle* h, double lambda, int fix, int* iters_out, double* err_out);
int make_candidate(Handle* h, double damping, int fix);
}  // namespace emba

using namespace emba;

extern "C" int emba_solve_time_window(emba_handle_t hh, const emba_lm_settings_t* s, emba_lm_log_t* log,
                                      int32_t log_cap, int32_t* n_log, double* final_cost) {
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
    if (cost_has_decreased) {
      if (iter == 0) {  // :70-92
        EMBA_TRY(evaluate_slot(h, h->cur, cost_type, s->eta, s->alpha));
        cost_min = h->st[h->cur].cost_data + h->st[h->cur].cost_reg;
      }
      // :96-102 is a no-op here: accepting a candidate swaps the device slots, so the current slot already holds
      // ep_data_new / num_ev_map_new
      EMBA_TRY(form_normal_eq(h, s->thres_valid_pixel, cost_type, s->eta, s->alpha));  // :114-130
    }
    const double lambda_used = lambda, cost_min_before = cost_min;
    if (s->use_cg) EMBA_TRY(solve_pcg(h, lambda, fix, nullptr, nullptr));  // :190-202
    else EMBA_TRY(solve_schur(h, lambda, fix));
    h->solved = true;
    h->solved_fix = fix;
    EMBA_TRY(make_candidate(h, s->damping_factor, fix));  // :226-240
    const int cand = 1 - h->cur;
    EMBA_TRY(evaluate_slot(h, cand, cost_type, s->eta, s->alpha));  // :251-268
    cost_new = h->st[cand].cost_data + h->st[cand].cost_reg;
    iter += 1;
    const bool accepted = cost_new < cost_min;  // :299
    if (log && nlog < log_cap) {
      emba_lm_log_t& r = log[nlog];
      r.iter = iter - 1; r.lambda = lambda_used; r.cost_min = cost_min_before; r.cost_new = cost_new;
      r.accepted = accepted ? 1 : 0; r.num_active_pixels = h->Np; r.num_measurements = h->st[cand].M;
    }
    nlog++;
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
  }
  if (n_log) *n_log = nlog;
  if (final_cost) *final_cost = cost_min;
  return EMBA_OK;
}
