/*
 * emba_b200 C ABI -- B200 (sm_100a) drop-in for EMBA's per-iteration
 * Levenberg-Marquardt hot path.
 *
 * Each entry point replaces one member of the reference's measurement-model
 * class `EMBA::LEGM : Model` (reference: include/emba/model.h:26-133), whose
 * only caller is `EMBA::solveTimeWindow` (reference: src/emba/solver.cpp:11-368).
 * The reference-side binding (a header-only adapter with LEGM's signatures on
 * top of these calls) is shown in INTEGRATION.md.
 *
 * Conventions (identical to the reference so results compare 1:1):
 *  - events are time-sorted; only the first 100*floor(N/100) are used
 *    (src/emba/model.cpp:78-79,102);
 *  - maps are row-major fp64, index = x + y*pano_w (include/emba/model.h:42-43);
 *  - control poses are unit quaternions, (x, y, z, w) per pose;
 *  - pose unknowns are 3-vector LEFT perturbations stacked in pose order
 *    (src/emba/model.cpp:25-35); map unknowns are (Gx, Gy) interleaved per ACTIVE
 *    pixel in ascending pixel index (src/emba/model.cpp:438-439,874-875);
 *  - the spline time base is passed as the two int64 nanosecond values the
 *    reference forms with int64_t(1e9*t_beg), int64_t(1e9*dt_knots)
 *    (src/utils/trajectory.cpp:61-70) -- the adapter must use that expression.
 *
 * All functions return 0 on success and a negative EMBA_E_* code on failure;
 * emba_last_error() gives a message. Nothing aborts or throws across the ABI
 * (the reference aborts through glog CHECK / basalt asserts instead).
 * A handle is not thread-safe; one host thread drives one handle = one GPU.
 * There is no CPU fallback: without a CUDA device emba_create fails.
 * All pointers are HOST pointers unless a parameter name ends in `_dev`.
 */
#ifndef EMBA_B200_H_
#define EMBA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define EMBA_API __attribute__((visibility("default")))
#else
#define EMBA_API
#endif

typedef struct emba_handle_s* emba_handle_t;

enum {
  EMBA_OK = 0,
  EMBA_E_ARG = -1,      /* bad argument / call order */
  EMBA_E_CUDA = -2,     /* CUDA runtime error */
  EMBA_E_SUPPORT = -3,  /* batch time outside the spline support (basalt so3_spline.h:221-230 asserts) */
  EMBA_E_RANGE = -4,    /* strict mode only (emba_set_strict_range): a warped event rounds past the last panorama
                           element (out-of-bounds read/write in the reference, model.cpp:213,227) */
  EMBA_E_NCCL = -5,     /* collective error */
  EMBA_E_NUMERIC = -6   /* non-finite pivot in the Schur factorisation (exactly zero pivots are handled like Eigen's
                           LDLT::solve: that component of the solution is 0) */
};

/* which device-resident state a call refers to (solver.cpp keeps traj_ptr/Gx/Gy and traj_new_ptr/Gx_new/Gy_new) */
enum { EMBA_STATE_CURRENT = 0, EMBA_STATE_CANDIDATE = 1 };

/* robust cost (formNormalEqIRLS / evaluateRobustDataCost, model.cpp:279-314,493-687) */
enum { EMBA_COST_QUADRATIC = 0, EMBA_COST_CAUCHY = 1, EMBA_COST_HUBER = 2 };

typedef struct {
  int32_t sensor_w, sensor_h;   /* camera_info width/height (model.cpp:67-68) */
  int32_t pano_w, pano_h;       /* panorama size (model.cpp:64) */
  double C_th;                  /* contrast threshold (model.cpp:60) */
  const double* bearing_lut;    /* [sensor_h*sensor_w*3] rays b[y*W+x], computed by the host exactly as
                                   EventWarper::precomputeBearingVectors does (event_pano_warper.cpp:27-41) */
  int32_t device;               /* CUDA device ordinal */
} emba_config_t;

/* BASettings / LMSettings subset that drives the path (include/emba/params.h:4-61) */
typedef struct {
  int32_t max_num_iter;            /* LMSettings::max_num_iter */
  double tol_fun;                  /* LMSettings::tol_fun */
  int32_t num_times_tol_fun_sat;   /* LMSettings::num_times_tol_fun_sat */
  int32_t use_cg;                  /* BASettings::use_CG */
  int32_t cost_type;               /* EMBA_COST_*; != QUADRATIC means use_IRLS */
  double eta;                      /* BASettings::eta */
  int32_t thres_valid_pixel;       /* BASettings::thres_valid_pixel */
  double damping_factor;           /* BASettings::damping_factor */
  double alpha;                    /* BASettings::alpha */
  int32_t first_time_window;       /* EMBA::first_time_window_: fix the first control pose (solver.cpp:156-165) */
} emba_lm_settings_t;

/* one row per LM solve, the content of the reference's iteration log line (solver.cpp:170-171) and of its CG log
 * (solver.cpp:198-201) */
typedef struct {
  int32_t iter;
  double lambda;
  double cost_min;   /* before the step */
  double cost_new;   /* at the candidate */
  int32_t accepted;
  int64_t num_active_pixels;
  int64_t num_measurements;  /* inlier measurements M at the candidate */
  int32_t cg_iters;          /* solveNormalEqCG's pair (iterations, error); 0 / 0.0 for the Schur solve */
  double cg_error;
  double ms_form, ms_solve, ms_evaluate;  /* device time of this iteration's three regions, the ones the reference
                                             instruments (solver.cpp:105-151, 181-222, 242-294); ms_form = 0 after a
                                             rejected step (the equations are reused) */
} emba_lm_log_t;

/* optional per-iteration hook of emba_solve_time_window_cb: called on the host after every LM solve with that
 * iteration's log row (the point where the reference prints its log line and, with record_data, dumps the evolution
 * images, solver.cpp:170-179). The callback may call emba_get_state / emba_reconstruct_map on the handle (which = 1
 * is the candidate just evaluated); returning non-zero stops the loop after this iteration. */
typedef int (*emba_lm_callback_t)(const emba_lm_log_t* row, void* user);

/* ---- lifetime: replaces LEGM::LEGM / ~LEGM (model.cpp:56-70, model.h:80) ---- */
EMBA_API int emba_create(const emba_config_t* cfg, emba_handle_t* out);
EMBA_API int emba_destroy(emba_handle_t h);
EMBA_API const char* emba_last_error(emba_handle_t h);

/* ---- "next" row N1: device-resident event sequence. The whole recording is uploaded once
 * (pinned, chunked staging); the steps the reference does on the host before every window run on the device:
 *   emba_events_sort_by_time : std::sort by timestamp after parsing the bag (src/utils/rosbag_loading.cpp:61-65;
 *                              stable here, std::sort's order of equal stamps is unspecified)
 *   emba_events_subsample    : keep every rate-th event, the rate-th first (src/emba/emba.cpp:281-304)
 *   emba_events_window       : EMBA::getEventSubset (src/emba/emba.cpp:473-510): 1 ms robust margins, 100-event
 *                              stride search, including its unsigned wrap-around when the first probe already lies
 *                              past the window end; t_beg_ns / t_end_ns are ros::Time::toNSec()
 *   emba_set_events_dev      : the per-window pre-pass (what emba_set_events does) straight from the sequence,
 *                              events [idx_beg, idx_end), no host copy */
typedef struct emba_events_s* emba_events_t;
EMBA_API int emba_events_create(int32_t device, int64_t n_events, const uint16_t* x, const uint16_t* y,
                                const int64_t* t_ns, const uint8_t* polarity, emba_events_t* out);
EMBA_API int emba_events_destroy(emba_events_t ev);
EMBA_API int emba_events_count(emba_events_t ev, int64_t* out);
EMBA_API int emba_events_sort_by_time(emba_events_t ev);
EMBA_API int emba_events_subsample(emba_events_t ev, int32_t event_sampling_rate);
EMBA_API int emba_events_window(emba_events_t ev, int64_t t_beg_ns, int64_t t_end_ns, int64_t* idx_beg,
                                int64_t* idx_end);
EMBA_API int emba_events_download(emba_events_t ev, int64_t idx_beg, int64_t idx_end, uint16_t* x, uint16_t* y,
                                  int64_t* t_ns, uint8_t* polarity);
EMBA_API int emba_set_events_dev(emba_handle_t h, emba_events_t ev, int64_t idx_beg, int64_t idx_end);

/* ---- events: replaces the `const EventPacket& events` argument of evaluateDataError (model.h:83-84) and the
 * per-pixel EventMap pairing structure (include/emba/event_map.h:22-113). Called once per time window: uploads
 * the events, computes batch mid-times (model.cpp:115-119) and the static pair links. */
EMBA_API int emba_set_events(emba_handle_t h, int64_t n_events, const uint16_t* x, const uint16_t* y, const int64_t* t_ns,
                    const uint8_t* polarity);
/* number of event pairs (candidate measurements, before the outlier gate) */
EMBA_API int emba_num_pairs(emba_handle_t h, int64_t* out);

/* ---- time-sharding across GPUs: this handle evaluates the measurements of shard `rank`
 * of `world` (contiguous control-pose / time slices). Call before emba_set_state. Partial results are combined
 * by emba_comm_* below when a communicator is attached; with world == 1 it is a no-op. */
EMBA_API int emba_set_shard(emba_handle_t h, int32_t rank, int32_t world);
/* attach an NCCL communicator: `nccl_unique_id` is the 128-byte ncclUniqueId created by rank 0 and broadcast by
 * the host (e.g. torch.distributed); collectives run over NVLink/NVSwitch inside the handle's stream. */
EMBA_API int emba_comm_unique_id(void* nccl_unique_id_out128);
EMBA_API int emba_comm_init(emba_handle_t h, const void* nccl_unique_id128, int32_t rank, int32_t world);

/* ---- state: replaces the (Trajectory*, cv::Mat Gx, cv::Mat Gy) arguments. Reads what the adapter gets through
 * Trajectory::size/getControlPose/begTime/getDtCtrlPoses (include/utils/trajectory.h:30-47). */
EMBA_API int emba_set_state(emba_handle_t h, int32_t which, int64_t t0_ns, int64_t dt_ns, int32_t n_poses,
                   const double* quat_xyzw, const double* Gx, const double* Gy);
EMBA_API int emba_get_state(emba_handle_t h, int32_t which, double* quat_xyzw_out, double* Gx_out, double* Gy_out);

/* ---- LEGM::evaluateDataError (model.cpp:72-258) + evaluateRegError (:260-277) + the cost of solver.cpp:82-91,
 * 257-268. cost_data = 0.5*sum(e^2) or the robust cost; cost_reg = 0.5*alpha*sum(Gx^2+Gy^2) over ALL pixels.
 * Per-measurement residuals, displacements and target pixels stay on the device for emba_form_normal_eq. */
EMBA_API int emba_evaluate(emba_handle_t h, int32_t which, int32_t cost_type, double eta, double alpha, double* cost_data,
                  double* cost_reg, int64_t* num_measurements);
/* optional downloads of the last evaluation of `which` (parity / the adapter's VecXd return value):
 * ep_out [M] in the reference's order (sensor pixel row-major, then time; model.cpp:179-246),
 * num_ev_map_out [pano_h*pano_w] int32 (model.cpp:227). Either may be NULL. */
EMBA_API int emba_get_evaluation(emba_handle_t h, int32_t which, double* ep_out, int32_t* num_ev_map_out);
/* A warped event can round to column pano_w (within half a pixel of the phi = +-pi seam). The reference's release
 * build then reads and writes element (y, pano_w) of its row-major cv::Mat, i.e. (y + 1, 0) (model.cpp:213-227,
 * 396-412): the default here reproduces exactly that linear index. Only an index past the LAST element (row
 * pano_h - 1 at the seam, or row pano_h at the pole) is undefined in the reference; such pairs are dropped as
 * outliers. on != 0 makes emba_evaluate fail with EMBA_E_RANGE on those instead (all ranks agree on it). */
EMBA_API int emba_set_strict_range(emba_handle_t h, int32_t on);

/* ---- LEGM::formNormalEq / formNormalEqIRLS (model.cpp:316-687) fused with applyL2Reg (:689-719), on the last
 * evaluation of the CURRENT state. */
EMBA_API int emba_form_normal_eq(emba_handle_t h, int32_t thres_valid_pixel, int32_t cost_type, double eta, double alpha,
                        int64_t* num_active_pixels);
/* Map-block reduction path of emba_form_normal_eq: 0 (default) = deterministic segmented reduction over
 * pixel-sorted rows (stable radix sort + one warp per pixel, bit-reproducible); 1 = fp64-atomic path (every row
 * adds its 24 A12 + 5 A22/b2 products with red.global.add.f64, no sort; same values up to summation order). */
enum { EMBA_MAP_SORTED = 0, EMBA_MAP_ATOMIC = 1 };
EMBA_API int emba_set_map_path(emba_handle_t h, int32_t mode);

/* ---- LEGM::applyL2Reg (model.cpp:689-719) as a separate step, for callers that form with alpha = 0 (the
 * reference calls it right after formNormalEq, solver.cpp:130): A22 += alpha*I, b2 -= alpha*(Gx,Gy)[active]. */
EMBA_API int emba_apply_l2_reg(emba_handle_t h, double alpha);
/* parity / adapter download. A11 [3n*3n] row-major, b1 [3n], A22 [Np*4] (2x2 row-major blocks), b2 [2Np],
 * active [Np] ascending pixel indices (the std::set order, model.cpp:370-377). A12 is returned dense row-major
 * [3n * 2Np] exactly like the reference's MatXd (model.cpp:358) -- only for small problems. Any may be NULL. */
EMBA_API int emba_get_normal_eq(emba_handle_t h, double* A11, double* b1, double* A22, double* b2, int64_t* active,
                       double* A12_dense);
/* parity download of the per-measurement Jacobian rows of the last emba_form_normal_eq, in the reference's
 * measurement order (inlier index = position in ep; model.cpp:179-246): 16 doubles per row
 * [Jc(6) = temp*dpm_ddrot_cp (model.cpp:449) | Jp(6) = -Gpm*dpm_ddrot_cp (model.cpp:459) | e | dp(2) | 0].
 * Rows of measurements on inactive pixels are zero (the reference skips them, model.cpp:409-412): form with
 * thres_valid_pixel = 1 to get every inlier's row. Single GPU, windows of at most 2^24 pairs (EMBA_E_SUPPORT). */
EMBA_API int emba_get_jacobian_rows(emba_handle_t h, double* rows_out, int64_t cap_rows, int64_t* n_rows);
/* number of structurally non-zero A12 entries held on the device (windowed per-pixel strips) */
EMBA_API int emba_a12_entries(emba_handle_t h, int64_t* out);

/* ---- LEGM::solveNormalEq (model.cpp:721-792, Schur complement onto the control poses) or
 * LEGM::solveNormalEqCG (:794-840, Jacobi-PCG, <=100 iterations, tolerance 1e-6). fix_first_pose applies the
 * gauge shrink of solver.cpp:156-165. x1_out [3(n - fix)] and x2_out [2Np] may be NULL. */
EMBA_API int emba_solve(emba_handle_t h, double lambda, int32_t use_cg, int32_t fix_first_pose, double* x1_out,
               double* x2_out, int32_t* cg_iters, double* cg_error);

/* ---- Model::updateTraj (model.cpp:22-53) + LEGM::updateMap (:863-903): candidate <- current (+) last solution */
EMBA_API int emba_make_candidate(emba_handle_t h, double damping_factor, int32_t fix_first_pose);
/* solver.cpp:299-317: candidate becomes current (state, residuals, num_ev_map) */
EMBA_API int emba_accept_candidate(emba_handle_t h);

/* ---- EMBA::solveTimeWindow (solver.cpp:11-368): the whole LM loop with all state resident on the device.
 * Starts from the CURRENT state; on return the CURRENT state is the refined one. log may be NULL. */
EMBA_API int emba_solve_time_window(emba_handle_t h, const emba_lm_settings_t* s, emba_lm_log_t* log, int32_t log_cap,
                           int32_t* n_log, double* final_cost);
/* the same with the per-iteration hook (cb may be NULL) */
EMBA_API int emba_solve_time_window_cb(emba_handle_t h, const emba_lm_settings_t* s, emba_lm_log_t* log,
                                       int32_t log_cap, int32_t* n_log, double* final_cost, emba_lm_callback_t cb,
                                       void* user);

/* ---- "next" row N2: control-pose initialisation from a dense front-end trajectory,
 * LinearTrajectory::generateCtrlPosesLong (src/utils/trajectory.cpp:258-294; generateCtrlPoses :245-256,
 * fitCtrlPoses :149-229) as EMBA::Run calls it (src/emba/emba.cpp:416, sub-interval = dt_knots). poses: time-sorted
 * (t_ns, quaternion xyzw). t_beg_ns / t_end_ns are the ns-exact ros::Time::toNSec() of the interval the reference
 * passes (a double second at epoch scale resolves only ~240 ns). Writes floor((t_end - t_beg)/dt_knots + 1e-6) + 1
 * control poses (capacity cap).
 * EMBA_E_SUPPORT if a knot interval holds fewer than 2 front-end poses (the reference aborts there). No handle. */
EMBA_API int emba_fit_control_poses(int32_t device, int64_t n_poses, const int64_t* t_ns, const double* quat_xyzw,
                                    int64_t t_beg_ns, int64_t t_end_ns, double dt_knots, double* ctrl_quat_xyzw_out,
                                    int32_t cap, int32_t* n_ctrl_out);

/* ---- "next" row N3: intensity map from the gradient map,
 * poisson_reconstruction::reconstructFromGradient (src/image_rec/poisson_reconstruction.cpp:9-50; pde::poisolve with
 * zero Dirichlet boundary, src/image_rec/laplace.cpp:587-797), which the solver applies to cv::merge({Gx, Gy})
 * (src/emba/solver.cpp:412-417, :466-471). A plan owns the sine-transform matrices of one panorama size (the
 * counterpart of the FFTW plans the reference makes per call). Width and height must be even (EMBA_E_SUPPORT).
 * Gx, Gy, img_out: pano_h x pano_w row-major fp64 host buffers. */
typedef struct emba_poisson_s* emba_poisson_t;
EMBA_API int emba_poisson_create(int32_t device, int32_t pano_w, int32_t pano_h, emba_poisson_t* out);
EMBA_API int emba_poisson_destroy(emba_poisson_t p);
EMBA_API int emba_poisson_reconstruct(emba_poisson_t p, const double* Gx, const double* Gy, double* img_out);
/* device time of the last reconstruction (ms) and kernels launched by the plan so far; either may be NULL */
EMBA_API int emba_poisson_last_ms(emba_poisson_t p, double* ms_out, int64_t* launches_out);
/* the same reconstruction of the optimiser's own device-resident map (which: 0 = CURRENT, 1 = CANDIDATE): only the
 * image crosses PCIe */
EMBA_API int emba_reconstruct_map(emba_handle_t h, int32_t which, double* img_out);

/* ---- "next" row N4: EXTENSION MODE, the variant BASELINE.json's north star words literally --
 * cubic cumulative SO(3) B-spline evaluated PER EVENT with its 4 active control poses (re-derived from
 * basalt::So3Spline<4>::evaluate, thirdparty/basalt-headers/include/basalt/spline/so3_spline.h:218-274, which the
 * reference wraps in CubicTrajectory::evaluate, src/utils/trajectory.cpp:453-479), bilinear sampling of the gradient
 * map (8 map Jacobian entries instead of the nearest pixel's 2, model.cpp:209-214), 2 x 12 rotation Jacobian
 * entries per event pair (parity mode: 2 x 6, model.cpp:449,459). The reference has no implementation of this mode:
 * it is checked against oracle/ext_capi.cpp (the same model on the reference's vendored basalt) and reported
 * separately from the parity mode. The pairing of events comes from `h` (emba_ext_create runs emba_set_events_dev on
 * it), timestamps from the device-resident sequence, which must outlive the extension handle.
 *   emba_ext_evaluate : residuals, cost, the explicit sparse Jacobian rows, footprint counts
 *   emba_ext_form     : active pixels (footprint hits >= thres; a row is used iff its 4 pixels are active),
 *                       g = J^T e (- alpha G), block diagonal of H = J^T J (+ alpha I on the map block)
 *   emba_ext_solve    : block-Jacobi PCG on (H + lambda diag H) x = g, matrix free through the stored rows
 *   emba_ext_apply    : R_i <- Exp(x1_i) R_i, G[active] += damping * x2, G[inactive] = 0
 * Unknown order: 3 n_poses rotation components, then (Gx, Gy) per ACTIVE pixel in ascending pixel index. */
typedef struct emba_ext_s* emba_ext_t;
EMBA_API int emba_ext_create(emba_handle_t h, emba_events_t ev, int64_t idx_beg, int64_t idx_end, emba_ext_t* out);
EMBA_API int emba_ext_destroy(emba_ext_t x);
EMBA_API int emba_ext_num_pairs(emba_ext_t x, int64_t* out);
EMBA_API int emba_ext_set_state(emba_ext_t x, int64_t t0_ns, int64_t dt_ns, int32_t n_poses, const double* quat_xyzw,
                                const double* Gx, const double* Gy);
EMBA_API int emba_ext_get_state(emba_ext_t x, double* quat_xyzw_out, double* Gx_out, double* Gy_out);
EMBA_API int emba_ext_evaluate(emba_ext_t x, double alpha, double* cost_data, double* cost_reg, int64_t* num_measurements);
EMBA_API int emba_ext_form(emba_ext_t x, int32_t thres_valid_pixel, double alpha, int64_t* num_active_pixels,
                           int64_t* num_rows_used);
/* parity downloads: one row per event pair in time order of the current event (cap rows each, any may be NULL):
 * e, dp[2], pm[2] (warped current event), J[24] = Jc(12) | Jp(12), axy[2] (bilinear fractions), cp[2] (first control
 * pose of the current / previous event), pix[4] (footprint), ev (current event), flag (0 outlier, 1 inlier, 3 used) */
EMBA_API int emba_ext_get_rows(emba_ext_t x, int64_t cap, double* e, double* dp, double* pm, double* J24, double* axy,
                               int32_t* cp, int32_t* pix, int32_t* ev, int32_t* flag, int64_t* n_rows);
/* g [3n + 2Np], pose_blocks [n][9], pixel_blocks [Np][3] (xx, xy, yy), active [Np] */
EMBA_API int emba_ext_get_normal_eq(emba_ext_t x, double* g, double* pose_blocks, double* pixel_blocks, int32_t* active);
/* y = (J^T J + alpha I_map + lambda diag H) v, the operator emba_ext_solve iterates with (v, y: 3n + 2Np, host) */
EMBA_API int emba_ext_matvec(emba_ext_t x, double lambda, double alpha, const double* v, double* y);
EMBA_API int emba_ext_solve(emba_ext_t x, double lambda, double alpha, int32_t max_iter, double tol, double* x_out,
                            int32_t* iters, double* error);
EMBA_API int emba_ext_apply(emba_ext_t x, double damping_factor);
/* device ms of the last evaluate / form / solve */
EMBA_API int emba_ext_last_ms(emba_ext_t x, double* out3);

/* ---- measurement hooks used by bench.py (CUDA events on the handle's stream) */
/* elapsed device time (CUDA events on the handle's stream) of the last emba_evaluate / emba_form_normal_eq /
 * emba_solve, ms: out[0]=evaluate total, out[1]=per-measurement residual kernel (k_eval), out[2]=form total,
 * out[3]=pose-block assembly kernel (k_asm_pose), out[4]=everything after it (segments + k_pix + multi-GPU exchange),
 * out[5]=solve total, out[6]=map-block assembly kernel alone (k_pix), out[7]=row sort on its side stream (it runs
 * beside k_asm_pose, so its elapsed time includes the contention and is not additive) */
EMBA_API int emba_last_timings_ms(emba_handle_t h, double* out8);
/* device time (ms) of the last window set-up: out[0] = emba_set_events / emba_set_events_dev (upload + pair links),
 * out[1] = the static rebuild the first emba_set_state of a window triggers (canonical order, work items) */
EMBA_API int emba_last_setup_ms(emba_handle_t h, double* out2);
/* sizes of the last assembly on this handle (this rank): out[0] = measurements (pairs) of this rank's time slice,
 * out[1] = pairs of the whole window, out[2] = active pixels, out[3] = this rank's rows on active pixels,
 * out[4] = A12 entries this rank's own rows produced (local strips), out[5] = A12 entries it holds for the solve
 * (== out[4] with one GPU; with several, the merged strips of the pixels it owns), out[6] = work items,
 * out[7] = pixel segments longer than the warp sort's 1024 rows */
EMBA_API int emba_get_counters(emba_handle_t h, int64_t* out8);
/* several GPUs: device ms of the collective phases of the last pass on this rank. out[0] = histogram / cost
 * all-reduce of emba_evaluate, out[1] = window all-gather + exchange bookkeeping (before the map-side kernel),
 * out[2] = map-side kernel with the overlapped strip sends (until the last chunk has arrived), out[3] = A22 / b2
 * all-reduce + merge of the received sub-strips; out[7] = 1 if the A12 sub-strips of that pass went through peer memory
 * (the map-side kernel stores them straight into the owner rank's receive buffer over NVLink, mapped with CUDA IPC;
 * the default wherever every rank can map every other rank's buffer), 0 if through the ncclSend/ncclRecv all-to-all
 * (EMBA_XCHG_PEER=0, or the mapping failed on some rank); the rest 0 */
EMBA_API int emba_last_comm_ms(emba_handle_t h, double* out8);
/* number of kernel launches issued by this handle so far */
EMBA_API int emba_launch_count(emba_handle_t h, int64_t* out);
EMBA_API int emba_synchronize(emba_handle_t h);

/* library / build identification ("emba_b200 <version> sm_100a") */
EMBA_API const char* emba_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EMBA_B200_H_ */
