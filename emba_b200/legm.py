"""Host-side mirror of the reference's optimiser interface for the LM hot path.

`Engine` is a thin numpy-in / numpy-out wrapper of the C ABI (include/emba_b200.h).
`LEGM` re-creates the method names, argument meaning and call order of the reference's
`EMBA::LEGM` (include/emba/model.h:72-133) and `Trajectory` the part of
`LinearTrajectory` (include/utils/trajectory.h:106-191) that the LM loop touches, so that
`solveTimeWindow` below reads like the reference's `EMBA::solveTimeWindow`
(src/emba/solver.cpp:11-368) and the parity tests read like tests of the reference.

All numerics run in the CUDA library; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import EmbaError, LMLog, LMSettings, ptr


def spline_base_ns(t_beg: float, dt_knots: float):
    """int64_t(1e9*t_beg), int64_t(1e9*dt_knots): the truncating expression of
    LinearTrajectory(double, double, cps) (src/utils/trajectory.cpp:61-70)."""
    return int(np.int64(np.float64(1e9) * np.float64(t_beg))), int(np.int64(np.float64(1e9) * np.float64(dt_knots)))


def ros_time_ns(t_sec: float) -> int:
    """ros::Time(double).toNSec(): sec = floor(t), nsec = round((t - sec) * 1e9) (rostime)."""
    s = np.floor(np.float64(t_sec))
    return int(s) * 1_000_000_000 + int(np.round((np.float64(t_sec) - s) * 1e9))


def fit_control_poses(t_ns, quat_xyzw, t_beg, t_end, dt_knots, device=0):
    """LinearTrajectory::generateCtrlPosesLong (src/utils/trajectory.cpp:258-294) on the GPU: initial control poses
    from a dense front-end trajectory (time-sorted t_ns, xyzw quaternions). t_beg / t_end: seconds (converted like
    ros::Time(double)) or integer nanoseconds."""
    L = capi.load()
    t = np.ascontiguousarray(t_ns, dtype=np.int64)
    q = np.ascontiguousarray(quat_xyzw, dtype=np.float64)
    n = C.c_int32(0)
    tb = int(t_beg) if isinstance(t_beg, (int, np.integer)) else ros_time_ns(t_beg)
    te = int(t_end) if isinstance(t_end, (int, np.integer)) else ros_time_ns(t_end)
    cap = int(round((te - tb) * 1e-9 / dt_knots)) + 8
    out = np.empty((cap, 4))
    rc = L.emba_fit_control_poses(device, t.size, ptr(t, C.c_int64), ptr(q), tb, te,
                                  float(dt_knots), ptr(out), cap, C.byref(n))
    if rc != 0:
        raise EmbaError(rc, "emba_fit_control_poses")
    return out[: n.value].copy()


class PoissonPlan:
    """poisson_reconstruction::reconstructFromGradient (src/image_rec/poisson_reconstruction.cpp:9-50) on the GPU for
    one panorama size; `reconstruct(Gx, Gy)` returns the intensity map (pano_h x pano_w)."""

    def __init__(self, pano_w, pano_h, device=0):
        self.L = capi.load()
        self.pano_w, self.pano_h = int(pano_w), int(pano_h)
        self.h = C.c_void_p()
        rc = self.L.emba_poisson_create(int(device), self.pano_w, self.pano_h, C.byref(self.h))
        if rc != 0:
            raise EmbaError(rc, "emba_poisson_create")

    def reconstruct(self, Gx, Gy):
        gx = np.ascontiguousarray(Gx, dtype=np.float64)
        gy = np.ascontiguousarray(Gy, dtype=np.float64)
        assert gx.shape == (self.pano_h, self.pano_w) and gy.shape == gx.shape
        out = np.empty_like(gx)
        rc = self.L.emba_poisson_reconstruct(self.h, ptr(gx), ptr(gy), ptr(out))
        if rc != 0:
            raise EmbaError(rc, "emba_poisson_reconstruct")
        return out

    def last_ms(self):
        ms = C.c_double(0.0)
        n = C.c_int64(0)
        self.L.emba_poisson_last_ms(self.h, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def close(self):
        if self.h:
            self.L.emba_poisson_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def reconstructFromGradient(gradients, device=0):
    """Reference-shaped entry point: `gradients` is the H x W x 2 array cv::merge({Gx, Gy}) produces
    (src/emba/solver.cpp:412-417); returns the H x W intensity map."""
    g = np.asarray(gradients, dtype=np.float64)
    assert g.ndim == 3 and g.shape[2] == 2
    plan = PoissonPlan(g.shape[1], g.shape[0], device)
    try:
        return plan.reconstruct(g[:, :, 0], g[:, :, 1])
    finally:
        plan.close()


class EventSequence:
    """Device-resident event recording (include/emba_b200.h: emba_events_*): the steps the reference runs on the host
    before every window -- time sort (src/utils/rosbag_loading.cpp:61-65), subsampling (src/emba/emba.cpp:281-304),
    robust window extraction EMBA::getEventSubset (src/emba/emba.cpp:473-510) -- on the GPU."""

    def __init__(self, x, y, t_ns, pol, device=0):
        self.L = capi.load()
        x = np.ascontiguousarray(x, dtype=np.uint16)
        y = np.ascontiguousarray(y, dtype=np.uint16)
        t = np.ascontiguousarray(t_ns, dtype=np.int64)
        p = np.ascontiguousarray(pol, dtype=np.uint8)
        self.h = C.c_void_p()
        rc = self.L.emba_events_create(int(device), x.size, ptr(x, C.c_uint16), ptr(y, C.c_uint16), ptr(t, C.c_int64),
                                       ptr(p, C.c_uint8), C.byref(self.h))
        if rc != 0:
            raise EmbaError(rc, "emba_events_create")

    def _chk(self, rc, what):
        if rc != 0:
            raise EmbaError(rc, what)

    def count(self):
        v = C.c_int64(0)
        self._chk(self.L.emba_events_count(self.h, C.byref(v)), "emba_events_count")
        return v.value

    def sort_by_time(self):
        self._chk(self.L.emba_events_sort_by_time(self.h), "emba_events_sort_by_time")

    def subsample(self, event_sampling_rate):
        self._chk(self.L.emba_events_subsample(self.h, int(event_sampling_rate)), "emba_events_subsample")

    def window(self, t_beg_ns, t_end_ns):
        """EMBA::getEventSubset: returns (idx_beg, idx_end) of the event subset."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._chk(self.L.emba_events_window(self.h, int(t_beg_ns), int(t_end_ns), C.byref(a), C.byref(b)),
                  "emba_events_window")
        return a.value, b.value

    def download(self, i0=0, i1=None):
        i1 = self.count() if i1 is None else i1
        n = max(0, i1 - i0)
        x, y = np.empty(n, np.uint16), np.empty(n, np.uint16)
        t, p = np.empty(n, np.int64), np.empty(n, np.uint8)
        self._chk(self.L.emba_events_download(self.h, i0, i1, ptr(x, C.c_uint16), ptr(y, C.c_uint16),
                                              ptr(t, C.c_int64), ptr(p, C.c_uint8)), "emba_events_download")
        return x, y, t, p

    def close(self):
        if getattr(self, "h", None):
            self.L.emba_events_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    def __init__(self, sensor_w, sensor_h, bearing_lut, C_th, pano_w, pano_h, device=0):
        self.L = capi.load()
        self.sensor_w, self.sensor_h, self.pano_w, self.pano_h = sensor_w, sensor_h, pano_w, pano_h
        self.P = pano_w * pano_h
        lut = np.ascontiguousarray(bearing_lut, dtype=np.float64).reshape(-1)
        assert lut.size == 3 * sensor_w * sensor_h
        cfg = capi.Config(sensor_w, sensor_h, pano_w, pano_h, float(C_th), ptr(lut), device)
        self.h = C.c_void_p()
        rc = self.L.emba_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise EmbaError(rc, "emba_create failed (no CUDA device? there is no CPU fallback)")
        self.n = 0
        self.N = 0

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.emba_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise EmbaError(rc, self.L.emba_last_error(self.h).decode())

    # -- events / shard -------------------------------------------------------------------------
    def set_events(self, x, y, t_ns, pol):
        x = np.ascontiguousarray(x, dtype=np.uint16)
        y = np.ascontiguousarray(y, dtype=np.uint16)
        t = np.ascontiguousarray(t_ns, dtype=np.int64)
        p = np.ascontiguousarray(pol, dtype=np.uint8)
        self.N = x.size
        self._chk(self.L.emba_set_events(self.h, x.size, ptr(x, C.c_uint16), ptr(y, C.c_uint16), ptr(t, C.c_int64),
                                         ptr(p, C.c_uint8)))

    def set_events_dev(self, seq: "EventSequence", idx_beg, idx_end):
        """per-window pre-pass straight from a device-resident sequence (no host copy)"""
        self.N = int(idx_end - idx_beg)
        self._chk(self.L.emba_set_events_dev(self.h, seq.h, int(idx_beg), int(idx_end)))

    def setup_ms(self):
        """device ms of the last (set_events, static rebuild)"""
        out = np.zeros(2)
        self._chk(self.L.emba_last_setup_ms(self.h, ptr(out)))
        return float(out[0]), float(out[1])

    def counters(self):
        out = np.zeros(8, dtype=np.int64)
        self._chk(self.L.emba_get_counters(self.h, ptr(out, C.c_int64)))
        return dict(measurements_rank=int(out[0]), pairs_window=int(out[1]), active_pixels=int(out[2]),
                    rows_on_active_rank=int(out[3]), a12_entries_local=int(out[4]), a12_entries_solve=int(out[5]),
                    work_items=int(out[6]), long_segments=int(out[7]))

    def comm_ms(self):
        out = np.zeros(8)
        self._chk(self.L.emba_last_comm_ms(self.h, ptr(out)))
        return dict(eval_allreduce=out[0], exchange_prepare=out[1], pix_and_sends=out[2], allreduce_and_merge=out[3],
                    strip_exchange="peer" if out[7] == 1.0 else "nccl")

    def set_strict_range(self, on):
        self._chk(self.L.emba_set_strict_range(self.h, int(bool(on))))

    def get_jacobian_rows(self):
        """[M, 16] rows [Jc(6) Jp(6) e dp(2) 0] of the last assembly in the reference's measurement order"""
        cap = max(self.num_pairs(), 1)
        rows = np.empty((cap, 16))
        n = C.c_int64(0)
        self._chk(self.L.emba_get_jacobian_rows(self.h, ptr(rows), cap, C.byref(n)))
        return rows[: n.value].copy()

    def num_pairs(self):
        v = C.c_int64(0)
        self._chk(self.L.emba_num_pairs(self.h, C.byref(v)))
        return v.value

    def set_shard(self, rank, world):
        self._chk(self.L.emba_set_shard(self.h, rank, world))

    def comm_init(self, uid_bytes, rank, world):
        buf = C.create_string_buffer(bytes(uid_bytes), 128)
        self._chk(self.L.emba_comm_init(self.h, buf, rank, world))

    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        rc = self.L.emba_comm_unique_id(buf)
        if rc != 0:
            raise EmbaError(rc, "emba_comm_unique_id")
        return buf.raw

    # -- state -----------------------------------------------------------------------------------
    def set_state(self, which, t0_ns, dt_ns, quat, Gx, Gy):
        q = np.ascontiguousarray(quat, dtype=np.float64)
        gx = np.ascontiguousarray(Gx, dtype=np.float64)
        gy = np.ascontiguousarray(Gy, dtype=np.float64)
        assert gx.size == self.P and gy.size == self.P
        self.n = q.shape[0]
        self._chk(self.L.emba_set_state(self.h, which, int(t0_ns), int(dt_ns), self.n, ptr(q), ptr(gx), ptr(gy)))

    def get_state(self, which=capi.STATE_CURRENT):
        q = np.empty((self.n, 4))
        gx = np.empty((self.pano_h, self.pano_w))
        gy = np.empty((self.pano_h, self.pano_w))
        self._chk(self.L.emba_get_state(self.h, which, ptr(q), ptr(gx), ptr(gy)))
        return q, gx, gy

    def reconstruct_map(self, which=capi.STATE_CURRENT):
        """Poisson intensity map of the device-resident gradient maps (solver.cpp:412-417)."""
        img = np.empty((self.pano_h, self.pano_w))
        self._chk(self.L.emba_reconstruct_map(self.h, which, ptr(img)))
        return img

    # -- evaluate ----------------------------------------------------------------------------------
    def evaluate(self, which=capi.STATE_CURRENT, cost_type=0, eta=1.0, alpha=0.0):
        cd, cr, M = C.c_double(0), C.c_double(0), C.c_int64(0)
        self._chk(self.L.emba_evaluate(self.h, which, cost_type, float(eta), float(alpha), C.byref(cd), C.byref(cr),
                                       C.byref(M)))
        return cd.value, cr.value, M.value

    def get_evaluation(self, which=capi.STATE_CURRENT, M=None, want_ep=True, want_num=True):
        ep = np.empty(max(self.num_pairs(), 1)) if want_ep else None
        num = np.empty((self.pano_h, self.pano_w), dtype=np.int32) if want_num else None
        self._chk(self.L.emba_get_evaluation(self.h, which, ptr(ep), ptr(num, C.c_int32)))
        if ep is not None and M is not None:
            ep = ep[:M]
        return ep, num

    # -- normal equations ----------------------------------------------------------------------------
    def form_normal_eq(self, thres, cost_type=0, eta=1.0, alpha=0.0):
        Np = C.c_int64(0)
        self._chk(self.L.emba_form_normal_eq(self.h, thres, cost_type, float(eta), float(alpha), C.byref(Np)))
        self.Np = Np.value
        return self.Np

    def set_map_path(self, mode):
        """0 = deterministic sorted segmented reduction (default), 1 = fp64-atomic path."""
        self._chk(self.L.emba_set_map_path(self.h, int(mode)))

    def apply_l2_reg(self, alpha):
        self._chk(self.L.emba_apply_l2_reg(self.h, float(alpha)))

    def get_normal_eq(self, want_A12=True):
        n, Np = self.n, self.Np
        A11 = np.empty((3 * n, 3 * n))
        b1 = np.empty(3 * n)
        A22 = np.empty((Np, 2, 2))
        b2 = np.empty(2 * Np)
        act = np.empty(Np, dtype=np.int64)
        A12 = np.empty((3 * n, 2 * Np)) if want_A12 else None
        self._chk(self.L.emba_get_normal_eq(self.h, ptr(A11), ptr(b1), ptr(A22), ptr(b2), ptr(act, C.c_int64),
                                            ptr(A12)))
        return A11, A12, A22, b1, b2, act

    def a12_entries(self):
        v = C.c_int64(0)
        self._chk(self.L.emba_a12_entries(self.h, C.byref(v)))
        return v.value

    def solve(self, lam, use_cg=False, fix_first=True, want=True):
        d = 3 * (self.n - (1 if fix_first else 0))
        x1 = np.empty(d) if want else None
        x2 = np.empty(2 * self.Np) if want else None
        it, err = C.c_int32(0), C.c_double(0)
        self._chk(self.L.emba_solve(self.h, float(lam), int(use_cg), int(fix_first), ptr(x1), ptr(x2), C.byref(it),
                                    C.byref(err)))
        return x1, x2, it.value, err.value

    def make_candidate(self, damping, fix_first=True):
        self._chk(self.L.emba_make_candidate(self.h, float(damping), int(fix_first)))

    def accept_candidate(self):
        self._chk(self.L.emba_accept_candidate(self.h))

    def solve_time_window(self, *, max_num_iter=50, tol_fun=1e-3, num_times_tol_fun_sat=2, use_cg=False, cost_type=0,
                          eta=1.0, thres=5, damping=1.0, alpha=5.0, first_window=True, callback=None):
        """EMBA::solveTimeWindow on the device. Returns (log rows, final cost); row columns: iter, lambda, cost_min,
        cost_new, accepted, active pixels, measurements, cg_iters, cg_error, ms_form, ms_solve, ms_evaluate.
        callback(row_dict) is called after every solve (solver.cpp:170-179); a truthy return stops the loop."""
        s = LMSettings(max_num_iter, tol_fun, num_times_tol_fun_sat, int(use_cg), cost_type, eta, thres, damping,
                       alpha, int(first_window))
        cap = max_num_iter + 8
        log = (LMLog * cap)()
        nlog, fc = C.c_int32(0), C.c_double(0)

        def _cb(rowp, _user):
            r = rowp.contents
            return 1 if callback({"iter": r.iter, "lambda": r.lambda_, "cost_min": r.cost_min, "cost_new": r.cost_new,
                                  "accepted": r.accepted, "cg_iters": r.cg_iters, "cg_error": r.cg_error}) else 0

        cfn = capi.LM_CALLBACK(_cb) if callback is not None else C.cast(None, capi.LM_CALLBACK)
        self._chk(self.L.emba_solve_time_window_cb(self.h, C.byref(s), log, cap, C.byref(nlog), C.byref(fc), cfn, None))
        rows = np.array([[r.iter, r.lambda_, r.cost_min, r.cost_new, r.accepted, r.num_active_pixels,
                          r.num_measurements, r.cg_iters, r.cg_error, r.ms_form, r.ms_solve, r.ms_evaluate]
                         for r in log[: nlog.value]], dtype=np.float64).reshape(-1, 12)
        return rows, fc.value

    def timings_ms(self):
        out = np.zeros(8)
        self._chk(self.L.emba_last_timings_ms(self.h, ptr(out)))
        return dict(evaluate=out[0], eval_kernel=out[1], form=out[2], asm_pose_kernel=out[3], map_side=out[4],
                    solve=out[5], pix_kernel=out[6], sort=out[7])

    def launch_count(self):
        v = C.c_int64(0)
        self._chk(self.L.emba_launch_count(self.h, C.byref(v)))
        return v.value

    def synchronize(self):
        self._chk(self.L.emba_synchronize(self.h))


class ExtEngine:
    """Extension mode (include/emba_b200.h: emba_ext_*): cubic per-event SO(3) spline, bilinear map sampling,
    normal equations in Jacobian form solved by block-Jacobi PCG. `engine` provides the pairing, `seq` the timestamps."""

    def __init__(self, engine: Engine, seq: EventSequence, idx_beg, idx_end):
        self.L, self.eng, self.seq = engine.L, engine, seq
        self.h = C.c_void_p()
        engine._chk(self.L.emba_ext_create(engine.h, seq.h, int(idx_beg), int(idx_end), C.byref(self.h)))
        self.n, self.Np = 0, 0

    def _chk(self, rc):
        self.eng._chk(rc)

    def close(self):
        if getattr(self, "h", None):
            self.L.emba_ext_destroy(self.h)
            self.h = None

    def num_pairs(self):
        v = C.c_int64(0)
        self._chk(self.L.emba_ext_num_pairs(self.h, C.byref(v)))
        return v.value

    def set_state(self, t0_ns, dt_ns, quat, Gx, Gy):
        q = np.ascontiguousarray(quat, dtype=np.float64)
        gx = np.ascontiguousarray(Gx, dtype=np.float64)
        gy = np.ascontiguousarray(Gy, dtype=np.float64)
        self.n = q.shape[0]
        self._chk(self.L.emba_ext_set_state(self.h, int(t0_ns), int(dt_ns), self.n, ptr(q), ptr(gx), ptr(gy)))

    def get_state(self):
        q = np.empty((self.n, 4))
        gx = np.empty((self.eng.pano_h, self.eng.pano_w))
        gy = np.empty((self.eng.pano_h, self.eng.pano_w))
        self._chk(self.L.emba_ext_get_state(self.h, ptr(q), ptr(gx), ptr(gy)))
        return q, gx, gy

    def evaluate(self, alpha=0.0):
        cd, cr, M = C.c_double(0), C.c_double(0), C.c_int64(0)
        self._chk(self.L.emba_ext_evaluate(self.h, float(alpha), C.byref(cd), C.byref(cr), C.byref(M)))
        return cd.value, cr.value, M.value

    def form(self, thres, alpha):
        Np, Mu = C.c_int64(0), C.c_int64(0)
        self._chk(self.L.emba_ext_form(self.h, int(thres), float(alpha), C.byref(Np), C.byref(Mu)))
        self.Np = Np.value
        return Np.value, Mu.value

    def get_rows(self):
        M = self.num_pairs()
        cap = max(M, 1)
        out = dict(e=np.empty(cap), dp=np.empty((cap, 2)), pm=np.empty((cap, 2)), J=np.empty((cap, 24)),
                   axy=np.empty((cap, 2)), cp=np.empty((cap, 2), np.int32), pix=np.empty((cap, 4), np.int32),
                   ev=np.empty(cap, np.int32), flag=np.empty(cap, np.int32))
        n = C.c_int64(0)
        self._chk(self.L.emba_ext_get_rows(self.h, cap, ptr(out["e"]), ptr(out["dp"]), ptr(out["pm"]), ptr(out["J"]),
                                           ptr(out["axy"]), ptr(out["cp"], C.c_int32), ptr(out["pix"], C.c_int32),
                                           ptr(out["ev"], C.c_int32), ptr(out["flag"], C.c_int32), C.byref(n)))
        return {k: v[: n.value] for k, v in out.items()}

    def get_normal_eq(self):
        d = 3 * self.n + 2 * self.Np
        g, Bp, Bm, act = np.empty(d), np.empty((self.n, 3, 3)), np.empty((max(self.Np, 1), 3)), np.empty(max(self.Np, 1), np.int32)
        self._chk(self.L.emba_ext_get_normal_eq(self.h, ptr(g), ptr(Bp), ptr(Bm), ptr(act, C.c_int32)))
        return g, Bp, Bm[: self.Np], act[: self.Np]

    def matvec(self, lam, alpha, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        y = np.empty_like(v)
        self._chk(self.L.emba_ext_matvec(self.h, float(lam), float(alpha), ptr(v), ptr(y)))
        return y

    def solve(self, lam, alpha, max_iter=100, tol=1e-6):
        x = np.empty(3 * self.n + 2 * self.Np)
        it, err = C.c_int32(0), C.c_double(0)
        self._chk(self.L.emba_ext_solve(self.h, float(lam), float(alpha), int(max_iter), float(tol), ptr(x), C.byref(it),
                                        C.byref(err)))
        return x, it.value, err.value

    def apply(self, damping=1.0):
        self._chk(self.L.emba_ext_apply(self.h, float(damping)))

    def last_ms(self):
        out = np.zeros(3)
        self._chk(self.L.emba_ext_last_ms(self.h, ptr(out)))
        return dict(evaluate=out[0], form=out[1], solve=out[2])


# ------------------------------------------------------------------------------------------------
# reference-shaped interface
# ------------------------------------------------------------------------------------------------
class Trajectory:
    """The part of the reference's LinearTrajectory the LM loop uses (include/utils/trajectory.h:106-191):
    size(), getControlPose(), begTime(), getDtCtrlPoses(), clone(). Control poses are xyzw unit quaternions."""

    def __init__(self, t_beg, dt_knots, quat_xyzw):
        self.t_beg_ = float(t_beg)
        self.dt_knots_ = float(dt_knots)
        self.quat = np.array(quat_xyzw, dtype=np.float64, copy=True)

    def size(self):
        return self.quat.shape[0]

    def getControlPose(self, i):
        return self.quat[i]

    def begTime(self):
        return self.t_beg_

    def getDtCtrlPoses(self):
        return self.dt_knots_

    def clone(self):
        return Trajectory(self.t_beg_, self.dt_knots_, self.quat)


class EventPacket:
    """std::vector<dvs_msgs::Event> as flat arrays (include/emba/model.h:15)."""

    def __init__(self, x, y, t_ns, polarity):
        self.x, self.y, self.t_ns, self.polarity = x, y, t_ns, polarity

    def size(self):
        return int(np.asarray(self.t_ns).size)


class LEGM:
    """Drop-in shaped like EMBA::LEGM (include/emba/model.h:72-133), running on the GPU.

    Like the reference, the model keeps hidden state between calls: the per-measurement data of the LAST
    evaluated point (the reference's event_map_, model.h:131-132). evaluateDataError() evaluates into the device's
    candidate slot; formNormalEq() promotes that slot to current first, which is exactly the reference's rule that
    the normal equations are formed from the last evaluated point (solver.cpp:96-130)."""

    def __init__(self, sensor_w, sensor_h, bearing_lut, C_th, pano_width, pano_height, device=0):
        self.eng = Engine(sensor_w, sensor_h, bearing_lut, C_th, pano_width, pano_height, device)
        self._events_id = None
        self._pending = False
        self._M = 0
        self._fix = 1

    def _sync_events(self, events: EventPacket):
        key = (id(events.t_ns), events.size())
        if key != self._events_id:
            self.eng.set_events(events.x, events.y, events.t_ns, events.polarity)
            self._events_id = key

    def evaluateDataError(self, traj: Trajectory, Gx, Gy, events: EventPacket, eval_deriv=True, num_ev_map=None,
                          cost_type=0, eta=1.0):
        """model.cpp:72-258. Returns ep (reference order); fills num_ev_map in place if given."""
        self._sync_events(events)
        t0, dt = spline_base_ns(traj.begTime(), traj.getDtCtrlPoses())
        self.eng.set_state(capi.STATE_CANDIDATE, t0, dt, traj.quat, Gx, Gy)
        cd, _, M = self.eng.evaluate(capi.STATE_CANDIDATE, cost_type, eta, 0.0)
        self._pending, self._M, self._cost_data = True, M, cd
        ep, num = self.eng.get_evaluation(capi.STATE_CANDIDATE, M, True, num_ev_map is not None)
        if num_ev_map is not None:
            num_ev_map[...] = num
        return ep

    def evaluateRegError(self, Gx, Gy):
        """model.cpp:260-277 (a plain reshuffle of the map values; host side like the reference)."""
        return np.stack([np.asarray(Gx).reshape(-1), np.asarray(Gy).reshape(-1)], -1).reshape(-1)

    def lastDataCost(self):
        """0.5*ep.ep (or the robust cost) of the last evaluateDataError, reduced on the device."""
        return self._cost_data

    def formNormalEq(self, num_ctrl_poses, thres_valid_pixel, cost_type=0, eta=1.0, want_A12=True):
        """model.cpp:316-491 (IRLS: 493-687). Returns A11, A12, A22_blocks, b1, b2, active_pix_idxes."""
        if self._pending:
            self.eng.accept_candidate()
            self._pending = False
        assert num_ctrl_poses == self.eng.n
        self.eng.form_normal_eq(thres_valid_pixel, cost_type, eta, 0.0)
        return self.eng.get_normal_eq(want_A12)

    def applyL2Reg(self, alpha):
        """model.cpp:689-719 on the device-resident A22 / b2; returns the updated (A22_blocks, b2)."""
        self.eng.apply_l2_reg(alpha)
        _, _, A22, _, b2, _ = self.eng.get_normal_eq(False)
        return A22, b2

    def solveNormalEq(self, lam, fix_first=True):
        """model.cpp:721-792."""
        x1, x2, _, _ = self.eng.solve(lam, False, fix_first)
        self._fix = fix_first
        return x1, x2

    def solveNormalEqCG(self, lam, fix_first=True):
        """model.cpp:794-840. Returns x1, x2, (iterations, error)."""
        x1, x2, it, err = self.eng.solve(lam, True, fix_first)
        self._fix = fix_first
        return x1, x2, (it, err)

    def updateTrajAndMap(self, traj: Trajectory, damping_factor):
        """Model::updateTraj (model.cpp:22-53) + LEGM::updateMap (model.cpp:863-903) applied to clones of the
        current state with the last solution; returns (traj_new, Gx_new, Gy_new)."""
        self.eng.make_candidate(damping_factor, self._fix)
        q, gx, gy = self.eng.get_state(capi.STATE_CANDIDATE)
        return Trajectory(traj.begTime(), traj.getDtCtrlPoses(), q), gx, gy
