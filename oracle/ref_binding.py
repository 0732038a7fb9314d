"""ctypes binding of oracle/_ref/libemba_ref.so (the unmodified reference hot-path
sources + oracle/ref_capi.cpp). TEST INFRASTRUCTURE ONLY: never imported by
emba_b200/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libemba_ref.so")

_dp = C.POINTER(C.c_double)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _p(a, ty=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


class RefLib:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(LIB_PATH)
            L.embaref_create.restype = C.c_void_p
            L.embaref_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_double, C.c_int, C.c_int]
            L.embaref_destroy.argtypes = [C.c_void_p]
            L.embaref_get_bearing_lut.argtypes = [C.c_void_p, _dp]
            L.embaref_set_events.argtypes = [C.c_void_p, C.c_long, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16),
                                             C.POINTER(C.c_int64), C.POINTER(C.c_uint8)]
            L.embaref_traj_create.restype = C.c_void_p
            L.embaref_traj_create.argtypes = [C.c_double, C.c_double, C.c_int, _dp]
            L.embaref_traj_destroy.argtypes = [C.c_void_p]
            L.embaref_traj_size.argtypes = [C.c_void_p]
            L.embaref_traj_get.argtypes = [C.c_void_p, _dp]
            L.embaref_traj_evaluate.argtypes = [C.c_void_p, C.c_int64, _dp, _dp]
            L.embaref_update_traj.argtypes = [C.c_void_p, C.c_void_p, _dp, C.c_int, C.c_int]
            L.embaref_evaluate.restype = C.c_long
            L.embaref_evaluate.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, C.c_int, _dp, C.POINTER(C.c_int32)]
            L.embaref_dump_measurements.restype = C.c_long
            L.embaref_dump_measurements.argtypes = [C.c_void_p, _dp, C.c_long, C.POINTER(C.c_long)]
            L.embaref_reg_cost.restype = C.c_double
            L.embaref_reg_cost.argtypes = [C.c_void_p, _dp, _dp, C.c_double]
            L.embaref_data_cost.restype = C.c_double
            L.embaref_data_cost.argtypes = [C.c_void_p, C.c_int, C.c_double]
            L.embaref_form.restype = C.c_long
            L.embaref_form.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _dp, _dp]
            L.embaref_get_normal_eq.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int64)]
            L.embaref_a12_nnz.restype = C.c_long
            L.embaref_a12_nnz.argtypes = [C.c_void_p]
            L.embaref_solve.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, _dp, _dp, C.POINTER(C.c_int), _dp]
            L.embaref_update_map.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double]
            L.embaref_solve_time_window.restype = C.c_int
            L.embaref_solve_time_window.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), _dp, _dp, C.c_int, C.c_double,
                                                    C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double,
                                                    C.c_double, C.c_int, _dp, C.c_int, _dp]
            L.embaref_get_timers.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_long)]
            L.embaref_generate_ctrl_poses_long.restype = C.c_int
            L.embaref_generate_ctrl_poses_long.argtypes = [C.c_long, C.POINTER(C.c_int64), _dp, C.c_double, C.c_double,
                                                           C.c_double, C.c_double, _dp, C.c_int]
            L.embaref_poisson_reconstruct.restype = C.c_int
            L.embaref_poisson_reconstruct.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp]
            cls._lib = L
        return cls._lib


def ref_poisson_reconstruct(Gx, Gy):
    """poisson_reconstruction::reconstructFromGradient of the reference (poisson_reconstruction.cpp:9-50)."""
    L = RefLib.lib()
    gx = np.ascontiguousarray(Gx, dtype=np.float64)
    gy = np.ascontiguousarray(Gy, dtype=np.float64)
    H, W = gx.shape
    out = np.empty((H, W))
    rc = L.embaref_poisson_reconstruct(_p(gx), _p(gy), H, W, _p(out))
    assert rc == 0
    return out


def ref_generate_ctrl_poses_long(t_ns, quat_xyzw, t_beg, t_end, dt_knots, sub_interval=None):
    """LinearTrajectory::generateCtrlPosesLong of the reference (trajectory.cpp:258-294)."""
    L = RefLib.lib()
    t = np.ascontiguousarray(t_ns, dtype=np.int64)
    q = np.ascontiguousarray(quat_xyzw, dtype=np.float64)
    cap = int(round((t_end - t_beg) / dt_knots)) + 8
    out = np.empty((cap, 4))
    n = L.embaref_generate_ctrl_poses_long(t.size, _p(t, C.c_int64), _p(q), float(t_beg), float(t_end),
                                           float(dt_knots), float(dt_knots if sub_interval is None else sub_interval),
                                           _p(out), cap)
    assert n >= 0
    return out[:n].copy()


class RefTraj:
    """LinearTrajectory built with the (t_beg, dt_knots, cps) constructor (trajectory.cpp:61-74)."""

    def __init__(self, t_beg, dt_knots, quat_xyzw):
        self.L = RefLib.lib()
        q = np.ascontiguousarray(quat_xyzw, dtype=np.float64)
        self.n = q.shape[0]
        self.h = C.c_void_p(self.L.embaref_traj_create(float(t_beg), float(dt_knots), self.n, _p(q)))

    def quat(self):
        out = np.empty((self.n, 4))
        self.L.embaref_traj_get(self.h, _p(out))
        return out

    def evaluate(self, t_ns):
        R = np.empty(9)
        J = np.empty(18)
        idx = self.L.embaref_traj_evaluate(self.h, int(t_ns), _p(R), _p(J))
        return R.reshape(3, 3), J.reshape(3, 6), idx

    def __del__(self):
        try:
            if self.h:
                self.L.embaref_traj_destroy(self.h)
                self.h = None
        except Exception:
            pass


class RefLEGM:
    """EMBA::LEGM of the reference (include/emba/model.h:72-133), zero-distortion pinhole camera."""

    def __init__(self, sensor_w, sensor_h, fx, fy, cx, cy, C_th, pano_w, pano_h):
        self.L = RefLib.lib()
        self.sw, self.sh, self.pw, self.ph = sensor_w, sensor_h, pano_w, pano_h
        self.h = C.c_void_p(self.L.embaref_create(sensor_w, sensor_h, fx, fy, cx, cy, C_th, pano_w, pano_h))
        self.N = 0
        self.n_poses = 0
        self.Np = 0

    def __del__(self):
        try:
            if self.h:
                self.L.embaref_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def bearing_lut(self):
        out = np.empty((self.sw * self.sh, 3))
        self.L.embaref_get_bearing_lut(self.h, _p(out))
        return out

    def set_events(self, x, y, t_ns, pol):
        x = np.ascontiguousarray(x, dtype=np.uint16)
        y = np.ascontiguousarray(y, dtype=np.uint16)
        t = np.ascontiguousarray(t_ns, dtype=np.int64)
        p = np.ascontiguousarray(pol, dtype=np.uint8)
        self.N = x.size
        self.L.embaref_set_events(self.h, self.N, _p(x, C.c_uint16), _p(y, C.c_uint16), _p(t, C.c_int64),
                                  _p(p, C.c_uint8))

    def evaluate(self, traj: RefTraj, Gx, Gy, eval_deriv=True):
        Gx = np.ascontiguousarray(Gx, dtype=np.float64)
        Gy = np.ascontiguousarray(Gy, dtype=np.float64)
        ep = np.empty(max(self.N, 1))
        num = np.empty((self.ph, self.pw), dtype=np.int32)
        M = self.L.embaref_evaluate(self.h, traj.h, _p(Gx), _p(Gy), int(eval_deriv), _p(ep), _p(num, C.c_int32))
        return ep[:M].copy(), num

    def dump_measurements(self):
        rec = np.empty((max(self.N, 1), 25))
        nout = C.c_long(0)
        M = self.L.embaref_dump_measurements(self.h, _p(rec), rec.shape[0], C.byref(nout))
        assert M >= 0
        return rec[:M].copy(), nout.value

    def reg_cost(self, Gx, Gy, alpha):
        Gx = np.ascontiguousarray(Gx, dtype=np.float64)
        Gy = np.ascontiguousarray(Gy, dtype=np.float64)
        return self.L.embaref_reg_cost(self.h, _p(Gx), _p(Gy), float(alpha))

    def data_cost(self, irls_type=0, a=1.0):
        return self.L.embaref_data_cost(self.h, irls_type, float(a))

    def form(self, n_poses, thres, Gx, Gy, alpha, irls_type=0, a=1.0, apply_l2=True, want_A12=True):
        Gx = np.ascontiguousarray(Gx, dtype=np.float64)
        Gy = np.ascontiguousarray(Gy, dtype=np.float64)
        self.n_poses = n_poses
        Np = self.L.embaref_form(self.h, n_poses, thres, irls_type, float(a), float(alpha), int(apply_l2), _p(Gx),
                                 _p(Gy))
        self.Np = Np
        d = 3 * n_poses
        A11 = np.empty((d, d))
        A12 = np.empty((d, 2 * Np)) if want_A12 else None
        A22 = np.empty((Np, 2, 2))
        b1 = np.empty(d)
        b2 = np.empty(2 * Np)
        act = np.empty(Np, dtype=np.int64)
        self.L.embaref_get_normal_eq(self.h, _p(A11), _p(A12), _p(A22), _p(b1), _p(b2), _p(act, C.c_int64))
        return A11, A12, A22, b1, b2, act

    def a12_nnz(self):
        return self.L.embaref_a12_nnz(self.h)

    def solve(self, lam, use_cg=False, fix_first=True):
        d = 3 * (self.n_poses - (1 if fix_first else 0))
        x1 = np.empty(d)
        x2 = np.empty(2 * self.Np)
        it = C.c_int(0)
        err = C.c_double(0)
        self.L.embaref_solve(self.h, float(lam), int(use_cg), int(fix_first), _p(x1), _p(x2), C.byref(it),
                             C.byref(err))
        return x1, x2, it.value, err.value

    def update_map(self, Gx, Gy, x2, damping):
        Gx = np.array(Gx, dtype=np.float64, copy=True, order="C")
        Gy = np.array(Gy, dtype=np.float64, copy=True, order="C")
        x2 = np.ascontiguousarray(x2, dtype=np.float64)
        self.L.embaref_update_map(self.h, _p(Gx), _p(Gy), _p(x2), float(damping))
        return Gx, Gy

    def update_traj(self, traj: RefTraj, x1, fix_first):
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        self.L.embaref_update_traj(self.h, traj.h, _p(x1), x1.size, int(fix_first))

    def solve_time_window(self, traj: RefTraj, Gx, Gy, *, max_num_iter=50, tol_fun=1e-3, num_times_tol_fun_sat=2,
                          use_cg=False, irls_type=0, eta=1.0, thres=5, damping=1.0, alpha=5.0, first_window=True):
        Gx = np.array(Gx, dtype=np.float64, copy=True, order="C")
        Gy = np.array(Gy, dtype=np.float64, copy=True, order="C")
        cap = max_num_iter + 8
        log = np.zeros((cap, 6))
        fc = C.c_double(0)
        hp = C.c_void_p(traj.h.value)
        n = self.L.embaref_solve_time_window(self.h, C.byref(hp), _p(Gx), _p(Gy), max_num_iter, tol_fun,
                                             num_times_tol_fun_sat, int(use_cg), irls_type, float(eta), thres,
                                             float(damping), float(alpha), int(first_window), _p(log), cap,
                                             C.byref(fc))
        traj.h = hp  # the pointee was replaced on every accepted step
        return traj.quat(), Gx, Gy, log[:n].copy(), fc.value

    def timers(self):
        t = np.zeros(3)
        c = (C.c_long * 3)()
        self.L.embaref_get_timers(self.h, _p(t), c)
        return dict(form_s=t[0], solve_s=t[1], obj_s=t[2], form_n=c[0], solve_n=c[1], obj_n=c[2])
