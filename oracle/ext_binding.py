"""ctypes binding of oracle/_ref/libemba_ext_ref.so (oracle/ext_capi.cpp): the fp64 CPU restatement of the EXTENSION
MODE (cubic per-event SO(3) spline on the reference's vendored basalt, bilinear map sampling), plus a numpy assembly
of its normal equations. TEST INFRASTRUCTURE ONLY: never imported by emba_b200/."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libemba_ext_ref.so")
_dp = C.POINTER(C.c_double)
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def _p(a, ty=C.c_double):
    return a.ctypes.data_as(C.POINTER(ty))


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.embaext_rows.restype = C.c_long
        L.embaext_rows.argtypes = [C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_double, C.c_long, C.POINTER(C.c_uint16),
                                   C.POINTER(C.c_uint16), C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.c_int, C.c_int64,
                                   C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.embaext_cpred.restype = C.c_double
        L.embaext_cpred.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int64,
                                    _dp, _dp, _dp]
        _lib = L
    return _lib


def rows(sensor_w, sensor_h, lut, pano_w, pano_h, C_th, x, y, t_ns, pol, t0_ns, dt_ns, quat, Gx, Gy):
    """dict of per-measurement arrays: e, dp, pm, Jc, Jp, w, cp, pix, ev (time order of the current event)"""
    L = lib()
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.uint16); y = np.ascontiguousarray(y, dtype=np.uint16)
    t = np.ascontiguousarray(t_ns, dtype=np.int64); p = np.ascontiguousarray(pol, dtype=np.uint8)
    q = np.ascontiguousarray(quat, dtype=np.float64)
    gx = np.ascontiguousarray(Gx, dtype=np.float64); gy = np.ascontiguousarray(Gy, dtype=np.float64)
    N = t.size
    e = np.empty(N); dp = np.empty((N, 2)); pm = np.empty((N, 2)); Jc = np.empty((N, 12)); Jp = np.empty((N, 12))
    w = np.empty((N, 4)); cp = np.empty((N, 2), np.int32); pix = np.empty((N, 4), np.int32); ev = np.empty(N, np.int32)
    M = L.embaext_rows(sensor_w, sensor_h, _p(lut), pano_w, pano_h, float(C_th), N, _p(x, C.c_uint16), _p(y, C.c_uint16),
                       _p(t, C.c_int64), _p(p, C.c_uint8), q.shape[0], int(t0_ns), int(dt_ns), _p(q), _p(gx), _p(gy), _p(e),
                       _p(dp), _p(pm), _p(Jc), _p(Jp), _p(w), _p(cp, C.c_int32), _p(pix, C.c_int32), _p(ev, C.c_int32))
    if M < 0:
        raise ValueError("timestamp outside the cubic spline's support")
    return dict(e=e[:M].copy(), dp=dp[:M].copy(), pm=pm[:M].copy(), Jc=Jc[:M].copy(), Jp=Jp[:M].copy(), w=w[:M].copy(),
                cp=cp[:M].copy(), pix=pix[:M].copy(), ev=ev[:M].copy())


def cpred(lut, sensor_pix, pano_w, pano_h, t_c, t_p, t0_ns, dt_ns, quat, Gx, Gy):
    L = lib()
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    q = np.ascontiguousarray(quat, dtype=np.float64)
    gx = np.ascontiguousarray(Gx, dtype=np.float64); gy = np.ascontiguousarray(Gy, dtype=np.float64)
    return L.embaext_cpred(_p(lut), int(sensor_pix), pano_w, pano_h, int(t_c), int(t_p), q.shape[0], int(t0_ns), int(dt_ns),
                           _p(q), _p(gx), _p(gy))


def active_set(r, P, thres):
    """pixels whose bilinear footprints are hit by at least `thres` measurements; a measurement is used iff all four
    pixels of its footprint are active"""
    cnt = np.bincount(r["pix"].reshape(-1), minlength=P)
    act = cnt >= thres
    use = act[r["pix"]].all(1)
    return cnt, np.nonzero(act)[0], use


def normal_equations(r, n_poses, P, thres, alpha, Gx, Gy):
    """dense H = J^T J (+ alpha on the map block), g = J^T e (- alpha G) of the extension mode on the used rows;
    unknown order: 3 n pose components, then (Gx, Gy) per ACTIVE pixel in ascending pixel index. Small problems only."""
    cnt, act, use = active_set(r, P, thres)
    amap = -np.ones(P, dtype=np.int64)
    amap[act] = np.arange(act.size)
    d = 3 * n_poses + 2 * act.size
    H = np.zeros((d, d)); g = np.zeros(d)
    for m in np.nonzero(use)[0]:
        cols = np.concatenate([3 * r["cp"][m, 0] + np.arange(12), 3 * r["cp"][m, 1] + np.arange(12),
                               (3 * n_poses + 2 * amap[r["pix"][m]][:, None] + np.arange(2)[None, :]).reshape(-1)])
        vals = np.concatenate([r["Jc"][m], r["Jp"][m], (r["w"][m][:, None] * r["dp"][m][None, :]).reshape(-1)])
        J = np.zeros(d)
        np.add.at(J, cols, vals)  # the two pose blocks (and, at the y border, footprint pixels) may overlap
        H += np.outer(J, J)
        g += J * r["e"][m]
    ia = 3 * n_poses + 2 * np.arange(act.size)
    H[ia, ia] += alpha; H[ia + 1, ia + 1] += alpha
    g[ia] -= alpha * np.asarray(Gx).reshape(-1)[act]
    g[ia + 1] -= alpha * np.asarray(Gy).reshape(-1)[act]
    return H, g, act, use
