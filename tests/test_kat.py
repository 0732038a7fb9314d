"""Known-answer tests that the reference's own (stale) test programs contain for this path.

* src/test/event_map_test.cpp:131-155 prints the closed-form linear-spline Jacobian
      d(dphi_R)/d(dphi_cp1) = I - A_u,  d(dphi_R)/d(dphi_cp2) = A_u,  A_u = u Jl(u d) Jl^-1(d),  d = Log(R2 R1^-1)
  next to basalt's output (include/utils/so3_funcs.h:39-59 holds Jl / Jl_inv).
* src/test/so3_funcs_test.cpp:8 uses v = (0.5377, 1.8339, -2.2588).
* thirdparty/basalt-headers/test/src/test_spline.cpp:95-132 fixes the perturbation convention (left/left) by
  central differences.
"""
import numpy as np

from oracle import emba_oracle as O


def _Jl_closed(v):  # include/utils/so3_funcs.h:39-47
    phi = np.linalg.norm(v)
    a = v / phi
    t = np.sin(phi) / phi
    return t * np.eye(3) + (1 - t) * np.outer(a, a) + (1 - np.cos(phi)) / phi * O.hat(a)


def _Jl_inv_closed(v):  # include/utils/so3_funcs.h:50-59
    phi = np.linalg.norm(v)
    a = v / phi
    t1 = phi / 2
    t2 = t1 * np.cos(t1) / np.sin(t1)
    return t2 * np.eye(3) + (1 - t2) * np.outer(a, a) - t1 * O.hat(a)


def test_so3_funcs_vector():
    v = np.array([0.5377, 1.8339, -2.2588])  # so3_funcs_test.cpp:8
    assert np.max(np.abs(O.left_jacobian(v) - _Jl_closed(v))) < 1e-14
    assert np.max(np.abs(O.left_jacobian_inv(v) - _Jl_inv_closed(v))) < 1e-14
    assert np.max(np.abs(O.left_jacobian(v) @ O.left_jacobian_inv(v) - np.eye(3))) < 1e-13


def test_linear_spline_closed_form_jacobian():
    rng = np.random.default_rng(0)
    quat = O.quat_normalize(rng.standard_normal((6, 4)) * 0.2 + np.array([0, 0, 0, 1.0]))
    dt = 50_000_000
    tq = np.array([dt // 3, dt + 7, 3 * dt + dt // 2, 4 * dt + dt - 1], dtype=np.int64)
    R, D, cp = O.spline_eval(quat, 0, dt, tq)
    for i, t in enumerate(tq):
        s, u = t // dt, (t % dt) / dt
        R1, R2 = O.quat_to_R(quat[s]), O.quat_to_R(quat[s + 1])
        d = O.so3_log(O.quat_mul(quat[s + 1], O.quat_conj(quat[s])))  # Log(R2 R1^-1), world frame
        A = u * _Jl_closed(u * d) @ _Jl_inv_closed(d)
        assert cp[i] == s
        assert np.max(np.abs(D[i][:, :3] - (np.eye(3) - A))) < 1e-13  # event_map_test.cpp:146-150
        assert np.max(np.abs(D[i][:, 3:] - A)) < 1e-13
        Ru = O.quat_to_R(O.so3_exp(u * d)) @ R1  # R(u) = Exp(u d) R1
        assert np.max(np.abs(R[i] - Ru)) < 1e-14
        assert np.max(np.abs(R2 - O.quat_to_R(O.so3_exp(d)) @ R1)) < 1e-14


def test_spline_jacobian_central_differences():
    """basalt testEvaluateSo3 convention: knot <- Exp(x) knot, output increment Log(R' R^-1)."""
    rng = np.random.default_rng(1)
    quat = O.quat_normalize(rng.standard_normal((4, 4)) * 0.3 + np.array([0, 0, 0, 1.0]))
    dt = 10_000_000
    t = np.array([dt + 3_333_333], dtype=np.int64)
    R, D, cp = O.spline_eval(quat, 0, dt, t)
    eps = 1e-6
    for k in range(2):
        J = np.zeros((3, 3))
        for a in range(3):
            e = np.zeros(3)
            e[a] = eps
            out = []
            for sgn in (+1, -1):
                q = quat.copy()
                q[cp[0] + k] = O.quat_normalize(O.quat_mul(O.so3_exp(sgn * e), q[cp[0] + k]))
                Rp, _, _ = O.spline_eval(q, 0, dt, t)
                dR = Rp[0] @ R[0].T
                out.append(np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2)
            J[:, a] = (out[0] - out[1]) / (2 * eps)
        assert np.max(np.abs(J - D[0][:, 3 * k:3 * k + 3])) < 1e-6


def test_std_round_half_away_from_zero():
    v = np.array([0.5, 1.5, 2.5, -0.5, -1.5, 2.4999999999999996, 3.0, -2.5])
    assert np.array_equal(O.std_round(v), np.array([1, 2, 3, -1, -2, 2, 3, -3], dtype=np.float64))


def test_pairing_is_reference_order():
    x = np.array([1, 0, 1, 1, 0, 2], dtype=np.uint16)
    y = np.zeros(6, dtype=np.uint16)
    cur, prev = O.pairing(x, y, 3, 6)
    # pixel 0: events 1,4 -> pair (4,1); pixel 1: events 0,2,3 -> pairs (2,0),(3,2); pixel 2: single event
    assert cur.tolist() == [4, 2, 3] and prev.tolist() == [1, 0, 2]
