"""Extension mode (cubic per-event SO(3) spline, bilinear map sampling). CPU part: the oracle
(oracle/ext_capi.cpp on the reference's vendored basalt) against central finite differences of the model itself --
there is no reference implementation of this mode, so this is what pins the oracle. GPU part: the CUDA extension
kernels against that oracle."""
import numpy as np
import pytest

from conftest import GoldenScene, rel

ALPHA, THRES = 5.0, 5


def _scene():
    """the tiny golden scene with a cubic control-pose set around it (3 more knots, start one interval earlier)"""
    from oracle import emba_oracle as O

    sc = GoldenScene("tiny")
    t0, dt = O.spline_base_ns(sc.t_beg, sc.dt_knots)
    t0c = t0 - dt
    n = sc.n_poses + 3
    rng = np.random.default_rng(4)
    # knots: the linear scene's control poses extended at both ends, slightly perturbed
    q = np.concatenate([sc.quat_init[:1], sc.quat_init, sc.quat_init[-1:], sc.quat_init[-1:]], 0)
    dq = np.concatenate([0.5 * rng.standard_normal((n, 3)) * np.deg2rad(0.5), np.ones((n, 1))], 1)
    q = O.quat_normalize(O.quat_mul(O.quat_normalize(dq), q))
    return sc, t0c, dt, n, q


def _need_ext():
    from oracle import ext_binding as XB

    if not XB.available():
        pytest.skip("oracle/_ref/libemba_ext_ref.so not built")
    return XB


def test_ext_oracle_jacobians_match_finite_differences():
    from oracle import emba_oracle as O

    XB = _need_ext()
    sc, t0c, dt, n, q = _scene()
    lut = sc.bearing_lut()
    r = XB.rows(sc.sensor_w, sc.sensor_h, lut, sc.pano_w, sc.pano_h, sc.C_th, sc.x, sc.y, sc.t_ns, sc.pol, t0c, dt, q,
                sc.Gx_init, sc.Gy_init)
    M = r["e"].size
    assert M > 5000 and np.all(r["cp"] >= 0) and np.all(r["cp"] + 4 <= n)
    assert np.allclose(r["w"].sum(1), 1.0) and np.all(r["w"] >= 0)
    # previous-event index of every row
    last = {}
    prev = np.full(sc.t_ns.size, -1)
    for i in range(sc.t_ns.size):
        k = int(sc.y[i]) * sc.sensor_w + int(sc.x[i])
        prev[i] = last.get(k, -1)
        last[k] = i
    rng = np.random.default_rng(0)
    eps = 1e-6
    for m in rng.choice(M, 12, replace=False):
        i = int(r["ev"][m]); p = int(prev[i])
        sp = int(sc.y[i]) * sc.sensor_w + int(sc.x[i])
        # the interpolant is only piecewise smooth: stay away from cell borders for the finite differences
        fx, fy = r["pm"][m] - np.floor(r["pm"][m])
        if min(fx, 1 - fx, fy, 1 - fy) < 1e-3:
            continue

        def f(qq, gx=sc.Gx_init, gy=sc.Gy_init):
            return XB.cpred(lut, sp, sc.pano_w, sc.pano_h, sc.t_ns[i], sc.t_ns[p], t0c, dt, qq, gx, gy)

        # rotation part: left perturbation of control pose k, component a (basalt's convention, test_spline.cpp:95-132)
        J = np.zeros(3 * n)
        np.add.at(J, 3 * r["cp"][m, 0] + np.arange(12), r["Jc"][m])
        np.add.at(J, 3 * r["cp"][m, 1] + np.arange(12), r["Jp"][m])
        for k in range(max(0, r["cp"][m].min()), min(n, r["cp"][m].max() + 4)):
            for a in range(3):
                d = np.zeros(3); d[a] = eps
                fd = []
                for sgn in (+1, -1):
                    qq = q.copy()
                    dq = np.concatenate([0.5 * sgn * d, [1.0]])
                    qq[k] = O.quat_normalize(O.quat_mul(O.quat_normalize(dq[None]), q[k][None]))[0]
                    fd.append(f(qq))
                num = (fd[0] - fd[1]) / (2 * eps)
                assert abs(num - J[3 * k + a]) < 2e-5 * max(1.0, np.abs(J).max()), (m, k, a, num, J[3 * k + a])
        # map part: d C_pred / d G[pix_j] = w_j * dp
        for j in range(4):
            for comp, G in ((0, sc.Gx_init), (1, sc.Gy_init)):
                Gp = G.copy().reshape(-1); Gm = G.copy().reshape(-1)
                Gp[r["pix"][m, j]] += eps; Gm[r["pix"][m, j]] -= eps
                args = (Gp.reshape(G.shape), sc.Gy_init) if comp == 0 else (sc.Gx_init, Gp.reshape(G.shape))
                args_m = (Gm.reshape(G.shape), sc.Gy_init) if comp == 0 else (sc.Gx_init, Gm.reshape(G.shape))
                num = (f(q, *args) - f(q, *args_m)) / (2 * eps)
                same = r["pix"][m] == r["pix"][m, j]  # clamped rows share a pixel at the y border
                assert abs(num - r["w"][m][same].sum() * r["dp"][m, comp]) < 1e-6


def test_ext_oracle_reduces_to_nearest_pixel_model_on_a_constant_map():
    """with a constant gradient map the bilinear sample equals the nearest-pixel one: residuals depend on the spline
    only, and a cubic spline whose knots are all equal is the constant rotation -> e = C_meas exactly"""
    XB = _need_ext()
    sc, t0c, dt, n, q = _scene()
    q0 = np.repeat(q[:1], n, 0)
    G = np.full((sc.pano_h, sc.pano_w), 0.3)
    r = XB.rows(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.pano_w, sc.pano_h, sc.C_th, sc.x, sc.y, sc.t_ns, sc.pol,
                t0c, dt, q0, G, G)
    assert np.max(np.abs(r["dp"])) < 1e-9
    pol = sc.pol[r["ev"]].astype(np.float64)
    assert np.allclose(r["e"], 2 * (pol - 0.5) * sc.C_th, atol=1e-9)


@pytest.mark.gpu
def test_ext_cuda_rows_and_normal_equations_vs_oracle():
    """CUDA extension kernels (emba_b200/csrc/ext.cu) against the basalt-based oracle: per-pair residuals and the
    2 x 12 + 8 Jacobian entries (spec 1e-6), g = J^T e, the block diagonal of H, the matrix-free operator and the PCG
    solution against a dense numpy solve of the same damped system."""
    from emba_b200.legm import Engine, EventSequence, ExtEngine

    XB = _need_ext()
    sc, t0c, dt, n, q = _scene()
    lut = sc.bearing_lut()
    r = XB.rows(sc.sensor_w, sc.sensor_h, lut, sc.pano_w, sc.pano_h, sc.C_th, sc.x, sc.y, sc.t_ns, sc.pol, t0c, dt, q,
                sc.Gx_init, sc.Gy_init)
    eng = Engine(sc.sensor_w, sc.sensor_h, lut, sc.C_th, sc.pano_w, sc.pano_h)
    seq = EventSequence(sc.x, sc.y, sc.t_ns, sc.pol)
    ext = ExtEngine(eng, seq, 0, sc.t_ns.size)
    ext.set_state(t0c, dt, q, sc.Gx_init, sc.Gy_init)
    cd, cr, M = ext.evaluate(ALPHA)
    g = ext.get_rows()
    inl = g["flag"] > 0
    assert M == r["e"].size == int(inl.sum()) and np.array_equal(g["ev"][inl], r["ev"])
    assert np.array_equal(g["cp"][inl], r["cp"]) and np.array_equal(g["pix"][inl], r["pix"])
    assert rel(r["e"], g["e"][inl]) < 1e-10 and rel(r["dp"], g["dp"][inl]) < 1e-10 and rel(r["pm"], g["pm"][inl]) < 1e-12
    assert rel(r["Jc"], g["J"][inl, :12]) < 1e-9 and rel(r["Jp"], g["J"][inl, 12:]) < 1e-9
    scale = max(np.abs(r["Jc"]).max(), np.abs(r["Jp"]).max())
    assert np.max(np.abs(r["Jc"] - g["J"][inl, :12])) < 1e-9 * scale and np.max(np.abs(r["Jp"] - g["J"][inl, 12:])) < 1e-9 * scale
    assert abs(cd - 0.5 * float(r["e"] @ r["e"])) < 1e-11 * cd
    assert abs(cr - 0.5 * ALPHA * float(np.sum(sc.Gx_init ** 2) + np.sum(sc.Gy_init ** 2))) < 1e-12 * cr
    # normal equations in Jacobian form
    P = sc.pano_w * sc.pano_h
    H, gv, act, use = XB.normal_equations(r, n, P, THRES, ALPHA, sc.Gx_init, sc.Gy_init)
    Np, Mu = ext.form(THRES, ALPHA)
    gg, Bp, Bm, act_g = ext.get_normal_eq()
    assert Np == act.size and np.array_equal(act_g, act) and Mu == int(use.sum())
    assert np.array_equal(g["flag"][inl] > 0, np.ones(M, bool)) and np.array_equal(ext.get_rows()["flag"][inl] == 3, use)
    assert rel(gv, gg) < 1e-9
    for k in range(n):
        assert np.max(np.abs(H[3 * k:3 * k + 3, 3 * k:3 * k + 3] - Bp[k])) < 1e-9 * max(1.0, np.abs(H[:3 * n, :3 * n]).max())
    ia = 3 * n + 2 * np.arange(Np)
    assert rel(H[ia, ia], Bm[:, 0]) < 1e-9 and rel(H[ia, ia + 1], Bm[:, 1]) < 1e-9 and rel(H[ia + 1, ia + 1], Bm[:, 2]) < 1e-9
    lam = 1e-2
    rng = np.random.default_rng(1)
    v = rng.standard_normal(H.shape[0])
    Hd = H + lam * np.diag(np.diag(H))
    assert rel(Hd @ v, ext.matvec(lam, ALPHA, v)) < 1e-9
    # control poses no event reaches (the padding knots at both ends) have zero rows: their update is zero, the rest
    # of the damped system is SPD
    x, it, err = ext.solve(lam, ALPHA, max_iter=2000, tol=1e-11)
    live = np.nonzero(np.diag(Hd) != 0.0)[0]
    dead = np.nonzero(np.diag(Hd) == 0.0)[0]
    x_ref = np.linalg.solve(Hd[np.ix_(live, live)], gv[live])
    assert err < 1e-9 and np.all(x[dead] == 0.0) and rel(x_ref, x[live]) < 1e-5
    ext.close(); seq.close(); eng.close()


@pytest.mark.gpu
def test_ext_lm_steps_decrease_the_cost():
    """A short LM loop driven through the extension ABI (evaluate -> form -> PCG solve -> apply, accept / reject on the
    host like solver.cpp:299-352) on the small scene: the cost decreases and the state stays on SO(3)."""
    from emba_b200.legm import Engine, EventSequence, ExtEngine
    from oracle import emba_oracle as O

    sc = GoldenScene("small")
    t0, dt = O.spline_base_ns(sc.t_beg, sc.dt_knots)
    q = np.concatenate([sc.quat_init[:1], sc.quat_init, sc.quat_init[-1:], sc.quat_init[-1:]], 0)
    eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h)
    seq = EventSequence(sc.x, sc.y, sc.t_ns, sc.pol)
    ext = ExtEngine(eng, seq, 0, sc.t_ns.size)
    gx, gy = sc.Gx_init.copy(), sc.Gy_init.copy()
    ext.set_state(t0 - dt, dt, q, gx, gy)
    cd, cr, M = ext.evaluate(ALPHA)
    cost = cd + cr
    lam, costs, accepted = 1e-3, [cost], 0
    for _ in range(8):
        ext.form(THRES, ALPHA)
        x, it, err = ext.solve(lam, ALPHA, max_iter=100, tol=1e-6)
        assert np.isfinite(x).all()
        ext.apply(1.0)
        cd, cr, M2 = ext.evaluate(ALPHA)
        if cd + cr < cost:
            cost = cd + cr; lam /= 10; accepted += 1
            q, gx, gy = ext.get_state()
        else:
            lam *= 10
            ext.set_state(t0 - dt, dt, q, gx, gy)
            ext.evaluate(ALPHA)
        costs.append(cost)
    assert accepted >= 2 and costs[-1] < 0.7 * costs[0]
    qf, _, _ = ext.get_state()
    assert np.all(np.abs(np.linalg.norm(qf, axis=1) - 1) < 1e-12)
    ext.close(); seq.close(); eng.close()
