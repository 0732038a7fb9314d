"""The numpy oracle (oracle/emba_oracle.py) pinned against the UNMODIFIED reference sources compiled here
(oracle/_ref/libemba_ref.so, built by `make -C oracle`), function by function, and against the committed golden
outputs of that library. CPU only."""
import numpy as np
import pytest

from conftest import rel
from oracle import emba_oracle as O
from oracle import ref_binding as RB

needs_ref = pytest.mark.skipif(not RB.available(), reason="oracle/_ref/libemba_ref.so not built (make -C oracle)")


def _oracle(sc):
    orc = O.Oracle(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.pano_w, sc.pano_h, sc.C_th)
    orc.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    return orc


def _ref(sc):
    ref = RB.RefLEGM(sc.sensor_w, sc.sensor_h, sc.fx, sc.fy, sc.cx, sc.cy, sc.C_th, sc.pano_w, sc.pano_h)
    ref.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    return ref


# ---- against the golden vectors (always runs; the fixtures were produced by the reference library) -----------
@pytest.mark.parametrize("name", ["tiny", "small", "seam"])
def test_oracle_matches_golden(name):
    from conftest import GoldenScene, load_golden_ref

    sc, g = GoldenScene(name), load_golden_ref(name)
    orc = _oracle(sc)
    t0, dt = O.spline_base_ns(sc.t_beg, sc.dt_knots)
    ep, num = orc.evaluate(sc.quat_init, t0, dt, sc.Gx_init, sc.Gy_init, True)
    assert ep.size == g["ep"].size and np.array_equal(num, g["num_ev_map"])
    assert rel(g["ep"], ep) < 1e-11
    assert orc.cur.size == ep.size + int(g["n_outliers"])  # every pair is either an inlier or an outlier
    st = orc.state
    rec, idx = g["rec"], g["rec_idx"]
    assert rel(rec[:, 9:15], st["Jc"][idx]) < 1e-11 and rel(rec[:, 15:21], st["Jp"][idx]) < 1e-11
    assert rel(g["rec_jsum"], np.concatenate([st["Jc"].sum(0), st["Jp"].sum(0)])) < 1e-10
    assert np.array_equal(rec[:, 21], st["cp_c"][idx]) and np.array_equal(rec[:, 22], st["cp_p"][idx])
    assert abs(orc.data_cost(ep) - float(g["cost_data"])) < 1e-11 * float(g["cost_data"])
    assert abs(orc.data_cost(ep, 1, 0.1) - float(g["cost_cauchy"])) < 1e-11 * float(g["cost_cauchy"])
    assert abs(orc.data_cost(ep, 2, 0.1) - float(g["cost_huber"])) < 1e-11 * float(g["cost_huber"])
    alpha, thres, lam = float(g["alpha"]), int(g["thres"]), float(g["lam"])
    assert abs(orc.reg_cost(sc.Gx_init, sc.Gy_init, alpha) - float(g["cost_reg"])) < 1e-12 * float(g["cost_reg"])
    A11, A12, A22, b1, b2, act = orc.form_normal_eq(sc.n_poses, thres)
    A22, b2 = orc.apply_l2_reg(A22, b2, act, alpha, sc.Gx_init, sc.Gy_init)
    assert np.array_equal(act, g["active"])
    for a, b in ((g["A11"], A11), (g["A22"], A22), (g["b1"], b1), (g["b2"], b2), (g["A12_rowsum"], A12.sum(1)),
                 (g["A12_colsum"], A12.sum(0))):
        assert rel(a, b) < 1e-11
    assert np.count_nonzero(A12) == int(g["A12_nnz"])
    G11, G12, g1 = orc.gauge_fix(A11, A12, b1)
    x1, x2 = orc.solve_normal_eq(G11, G12, A22, g1, b2, lam)
    assert rel(g["x1"], x1) < 1e-9 and rel(g["x2"], x2) < 1e-9
    y1, y2, it, err = orc.solve_normal_eq_cg(G11, G12, A22, g1, b2, lam)
    assert it == int(g["cg_iters"]) and abs(err - float(g["cg_err"])) < 1e-3 * float(g["cg_err"])
    assert rel(g["x1_cg"], y1) < 1e-8 and rel(g["x2_cg"], y2) < 1e-8
    # IRLS (cauchy)
    I11, I12, I22, ib1, ib2, _ = orc.form_normal_eq(sc.n_poses, thres, 1, 0.1)
    I22, ib2 = orc.apply_l2_reg(I22, ib2, act, alpha, sc.Gx_init, sc.Gy_init)
    assert rel(g["irls_A11"], I11) < 1e-11 and rel(g["irls_b2"], ib2) < 1e-11


def test_oracle_full_lm_matches_golden(tiny, tiny_ref):
    sc, g = tiny, tiny_ref
    orc = _oracle(sc)
    t0, dt = O.spline_base_ns(sc.t_beg, sc.dt_knots)
    q, Gx, Gy, log = orc.solve_time_window(sc.quat_init, t0, dt, sc.Gx_init, sc.Gy_init, alpha=float(g["alpha"]),
                                           thres=int(g["thres"]))
    rlog = g["lm_log"]
    assert log.shape == rlog.shape and np.array_equal(log[:, 4], rlog[:, 4]) and np.array_equal(log[:, 5], rlog[:, 5])
    assert np.max(np.abs(log[:, 3] - rlog[:, 3]) / rlog[:, 3]) < 1e-10
    assert np.max(np.abs(q - g["q_final"])) < 1e-12
    assert rel(g["Gx_final"], Gx) < 1e-10 and rel(g["Gy_final"], Gy) < 1e-10


# ---- live against the compiled reference (here; the library also travels to the GPU box) ----------------------
@needs_ref
def test_bearing_lut_and_spline_vs_reference(tiny):
    sc = tiny
    ref = _ref(sc)
    assert np.array_equal(ref.bearing_lut(), O.bearing_lut(sc.sensor_w, sc.sensor_h, sc.fx, sc.fy, sc.cx, sc.cy))
    tr = RB.RefTraj(sc.t_beg, sc.dt_knots, sc.quat_init)
    t0, dt = O.spline_base_ns(sc.t_beg, sc.dt_knots)
    tq = np.array([t0, t0 + 1, t0 + dt - 1, t0 + dt, t0 + 3 * dt + 12345677, t0 + (sc.n_poses - 1) * dt - 1],
                  dtype=np.int64)
    R, D, cp = O.spline_eval(sc.quat_init, t0, dt, tq)
    for i, t in enumerate(tq):
        Rr, Jr, idx = tr.evaluate(int(t))
        assert idx == cp[i]
        assert np.max(np.abs(Rr - R[i])) < 1e-14 and np.max(np.abs(Jr - D[i])) < 1e-13
    with pytest.raises(ValueError):
        O.spline_eval(sc.quat_init, t0, dt, np.array([t0 + (sc.n_poses - 1) * dt], dtype=np.int64))


@needs_ref
def test_batch_mid_times_odd_spans_vs_reference():
    """ros::Duration*0.5 rounding of odd nanosecond spans (model.cpp:116-119): evaluate the reference spline at the
    batch mid-time indirectly -- two events per pixel far apart make every batch one measurement; compare residuals."""
    rng = np.random.default_rng(3)
    span = rng.integers(1, 2_000_000_000, size=2000) | 1  # odd spans
    t0 = rng.integers(0, 10**9, size=2000)
    t = np.stack([t0, t0 + span], 1)
    got = []
    for a, b in t:
        tt = np.full(100, a, dtype=np.int64)
        tt[-1] = b
        got.append(O.batch_mid_times(np.sort(tt))[0])
    got = np.array(got)
    # exact integer reference: floor/round of the double product as rostime does it
    sec = span // 10**9
    nsec = span - sec * 10**9
    d = (sec.astype(np.float64) + 1e-9 * nsec.astype(np.float64)) * 0.5
    s = np.floor(d)
    ns = np.floor((d - s) * 1e9 + 0.5)
    assert np.array_equal(got, t0 + s.astype(np.int64) * 10**9 + ns.astype(np.int64))
    assert np.all(np.abs(got - (t0 + span / 2.0)) <= 1.0)


@needs_ref
@pytest.mark.parametrize("irls", [0, 1, 2])
def test_oracle_vs_reference_live_perturbed(tiny, irls):
    """A state different from the fixture's (other poses, other map) through both implementations."""
    sc = tiny
    rng = np.random.default_rng(11 + irls)
    q = sc.quat_init + 0.002 * rng.standard_normal(sc.quat_init.shape)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    Gx = sc.Gx_init + 0.01 * rng.standard_normal(sc.Gx_init.shape)
    Gy = sc.Gy_init + 0.01 * rng.standard_normal(sc.Gy_init.shape)
    ref, orc = _ref(sc), _oracle(sc)
    tr = RB.RefTraj(sc.t_beg, sc.dt_knots, q)
    t0, dt = O.spline_base_ns(sc.t_beg, sc.dt_knots)
    ep_r, num_r = ref.evaluate(tr, Gx, Gy, True)
    ep_o, num_o = orc.evaluate(tr.quat(), t0, dt, Gx, Gy, True)
    assert np.array_equal(num_r, num_o) and rel(ep_r, ep_o) < 1e-11
    a = 0.1
    assert abs(ref.data_cost(irls, a) - orc.data_cost(ep_o, irls, a)) < 1e-11 * abs(ref.data_cost(irls, a))
    A11, A12, A22, b1, b2, act = ref.form(sc.n_poses, 3, Gx, Gy, 2.0, irls_type=irls, a=a)
    B11, B12, B22, c1, c2, act2 = orc.form_normal_eq(sc.n_poses, 3, irls, a)
    B22, c2 = orc.apply_l2_reg(B22, c2, act2, 2.0, Gx, Gy)
    assert np.array_equal(act, act2)
    for x, y in ((A11, B11), (A12, B12), (A22, B22), (b1, c1), (b2, c2)):
        assert rel(x, y) < 1e-11
    x1, x2, _, _ = ref.solve(1e-2, False, False)
    y1, y2 = orc.solve_normal_eq(B11, B12, B22, c1, c2, 1e-2)
    assert rel(x1, y1) < 1e-8 and rel(x2, y2) < 1e-8
    gx_r, gy_r = ref.update_map(Gx, Gy, x2, 0.7)
    gx_o, gy_o = orc.update_map(Gx, Gy, x2, 0.7, act)
    assert np.array_equal(gx_r, gx_o) and np.array_equal(gy_r, gy_o)
    ref.update_traj(tr, x1, False)
    assert np.max(np.abs(tr.quat() - orc.update_traj(q / np.linalg.norm(q, axis=1, keepdims=True), x1, False))) < 1e-14


@needs_ref
def test_sobel_shim_vs_opencv_python():
    """The Sobel of oracle/shim/opencv2/core.hpp (what the compiled reference runs) and the oracle's, against the
    real OpenCV (python cv2 is installed): same operator up to summation order."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    A = rng.standard_normal((37, 53))
    B = rng.standard_normal((37, 53))
    Gxx, Gxy, Gyy = O.sobel_hessian(A, B)
    cxx = 0.125 * cv2.Sobel(A, cv2.CV_64F, 1, 0)
    cxy = 0.125 * cv2.Sobel(A, cv2.CV_64F, 0, 1)
    cyx = 0.125 * cv2.Sobel(B, cv2.CV_64F, 1, 0)
    cyy = 0.125 * cv2.Sobel(B, cv2.CV_64F, 0, 1)
    assert np.max(np.abs(Gxx - cxx)) < 1e-14 and np.max(np.abs(Gyy - cyy)) < 1e-14
    assert np.max(np.abs(Gxy - 0.5 * (cxy + cyx))) < 1e-14
