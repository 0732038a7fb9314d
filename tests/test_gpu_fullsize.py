"""GPU tests at BASELINE.json's sizes: configuration C1 (~0.93 M events, playroom.launch shape) against the CPU
oracle element by element, configuration C2 (~10 M events) element by element against the compiled reference
(oracle/_ref) and through size-independent properties (bitwise determinism, count conservation, symmetry, two
independent solvers agreeing, monotone LM), C3 and the headline configuration C4 (~120 M events, 2048x1024, n = 201)
through the same properties."""
import numpy as np
import pytest

from conftest import rel

pytestmark = pytest.mark.gpu
ALPHA, THRES = 5.0, 5


def _setup(name):
    from emba_b200 import synth
    from emba_b200.legm import Engine, spline_base_ns

    sc = synth.make_config(name, device="cuda")
    eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h)
    eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    t0, dt = spline_base_ns(sc.t_beg, sc.dt_knots)
    eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
    return sc, eng, t0, dt


def test_c1_against_oracle():
    from oracle import emba_oracle as O

    sc, eng, t0, dt = _setup("C1")
    assert 800_000 < sc.n_events < 1_100_000 and sc.n_poses == 47
    orc = O.Oracle(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.pano_w, sc.pano_h, sc.C_th)
    orc.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    ep_o, num_o = orc.evaluate(sc.quat_init, t0, dt, sc.Gx_init, sc.Gy_init, True)
    cd, cr, M = eng.evaluate(0, 0, 1.0, ALPHA)
    ep, num = eng.get_evaluation(0, M)
    assert M == ep_o.size and np.array_equal(num, num_o)
    assert rel(ep_o, ep) < 1e-10 and abs(cd - 0.5 * ep_o @ ep_o) < 1e-11 * cd
    Np = eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
    A11, A12, A22, b1, b2, act = eng.get_normal_eq(True)
    B11, B12, B22, c1, c2, act_o = orc.form_normal_eq(sc.n_poses, THRES)
    B22, c2 = orc.apply_l2_reg(B22, c2, act_o, ALPHA, sc.Gx_init, sc.Gy_init)
    assert Np == act_o.size and np.array_equal(act, act_o)
    for a, b in ((B11, A11), (B12, A12), (B22, A22), (c1, b1), (c2, b2)):
        assert rel(a, b) < 1e-9
    x1, x2, _, _ = eng.solve(1e-3, False, True)
    G11, G12, g1 = orc.gauge_fix(B11, B12, c1)
    y1, y2 = orc.solve_normal_eq(G11, G12, B22, g1, c2, 1e-3)
    assert rel(y1, x1) < 1e-6 and rel(y2, x2) < 1e-6
    eng.close()


def test_c2_elementwise_vs_compiled_reference():
    """C2 (10.4 M events, 240x180, 1024x512, n = 97) against the UNMODIFIED reference sources compiled into
    oracle/_ref/libemba_ref.so: integer outputs identical, residuals 1e-10, assembled H / g 1e-9 (BASELINE.json)."""
    from oracle import ref_binding as RB

    if not RB.available():
        pytest.skip("oracle/_ref/libemba_ref.so not built")
    sc, eng, t0, dt = _setup("C2")
    ref = RB.RefLEGM(sc.sensor_w, sc.sensor_h, sc.fx, sc.fy, sc.cx, sc.cy, sc.C_th, sc.pano_w, sc.pano_h)
    ref.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    tr = RB.RefTraj(sc.t_beg, sc.dt_knots, sc.quat_init)
    ep_r, num_r = ref.evaluate(tr, sc.Gx_init, sc.Gy_init, True)
    cd, cr, M = eng.evaluate(0, 0, 1.0, ALPHA)
    ep, num = eng.get_evaluation(0, M)
    assert M == ep_r.size and np.array_equal(num, num_r)
    assert rel(ep_r, ep) < 1e-10 and np.max(np.abs(ep_r - ep)) < 1e-9
    assert abs(cd - 0.5 * float(ep_r @ ep_r)) < 1e-11 * cd
    assert abs(cr - ref.reg_cost(sc.Gx_init, sc.Gy_init, ALPHA)) < 1e-12 * cr
    B11, _, B22, c1, c2, act_r = ref.form(sc.n_poses, THRES, sc.Gx_init, sc.Gy_init, ALPHA, want_A12=False)
    Np = eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
    A11, _, A22, b1, b2, act = eng.get_normal_eq(False)
    assert Np == act_r.size and np.array_equal(act, act_r)
    for a, b in ((B11, A11), (B22, A22), (c1, b1), (c2, b2)):
        assert rel(a, b) < 1e-9
    assert np.max(np.abs(B11 - A11)) < 1e-9 * np.abs(B11).max()
    eng.close()


def test_c2_properties():
    sc, eng, t0, dt = _setup("C2")
    assert 9_000_000 < sc.n_events < 12_000_000 and sc.n_poses == 97
    # bitwise determinism of the whole pass + solve
    out = []
    for _ in range(2):
        eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
        cd, cr, M = eng.evaluate(0, 0, 1.0, ALPHA)
        _, num = eng.get_evaluation(0, None, False, True)
        Np = eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
        A11, _, A22, b1, b2, act = eng.get_normal_eq(False)
        x1, x2, _, _ = eng.solve(1e-3, False, True)
        out.append((cd, cr, M, num, Np, A11, A22, b1, b2, act, x1, x2))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    cd, cr, M, num, Np, A11, A22, b1, b2, act, x1, x2 = out[0]
    # conservation: every inlier measurement lands in exactly one panorama pixel; active set = thresholded histogram
    assert int(num.sum()) == M and M <= eng.num_pairs() and num.min() >= 0
    assert np.array_equal(act, np.nonzero(num.reshape(-1) >= THRES)[0])
    # structure: A11 symmetric, banded by the largest pose gap of a pair; A22 blocks positive definite (alpha > 0)
    assert np.max(np.abs(A11 - A11.T)) <= 1e-12 * np.abs(A11).max()
    assert np.all(A22[:, 0, 0] >= ALPHA) and np.all(A22[:, 1, 1] >= ALPHA)
    assert np.all(A22[:, 0, 0] * A22[:, 1, 1] - A22[:, 0, 1] ** 2 > 0)
    assert np.all(np.diag(A11)[3:] > 0)
    # two independent solvers (direct Schur vs Jacobi-PCG on the full system) agree on the pose update
    y1, y2, it, err = eng.solve(1e-1, True, True)
    z1, z2, _, _ = eng.solve(1e-1, False, True)
    assert err < 1e-4 and rel(z1, y1) < 5e-2 and rel(z2, y2) < 5e-2
    # LM: accepted steps decrease the cost monotonically, the gauge pose is untouched
    eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
    log, fc = eng.solve_time_window(max_num_iter=7, alpha=ALPHA, thres=THRES)
    acc = log[log[:, 4] == 1]
    assert acc.shape[0] >= 2 and np.all(np.diff(acc[:, 3]) < 0) and fc < log[0, 2]
    assert np.all(acc[:, 3] < acc[:, 2])
    q, gx, gy = eng.get_state(0)
    assert np.array_equal(q[0], sc.quat_init[0]) and np.all(np.abs(np.linalg.norm(q, axis=1) - 1) < 1e-12)
    inactive = np.ones(gx.size, dtype=bool)
    assert np.isfinite(gx).all() and np.isfinite(gy).all()
    eng.close()


@pytest.mark.parametrize("pano_w,pano_h,dt_knots", [(512, 256, 0.1), (4096, 2048, 0.01)])
def test_c5_sweep_corners(pano_w, pano_h, dt_knots):
    """BASELINE.json's C5 sweep, its two corners: the coarsest panorama with the widest control-point spacing (most
    rows per map pixel -> scatter contention on the map side) and the finest panorama with dense control points
    (largest per-pixel buffers, long LDL^T). Size-independent properties; Schur and PCG must agree."""
    from emba_b200 import synth
    from emba_b200.legm import Engine, spline_base_ns

    sc = synth.make_config("C2", pano_w=pano_w, pano_h=pano_h, dt_knots=dt_knots, t_end=1.1, device="cuda")
    assert sc.n_poses == int(round(1.0 / dt_knots)) + 1 and sc.n_events > 500_000
    eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h)
    eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    t0, dt = spline_base_ns(sc.t_beg, sc.dt_knots)
    eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
    cd, cr, M = eng.evaluate(0, 0, 1.0, ALPHA)
    _, num = eng.get_evaluation(0, None, False, True)
    assert int(num.sum()) == M and 0 < M <= eng.num_pairs()
    Np = eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
    A11, _, A22, b1, b2, act = eng.get_normal_eq(False)
    assert np.array_equal(act, np.nonzero(num.reshape(-1) >= THRES)[0]) and Np == act.size
    assert np.max(np.abs(A11 - A11.T)) <= 1e-12 * np.abs(A11).max() and np.all(np.diag(A11)[3:] > 0)
    assert np.all(A22[:, 0, 0] * A22[:, 1, 1] - A22[:, 0, 1] ** 2 > 0)
    y1, y2, it, err = eng.solve(1e-1, True, True)
    z1, z2, _, _ = eng.solve(1e-1, False, True)
    assert err < 1e-4 and rel(z1, y1) < 5e-2 and rel(z2, y2) < 5e-2
    log, fc = eng.solve_time_window(max_num_iter=5, alpha=ALPHA, thres=THRES)
    acc = log[log[:, 4] == 1]
    assert acc.shape[0] >= 1 and np.all(np.diff(acc[:, 3]) < 0) and fc < log[0, 2]
    img = eng.reconstruct_map(0)
    assert img.shape == (pano_h, pano_w) and np.isfinite(img).all()
    eng.close()


def test_c3_properties():
    """Configuration C3 (~35 M events, 2048x1024 panorama, n = 201): mean pose window 76 > 64 poses, so most strips take
    k_pix's global-memory path and the solve fills their occupancy masks lazily. Size-independent properties."""
    sc, eng, t0, dt = _setup("C3")
    assert 25_000_000 < sc.n_events < 45_000_000 and sc.n_poses == 201
    out = []
    for _ in range(2):
        eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
        cd, cr, M = eng.evaluate(0, 0, 1.0, ALPHA)
        Np = eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
        A11, _, A22, b1, b2, act = eng.get_normal_eq(False)
        x1, x2, _, _ = eng.solve(1e-3, False, True)
        out.append((cd, cr, M, Np, A11, A22, b1, b2, act, x1, x2))
    for a, b in zip(out[0], out[1]):  # bitwise determinism
        assert np.array_equal(np.asarray(a), np.asarray(b))
    cd, cr, M, Np, A11, A22, b1, b2, act, x1, x2 = out[0]
    _, num = eng.get_evaluation(0, None, False, True)
    assert int(num.sum()) == M and np.array_equal(act, np.nonzero(num.reshape(-1) >= THRES)[0])
    assert eng.a12_entries() / (6.0 * Np) > 64  # the long-window path really is exercised
    assert np.max(np.abs(A11 - A11.T)) <= 1e-12 * np.abs(A11).max()
    assert np.isfinite(x1).all() and np.isfinite(x2).all()
    # the direct solution satisfies the pose block of the damped normal equations after eliminating the map:
    # compare with the PCG solution of the full system (independent code path, no Schur tiles, no masks)
    y1, y2, it, err = eng.solve(1e-1, True, True)
    z1, z2, _, _ = eng.solve(1e-1, False, True)
    assert err < 1e-4 and rel(z1, y1) < 5e-2 and rel(z2, y2) < 5e-2
    log, fc = eng.solve_time_window(max_num_iter=3, alpha=ALPHA, thres=THRES)
    acc = log[log[:, 4] == 1]
    assert acc.shape[0] >= 1 and np.all(np.diff(acc[:, 3]) < 0) and fc < log[0, 2]
    eng.close()


def test_c4_headline_properties():
    """Configuration C4, the one the metric is quoted on (~120 M events, 2048x1024 panorama, n = 201, one GPU):
    size-independent properties. Bitwise equality of two passes is also the test of the per-pixel row ordering -- the
    rows reach their pixel segments through atomics in a different order every run, only the segment sort makes the
    summation order (and with it every bit of A12 / A22 / b2 / x) reproducible."""
    sc, eng, t0, dt = _setup("C4")
    assert 90_000_000 < sc.n_events < 140_000_000 and sc.n_poses == 201 and (sc.pano_w, sc.pano_h) == (2048, 1024)
    out = []
    for _ in range(2):
        eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
        cd, cr, M = eng.evaluate(0, 0, 1.0, ALPHA)
        Np = eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
        A11, _, A22, b1, b2, act = eng.get_normal_eq(False)
        x1, x2, _, _ = eng.solve(1e-3, False, True)
        out.append((cd, cr, M, Np, A11, A22, b1, b2, act, x1, x2))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    cd, cr, M, Np, A11, A22, b1, b2, act, x1, x2 = out[0]
    _, num = eng.get_evaluation(0, None, False, True)
    c = eng.counters()
    assert int(num.sum()) == M and M <= eng.num_pairs() == c["pairs_window"]
    assert np.array_equal(act, np.nonzero(num.reshape(-1) >= THRES)[0]) and Np == act.size
    assert c["rows_on_active_rank"] == int(num.reshape(-1)[act].sum())  # every row of an active pixel has a segment slot
    assert num.max() > 1024 and c["long_segments"] == int((num.reshape(-1)[act] > 1024).sum())  # long-segment path exercised
    assert np.max(np.abs(A11 - A11.T)) <= 1e-12 * np.abs(A11).max() and np.all(np.diag(A11)[3:] > 0)
    assert np.all(A22[:, 0, 0] >= ALPHA) and np.all(A22[:, 0, 0] * A22[:, 1, 1] - A22[:, 0, 1] ** 2 > 0)
    assert np.isfinite(x1).all() and np.isfinite(x2).all()
    # linearity of the assembly in the residual: the map-side right-hand side of the atomic path (different kernel,
    # different order) agrees to rounding
    eng.set_map_path(1)
    eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
    _, _, A22a, _, b2a, _ = eng.get_normal_eq(False)
    eng.set_map_path(0)
    assert rel(A22, A22a) < 1e-12 and rel(b2, b2a) < 1e-11
    # two independent solvers agree on the step; LM decreases the cost
    eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
    y1, y2, it, err = eng.solve(1e-1, True, True)
    z1, z2, _, _ = eng.solve(1e-1, False, True)
    assert err < 1e-4 and rel(z1, y1) < 5e-2 and rel(z2, y2) < 5e-2
    log, fc = eng.solve_time_window(max_num_iter=3, alpha=ALPHA, thres=THRES)
    acc = log[log[:, 4] == 1]
    assert acc.shape[0] >= 1 and np.all(np.diff(acc[:, 3]) < 0) and fc < log[0, 2]
    q, gx, gy = eng.get_state(0)
    assert np.array_equal(q[0], sc.quat_init[0]) and np.all(np.abs(np.linalg.norm(q, axis=1) - 1) < 1e-12)
    eng.close()


def test_long_and_huge_pixel_segments_vs_oracle():
    """C1's trajectory over a 48x24 panorama: ~170 panorama pixels share 0.26 M measurements, so every row segment of the map
    side is thousands of rows long -- the register merge path (1025 .. 4096 rows) and the CTA-wide path (> 4096 rows)
    of the segment sort, which the benchmark workloads hardly reach. Element-wise against the numpy oracle."""
    from emba_b200 import synth
    from emba_b200.legm import Engine, spline_base_ns
    from oracle import emba_oracle as O

    sc = synth.make_config("C1", pano_w=48, pano_h=24, device="cuda")
    eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h)
    eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    t0, dt = spline_base_ns(sc.t_beg, sc.dt_knots)
    eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
    orc = O.Oracle(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.pano_w, sc.pano_h, sc.C_th)
    orc.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    ep_o, num_o = orc.evaluate(sc.quat_init, t0, dt, sc.Gx_init, sc.Gy_init, True)
    cd, cr, M = eng.evaluate(0, 0, 1.0, ALPHA)
    _, num = eng.get_evaluation(0, None, False, True)
    assert M == ep_o.size and np.array_equal(num, num_o)
    assert (num > 4096).sum() >= 3 and ((num > 1024) & (num <= 4096)).sum() >= 3  # both long paths really run
    outs = []
    for _ in range(2):
        Np = eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
        outs.append(eng.get_normal_eq(True))
    for a, b in zip(outs[0], outs[1]):  # bit-reproducible although the rows reach their segments in a different order
        assert np.array_equal(a, b)
    A11, A12, A22, b1, b2, act = outs[0]
    assert eng.counters()["long_segments"] == int((num.reshape(-1)[act] > 1024).sum())
    B11, B12, B22, c1, c2, act_o = orc.form_normal_eq(sc.n_poses, THRES)
    B22, c2 = orc.apply_l2_reg(B22, c2, act_o, ALPHA, sc.Gx_init, sc.Gy_init)
    assert np.array_equal(act, act_o)
    for a, b in ((B11, A11), (B12, A12), (B22, A22), (c1, b1), (c2, b2)):
        assert rel(a, b) < 1e-9
    eng.close()
