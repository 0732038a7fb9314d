"""Generates the golden fixtures in this directory.

Run HERE (the build container), where /root/reference exists and oracle/_ref/libemba_ref.so -- the UNMODIFIED
reference hot-path sources compiled against oracle/shim -- has been built (make -C oracle):

    python tests/golden/make_golden.py

<name>_scene.npz : seeded synthetic scene (events, initial control poses, initial maps)
<name>_ref.npz   : outputs of the reference library on that scene: residuals, num_ev_map, per-measurement
                   Jacobians, normal equations, Schur and CG solutions, the LM iteration log and the refined state.
The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emba_b200 import synth  # noqa: E402
from oracle import ref_binding as RB  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ALPHA, THRES, LAM = 5.0, 5, 1e-3


def make(name):
    sc = synth.make_config(name)
    np.savez_compressed(
        os.path.join(HERE, f"{name}_scene.npz"), sensor_w=sc.sensor_w, sensor_h=sc.sensor_h, fx=sc.fx, fy=sc.fy,
        cx=sc.cx, cy=sc.cy, pano_w=sc.pano_w, pano_h=sc.pano_h, C_th=sc.C_th, t_beg=sc.t_beg, dt_knots=sc.dt_knots,
        n_poses=sc.n_poses, x=sc.x, y=sc.y, t_ns=sc.t_ns, pol=sc.pol, quat_init=sc.quat_init,
        Gx_init=sc.Gx_init, Gy_init=sc.Gy_init)
    ref = RB.RefLEGM(sc.sensor_w, sc.sensor_h, sc.fx, sc.fy, sc.cx, sc.cy, sc.C_th, sc.pano_w, sc.pano_h)
    ref.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    tr = RB.RefTraj(sc.t_beg, sc.dt_knots, sc.quat_init)
    ep, num = ref.evaluate(tr, sc.Gx_init, sc.Gy_init, True)
    rec, n_out = ref.dump_measurements()
    # keep the fixture small: every k-th per-measurement record (all of them for the tiny scene)
    stride = max(1, rec.shape[0] // 8000)
    rec_idx = np.arange(0, rec.shape[0], stride)
    rec_jsum = rec[:, 9:21].sum(0)  # column sums of all Jacobian rows (checksum over every measurement)
    rec = rec[rec_idx]
    cost_data = ref.data_cost()
    cost_cauchy = ref.data_cost(1, 0.1)
    cost_huber = ref.data_cost(2, 0.1)
    cost_reg = ref.reg_cost(sc.Gx_init, sc.Gy_init, ALPHA)
    A11, A12, A22, b1, b2, act = ref.form(sc.n_poses, THRES, sc.Gx_init, sc.Gy_init, ALPHA)
    x1, x2, _, _ = ref.solve(LAM, False, True)
    x1c, x2c, cg_it, cg_err = ref.solve(LAM, True, True)
    x1n, x2n, _, _ = ref.solve(LAM, False, False)
    # IRLS (cauchy) normal equations
    I11, I12, I22, ib1, ib2, iact = ref.form(sc.n_poses, THRES, sc.Gx_init, sc.Gy_init, ALPHA, irls_type=1, a=0.1)
    # full LM
    tr2 = RB.RefTraj(sc.t_beg, sc.dt_knots, sc.quat_init)
    q_fin, Gx_fin, Gy_fin, log, fcost = ref.solve_time_window(tr2, sc.Gx_init, sc.Gy_init, alpha=ALPHA, thres=THRES)
    np.savez_compressed(
        os.path.join(HERE, f"{name}_ref.npz"), ep=ep, num_ev_map=num, n_outliers=n_out, rec=rec, rec_idx=rec_idx, rec_jsum=rec_jsum, cost_data=cost_data,
        cost_cauchy=cost_cauchy, cost_huber=cost_huber, cost_reg=cost_reg, A11=A11, A12_nnz=np.count_nonzero(A12),
        A12_rowsum=A12.sum(1), A12_colsum=A12.sum(0), A12_fro=np.linalg.norm(A12), A22=A22, b1=b1, b2=b2,
        active=act, x1=x1, x2=x2, x1_cg=x1c, x2_cg=x2c, cg_iters=cg_it, cg_err=cg_err, x1_nofix=x1n, x2_nofix=x2n,
        irls_A11=I11, irls_b1=ib1, irls_A22=I22, irls_b2=ib2, irls_A12_rowsum=I12.sum(1),
        lm_log=log, lm_final_cost=fcost, q_final=q_fin, Gx_final=Gx_fin, Gy_final=Gy_fin,
        alpha=ALPHA, thres=THRES, lam=LAM)
    print(name, "N", sc.n_events, "M", ep.size, "Np", act.size, "LM solves", log.shape[0], "final cost", fcost)


def make_frontend():
    """Front-end trajectory fixture for the control-pose initialisation: 1 kHz poses of the
    synthetic ground-truth trajectory with 0.1 deg noise, and the control poses the reference fits to them."""
    rng = np.random.default_rng(3)
    t_beg, t_end, dt = 0.1, 2.4, 0.05
    tp = np.arange(0.0995, 2.4015, 0.001) + rng.uniform(-2e-4, 2e-4, 2302)
    tp.sort()
    t_ns = np.round(tp * 1e9).astype(np.int64)
    q = synth._rotvec_to_quat(synth.gt_rotvec(tp) + rng.standard_normal((tp.size, 3)) * np.deg2rad(0.1))
    cps = RB.ref_generate_ctrl_poses_long(t_ns, q, t_beg, t_end, dt)
    np.savez_compressed(os.path.join(HERE, "frontend_ref.npz"), t_ns=t_ns, quat=q, t_beg=t_beg, t_end=t_end,
                        dt_knots=dt, ctrl_poses=cps)
    print("frontend", t_ns.size, "poses ->", cps.shape[0], "control poses")


def make_poisson():
    """Poisson reconstruction fixture: a smooth seeded gradient map plus noise, 96 x 192,
    and the intensity map the compiled reference (poisson_reconstruction.cpp + laplace.cpp) reconstructs from it."""
    rng = np.random.default_rng(11)
    H, W = 96, 192
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.sin(xx / 17.0) * np.cos(yy / 11.0) + 0.3 * np.sin((xx + 2 * yy) / 29.0)
    Gx = np.zeros((H, W)); Gy = np.zeros((H, W))
    Gx[:, :-1] = img[:, 1:] - img[:, :-1]
    Gy[:-1, :] = img[1:, :] - img[:-1, :]
    Gx += 0.05 * rng.standard_normal((H, W))
    Gy += 0.05 * rng.standard_normal((H, W))
    M = RB.ref_poisson_reconstruct(Gx, Gy)
    np.savez_compressed(os.path.join(HERE, "poisson_ref.npz"), Gx=Gx, Gy=Gy, img=M)
    print("poisson", H, W, "max |img|", np.abs(M).max())


if __name__ == "__main__":
    if "poisson" in sys.argv[1:]:
        make_poisson()
        sys.exit(0)
    if not sys.argv[1:]:
        make_poisson()
    if not sys.argv[1:] or "frontend" in sys.argv[1:]:
        make_frontend()
        if "frontend" in sys.argv[1:]:
            sys.exit(0)
    for nm in (sys.argv[1:] or ["tiny", "small"]):
        make(nm)
