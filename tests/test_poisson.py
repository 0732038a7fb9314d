"""Poisson reconstruction: poisson_reconstruction::reconstructFromGradient of the reference
(src/image_rec/poisson_reconstruction.cpp:9-50, src/image_rec/laplace.cpp:587-797) vs the numpy oracle (CPU) and vs
the CUDA path (GPU, through the C ABI)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import emba_oracle as O
from oracle import ref_binding as RB


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(GOLDEN_DIR, "poisson_ref.npz"))


def divergence(Gx, Gy):
    """F of poisson_reconstruction.cpp:22-30."""
    H, W = Gx.shape
    F = np.zeros((H, W))
    F[:H - 1, :W - 1] = (Gx[:H - 1, 1:] - Gx[:H - 1, :W - 1]) + (Gy[1:, :W - 1] - Gy[:H - 1, :W - 1])
    return F


def laplacian_dirichlet(M):
    """5-point Laplacian with zero values outside the grid: what pde::poisolve inverts (Dirichlet 0, h = 1)."""
    P = np.pad(M, 1)
    return P[:-2, 1:-1] + P[2:, 1:-1] + P[1:-1, :-2] + P[1:-1, 2:] - 4.0 * M


def test_oracle_matches_golden(fx):
    img = O.poisson_reconstruct(fx["Gx"], fx["Gy"])
    assert img.shape == fx["img"].shape == (96, 192)
    assert np.max(np.abs(img - fx["img"])) < 1e-12 * np.max(np.abs(fx["img"]))


def test_oracle_solves_the_discrete_poisson_equation(fx):
    img = O.poisson_reconstruct(fx["Gx"], fx["Gy"])
    F = divergence(fx["Gx"], fx["Gy"])
    assert np.max(np.abs(laplacian_dirichlet(img) - F)) < 1e-11 * max(1.0, np.max(np.abs(F)))


@pytest.mark.skipif(not RB.available(), reason="oracle/_ref/libemba_ref.so not built")
@pytest.mark.parametrize("shape", [(2, 2), (7, 10), (64, 128), (50, 33)])
def test_oracle_matches_reference_live(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    Gx, Gy = rng.standard_normal(shape), rng.standard_normal(shape)
    a = RB.ref_poisson_reconstruct(Gx, Gy)
    b = O.poisson_reconstruct(Gx, Gy)
    assert np.max(np.abs(a - b)) < 1e-12 * max(1e-30, np.max(np.abs(a)))


@pytest.mark.gpu
def test_cuda_matches_golden(fx):
    from emba_b200.legm import PoissonPlan, reconstructFromGradient

    plan = PoissonPlan(192, 96)
    img = plan.reconstruct(fx["Gx"], fx["Gy"])
    ms, launches = plan.last_ms()
    assert launches >= 8 and ms > 0.0
    assert np.max(np.abs(img - fx["img"])) < 1e-11 * np.max(np.abs(fx["img"]))
    # same plan, second call (buffers reused), and the reference-shaped entry point
    img2 = plan.reconstruct(fx["Gx"], fx["Gy"])
    assert np.array_equal(img, img2)
    img3 = reconstructFromGradient(np.stack([fx["Gx"], fx["Gy"]], -1))
    assert np.array_equal(img, img3)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 2), (6, 10), (64, 128), (130, 70), (512, 1024)])
def test_cuda_matches_oracle(shape):
    from emba_b200.legm import PoissonPlan

    rng = np.random.default_rng(shape[0] + 7 * shape[1])
    Gx, Gy = rng.standard_normal(shape), rng.standard_normal(shape)
    img = PoissonPlan(shape[1], shape[0]).reconstruct(Gx, Gy)
    ref = O.poisson_reconstruct(Gx, Gy)
    assert np.max(np.abs(img - ref)) < 1e-11 * np.max(np.abs(ref))


@pytest.mark.gpu
def test_cuda_full_size_solves_the_poisson_equation():
    """2048 x 1024 (C3 / C4 panorama): size-independent property instead of the O(n^3) oracle."""
    from emba_b200.legm import PoissonPlan

    rng = np.random.default_rng(5)
    H, W = 1024, 2048
    Gx, Gy = rng.standard_normal((H, W)), rng.standard_normal((H, W))
    img = PoissonPlan(W, H).reconstruct(Gx, Gy)
    F = divergence(Gx, Gy)
    assert np.max(np.abs(laplacian_dirichlet(img) - F)) < 1e-9 * np.max(np.abs(F))


@pytest.mark.gpu
def test_cuda_odd_size_is_refused():
    from emba_b200.capi import EmbaError
    from emba_b200.legm import PoissonPlan

    with pytest.raises(EmbaError) as ei:
        PoissonPlan(33, 50)
    assert ei.value.code == -3


@pytest.mark.gpu
def test_reconstruct_map_of_the_optimiser_state():
    """emba_reconstruct_map works on the handle's device-resident maps (solver.cpp:412-417)."""
    from conftest import GoldenScene
    from emba_b200.legm import Engine, spline_base_ns

    sc = GoldenScene("tiny")
    eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h)
    t0, dt = spline_base_ns(sc.t_beg, sc.dt_knots)
    eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
    img = eng.reconstruct_map(0)
    ref = O.poisson_reconstruct(sc.Gx_init.reshape(sc.pano_h, sc.pano_w), sc.Gy_init.reshape(sc.pano_h, sc.pano_w))
    assert np.max(np.abs(img - ref)) < 1e-11 * np.max(np.abs(ref))
