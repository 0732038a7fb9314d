"""Control-pose initialisation: LinearTrajectory::generateCtrlPosesLong of the reference
(src/utils/trajectory.cpp:258-294) vs the numpy oracle (CPU) and vs the CUDA kernel (GPU)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import emba_oracle as O
from oracle import ref_binding as RB


@pytest.fixture(scope="module")
def fe():
    return np.load(os.path.join(GOLDEN_DIR, "frontend_ref.npz"))


def test_oracle_matches_golden(fe):
    cps = O.generate_ctrl_poses_long(fe["t_ns"], fe["quat"], float(fe["t_beg"]), float(fe["t_end"]), float(fe["dt_knots"]))
    assert cps.shape == fe["ctrl_poses"].shape == (47, 4)
    assert np.max(np.abs(cps - fe["ctrl_poses"])) < 1e-13


@pytest.mark.skipif(not RB.available(), reason="oracle/_ref/libemba_ref.so not built")
def test_oracle_matches_reference_live():
    rng = np.random.default_rng(9)
    t_beg, t_end, dt = 1.0, 3.0, 0.02
    tp = np.sort(rng.uniform(0.999, 3.001, 9000))
    t_ns = np.unique(np.round(tp * 1e9).astype(np.int64))
    q = O.quat_normalize(O.so3_exp(np.stack([0.3 * np.sin(t_ns * 1e-9), 0.2 * t_ns * 1e-9, 0.1 * np.cos(2e-9 * t_ns)], -1)
                                   + rng.standard_normal((t_ns.size, 3)) * 1e-3))
    a = RB.ref_generate_ctrl_poses_long(t_ns, q, t_beg, t_end, dt)
    b = O.generate_ctrl_poses_long(t_ns, q, t_beg, t_end, dt)
    assert a.shape == b.shape == (101, 4) and np.max(np.abs(a - b)) < 1e-13


def test_too_few_poses_is_an_error(fe):
    with pytest.raises(ValueError):  # the reference aborts (CHECK_GE, trajectory.cpp:153)
        O.generate_ctrl_poses_long(fe["t_ns"][::200], fe["quat"][::200], float(fe["t_beg"]), float(fe["t_end"]), 0.05)


@pytest.mark.gpu
def test_cuda_matches_golden_and_oracle(fe):
    from emba_b200.capi import EmbaError
    from emba_b200.legm import fit_control_poses

    t_beg, t_end, dt = float(fe["t_beg"]), float(fe["t_end"]), float(fe["dt_knots"])
    cps = fit_control_poses(fe["t_ns"], fe["quat"], t_beg, t_end, dt)
    assert cps.shape == (47, 4)
    assert np.max(np.abs(cps - fe["ctrl_poses"])) < 1e-12
    # other spacing, against the oracle
    cps2 = fit_control_poses(fe["t_ns"], fe["quat"], t_beg, t_end, 0.01)
    ref2 = O.generate_ctrl_poses_long(fe["t_ns"], fe["quat"], t_beg, t_end, 0.01)
    assert cps2.shape == ref2.shape and np.max(np.abs(cps2 - ref2)) < 1e-12
    with pytest.raises(EmbaError) as ei:
        fit_control_poses(fe["t_ns"][::200], fe["quat"][::200], t_beg, t_end, dt)
    assert ei.value.code == -3
