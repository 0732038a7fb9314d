"""The reference-side binding (include/emba_b200_legm.hpp) must compile against the reference's own headers.
Only possible where /root/reference exists (this container); skipped on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

TU = r'''
#include "emba_b200_legm.hpp"
int use_it(const sensor_msgs::CameraInfo& ci) {
  EMBA::LEGM_B200 m(ci, 0.2, 1024, 512);
  EMBA::EventPacket ev;
  cv::Mat Gx = cv::Mat::zeros(512, 1024, CV_64FC1), Gy = Gx.clone(), num;
  std::vector<Sophus::SO3d> cps(4);
  Trajectory* traj = new LinearTrajectory(0.1, 0.05, cps);
  EMBA::VecXd ep = m.evaluateDataError(traj, Gx, Gy, ev, true, num);
  EMBA::MatXd A11, A12; std::vector<EMBA::Mat2d> A22; EMBA::VecXd b1, b2, x1, x2;
  std::set<size_t> act, inact;
  m.formNormalEq(A11, A12, A22, b1, b2, ep, 4, num, 5, act, inact);
  m.formNormalEqIRLS(A11, A12, A22, b1, b2, ep, 4, num, 5, act, inact, "cauchy", 0.1);
  m.applyL2Reg(A22, b2, act, 5.0, Gx, Gy);
  m.solveNormalEq(A11, A12, A22, b1, b2, 1e-3, x1, x2);
  std::pair<int, double> r = m.solveNormalEqCG(A11, A12, A22, b1, b2, 1e-3, x1, x2);
  m.updateTraj(traj, x1, 1);
  m.updateMap(Gx, Gy, x2, 1.0, act, inact);
  emba_lm_settings_t s{};
  m.solveTimeWindowOnDevice(traj, Gx, Gy, ev, s);
  cv::Mat grad(512, 1024, CV_64FC2);
  cv::Mat img = poisson_reconstruction_b200::reconstructFromGradient(grad);
  return r.first + (int)m.evaluateRegError(Gx, Gy).size() + (int)m.evaluateRobustDataCost(ep, "huber", 0.1);
}
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include", "emba")), reason="reference headers not present")
def test_adapter_compiles_against_reference_headers(tmp_path):
    src = tmp_path / "adapter_tu.cpp"
    src.write_text(TU)
    B = os.path.join(REF, "thirdparty", "basalt-headers")
    cmd = ["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-w", f"-I{ROOT}/include", f"-I{ROOT}/oracle/shim",
           f"-I{REF}/include", f"-I{B}/include", f"-I{B}/thirdparty/Sophus", f"-I{B}/thirdparty/eigen", str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_c_header_is_plain_c(tmp_path):
    """include/emba_b200.h must be consumable from C (cgo / ctypes-style FFI generators)."""
    src = tmp_path / "t.c"
    src.write_text('#include "emba_b200.h"\nint main(void){ emba_config_t c; (void)c; return EMBA_OK; }\n')
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", f"-I{ROOT}/include", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
