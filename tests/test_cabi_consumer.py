"""A plain-C program (tests/cabi/cabi_consumer.c, compiled with gcc against include/emba_b200.h and linked to
libemba_b200.so) drives the hot path through the C ABI alone. CPU: it compiles and links. GPU: its output matches
the golden outputs of the reference."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cabi", "cabi_consumer.c")
LIBDIR = os.path.join(ROOT, "emba_b200")


def _build(tmp_path):
    exe = str(tmp_path / "cabi_consumer")
    cmd = ["/usr/bin/gcc", "-std=c99", "-O1", "-Wall", "-Werror", f"-I{ROOT}/include", SRC, "-o", exe, f"-L{LIBDIR}",
           "-lemba_b200", f"-Wl,-rpath,{LIBDIR}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _write_inputs(tmp_path, sc):
    from emba_b200 import eventio

    ev = str(tmp_path / "events.bin")
    st = str(tmp_path / "state.bin")
    eventio.save_events(ev, sc.x, sc.y, sc.t_ns, sc.pol, sc.sensor_w, sc.sensor_h)
    with open(st, "wb") as f:
        f.write(struct.pack("<6i", sc.sensor_w, sc.sensor_h, sc.pano_w, sc.pano_h, sc.n_poses, 0))
        f.write(struct.pack("<7d", sc.C_th, sc.fx, sc.fy, sc.cx, sc.cy, sc.t_beg, sc.dt_knots))
        f.write(np.ascontiguousarray(sc.quat_init, dtype="<f8").tobytes())
        f.write(np.ascontiguousarray(sc.Gx_init, dtype="<f8").tobytes())
        f.write(np.ascontiguousarray(sc.Gy_init, dtype="<f8").tobytes())
    return ev, st


def test_c_consumer_builds_and_fails_loudly_without_gpu(tmp_path, tiny):
    exe = _build(tmp_path)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("CUDA device present: covered by the gpu test")
    ev, st = _write_inputs(tmp_path, tiny)
    r = subprocess.run([exe, ev, st], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_consumer_matches_golden(tmp_path, tiny, tiny_ref):
    exe = _build(tmp_path)
    ev, st = _write_inputs(tmp_path, tiny)
    r = subprocess.run([exe, ev, st], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = {l.split()[0]: l.split()[1:] for l in r.stdout.strip().splitlines()}
    g = tiny_ref
    kv = dict(zip(out["M"][1::2], out["M"][2::2]))  # "M <v> Np <v> cost_data <v> ..."
    assert int(out["M"][0]) == g["ep"].size and int(kv["Np"]) == g["active"].size
    assert abs(float(kv["cost_data"]) - float(g["cost_data"])) < 1e-11 * float(g["cost_data"])
    assert abs(float(kv["cost_reg"]) - float(g["cost_reg"])) < 1e-12 * float(g["cost_reg"])
    assert abs(float(kv["x1_norm"]) - np.linalg.norm(g["x1"])) < 1e-7 * np.linalg.norm(g["x1"])
    assert abs(float(kv["x2_norm"]) - np.linalg.norm(g["x2"])) < 1e-7 * np.linalg.norm(g["x2"])
    lm = out["lm_solves"]
    assert int(lm[0]) == g["lm_log"].shape[0]
    assert abs(float(lm[2]) - float(g["lm_final_cost"])) < 1e-8 * float(g["lm_final_cost"])
    assert [int(v) for v in lm[4:]] == [int(v) for v in g["lm_log"][:, 4]]
    q_last = np.array([float(v) for v in out["q_last"]])
    assert np.max(np.abs(q_last - g["q_final"][-1])) < 1e-8
