"""Run under torchrun (one rank per GPU): time-sharded evaluation / assembly / LM with NCCL all-reduce of the partial
systems must reproduce the single-GPU golden results. Launched by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GoldenScene, load_golden_ref, rel  # noqa: E402
from emba_b200.legm import Engine, spline_base_ns  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    for name in ("tiny", "small"):
        sc, ref = GoldenScene(name), load_golden_ref(name)
        eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h, device=lr)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        eng.comm_init(uid.cpu().numpy().tobytes(), rank, world)
        eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
        t0, dt = spline_base_ns(sc.t_beg, sc.dt_knots)
        eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
        cd, cr, M = eng.evaluate(0, 0, 1.0, 5.0)
        _, num = eng.get_evaluation(0, None, False, True)
        assert M == ref["ep"].size, (M, ref["ep"].size)
        assert np.array_equal(num, ref["num_ev_map"])
        assert abs(cd - float(ref["cost_data"])) < 1e-11 * float(ref["cost_data"])
        Np = eng.form_normal_eq(5, 0, 1.0, 5.0)
        mode = eng.comm_ms()["strip_exchange"]
        want = os.environ.get("EMBA_EXPECT_EXCHANGE")
        assert want is None or mode == want, (mode, want)
        A11, A12, A22, b1, b2, act = eng.get_normal_eq(True)
        assert np.array_equal(act, ref["active"])
        # A12 is pixel-sharded: every rank holds the complete columns of the pixels it owns and zeros elsewhere
        sums = torch.from_numpy(np.concatenate([A12.sum(1), A12.sum(0), [np.sum(A12 * A12)]])).cuda()
        dist.all_reduce(sums)
        sums = sums.cpu().numpy()
        d3 = A12.shape[0]
        for a, b in ((ref["A11"], A11), (ref["b1"], b1), (ref["A22"], A22), (ref["b2"], b2),
                     (ref["A12_rowsum"], sums[:d3]), (ref["A12_colsum"], sums[d3:-1]),
                     (float(ref["A12_fro"]), np.sqrt(sums[-1]))):
            assert rel(a, b) < 1e-9, rel(a, b)
        own = np.nonzero(np.abs(A12).sum(0))[0] // 2
        assert own.size == 0 or (own.min() >= Np * rank // world and own.max() < Np * (rank + 1) // world)
        x1, x2, _, _ = eng.solve(1e-3, False, True)
        assert rel(ref["x1"], x1) < 1e-7 and rel(ref["x2"], x2) < 1e-7
        # PCG with the pixel-sharded A12 (one all-reduce of the product vector per iteration): same iteration count
        # and solution as the reference's Eigen::ConjugateGradient on one process
        y1, y2, it, err = eng.solve(1e-3, True, True)
        assert it == int(ref["cg_iters"]), (it, int(ref["cg_iters"]))
        assert rel(ref["x1_cg"], y1) < 1e-6 and rel(ref["x2_cg"], y2) < 1e-6
        eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
        log, fcost = eng.solve_time_window(alpha=5.0, thres=5)
        rlog = ref["lm_log"]
        assert log.shape[0] == rlog.shape[0] and np.array_equal(log[:, 4], rlog[:, 4])
        q, gx, gy = eng.get_state(0)
        ang = 2 * np.arccos(np.abs(np.sum(q * ref["q_final"], -1)).clip(0, 1))
        assert np.max(ang) < 1e-5 and rel(ref["Gx_final"], gx) < 1e-4
        eng.close()
        if rank == 0:
            print(f"mgpu_check {name}: world={world} OK (M={M}, Np={Np}, LM solves={log.shape[0]}, strips via {mode})", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
