"""torchrun script: per-phase device timings of the time-sharded pass (and LM) on a workload. usage (N ranks):
python -m torch.distributed.run --nproc-per-node N tools/mgpu_phases.py C4 [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from emba_b200 import synth
from emba_b200.legm import Engine, spline_base_ns
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
sc = synth.make_config(name, device=f"cuda:{lr}")
eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h, device=lr)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(uid, 0)
eng.comm_init(uid.cpu().numpy().tobytes(), rank, world)
eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
t0, dt = spline_base_ns(sc.t_beg, sc.dt_knots)
eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
rows = []
for i in range(reps + 3):
    dist.barrier()
    eng.evaluate(0, 0, 1.0, 5.0); te = eng.timings_ms(); ce = eng.comm_ms()["eval_allreduce"]
    eng.form_normal_eq(5, 0, 1.0, 5.0); tf = eng.timings_ms(); cf = eng.comm_ms()
    eng.solve(1e-3, False, True, want=False); ts = eng.timings_ms()
    if i >= 3:
        rows.append([te["evaluate"], te["eval_kernel"], ce, tf["form"], tf["asm_pose_kernel"], tf["pix_kernel"], tf["sort"],
                     cf["exchange_prepare"], cf["pix_and_sends"], cf["allreduce_and_merge"], ts["solve"]])
m = torch.tensor(np.mean(rows, 0), device="cuda")
allm = [torch.zeros_like(m) for _ in range(world)]
dist.all_gather(allm, m)
if rank == 0:
    A = torch.stack(allm).cpu().numpy()
    names = ["evaluate", "k_eval", "eval_allreduce", "form", "k_asm_pose", "k_pix(+sends)", "sort_side", "xchg_prepare",
             "pix_and_sends", "allreduce+merge", "solve"]
    print(f"{name} world={world} strips via {cf['strip_exchange']} pipeline={os.environ.get('EMBA_XCHG_PIPELINE', '0')} N={sc.n_events}  pass(max over ranks) = "
          f"{(A[:, 0] + A[:, 3]).max():.3f} ms")
    for j, nm in enumerate(names):
        print(f"  {nm:18s} mean {A[:, j].mean():7.3f}  max {A[:, j].max():7.3f}  per rank {np.round(A[:, j], 2).tolist()}")
eng.close()
dist.barrier()
dist.destroy_process_group()
