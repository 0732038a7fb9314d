"""world_size-2 gloo test (CPU) of the time-sharded decomposition the multi-GPU path uses:

  * events are partitioned into contiguous slices of whole batches, a measurement belongs to the rank of its CURRENT
    event; a pair whose previous event lies in an earlier slice is closed through all-gathered per-sensor-pixel
    "last event" tables (halo), the reference-order rank of every pair through all-gathered per-pixel pair counts;
    num_ev_map is all-reduced BEFORE the active-pixel decision, the cost is all-reduced;
  * A11/b1/A22/b2 partials are all-reduced;
  * A12 is exchanged as per-pixel sub-strips: every rank sends, for the pixels another rank owns, only its own
    (pose-window) sub-strip; the owner merges them -- no rank ever holds all of A12; the receive layout of the
    peer-memory exchange (every rank derives every owner's layout from the all-gathered windows) is restated too;
  * Schur: owners' partial sums all-reduced, replicated factorisation, x2 from the owners;
  * PCG: replicated vectors, partial matrix-vector products combined by one all-reduce per iteration.

Each rank computes its partial quantities with the numpy oracle restricted to its slice; the combined result must
equal the unsharded oracle (and therefore the reference golden vectors). This exercises the same collectives the
CUDA library issues through NCCL, on gloo.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, out):
    from conftest import GoldenScene, load_golden_ref, rel
    from oracle import emba_oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, g = GoldenScene("tiny"), load_golden_ref("tiny")
    orc = O.Oracle(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.pano_w, sc.pano_h, sc.C_th)
    orc.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    t0, dt = O.spline_base_ns(sc.t_beg, sc.dt_knots)
    n, thres, alpha = sc.n_poses, int(g["thres"]), float(g["alpha"])
    # ---- (0) the per-rank pairing pre-pass (emba_b200/csrc/prepass.cu, prepass_device with several ranks), restated:
    # sort my slice by sensor pixel, exchange the "last event of the pixel" tables, close the halo pairs, exchange the
    # per-pixel pair counts, derive the reference-order rank of every pair
    n_used = orc.n_used
    B = n_used // 100
    e_lo, e_hi = 100 * (B * rank // world), 100 * (B * (rank + 1) // world)
    S = sc.sensor_w * sc.sensor_h
    spix = (sc.y[e_lo:e_hi].astype(np.int64) * sc.sensor_w + sc.x[e_lo:e_hi].astype(np.int64))
    o = np.argsort(spix, kind="stable")
    ks, vs = spix[o], o + e_lo
    first = np.r_[True, ks[1:] != ks[:-1]] if ks.size else np.zeros(0, bool)
    lastm = np.r_[ks[1:] != ks[:-1], True] if ks.size else np.zeros(0, bool)
    last = -np.ones(S, dtype=np.int64)
    last[ks[lastm]] = vs[lastm]
    tabs = [torch.zeros(S, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(tabs, torch.from_numpy(last))
    halo = -np.ones(S, dtype=np.int64)
    for r in range(rank - 1, -1, -1):
        t = tabs[r].numpy()
        halo = np.where((halo < 0) & (t >= 0), t, halo)
    prev_loc = np.where(first, halo[ks], np.r_[-1, vs[:-1]])  # global ids, -1 = no pair
    is_pair = prev_loc >= 0
    cnt = np.bincount(ks[is_pair], minlength=S)
    cnts = [torch.zeros(S, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(cnts, torch.from_numpy(cnt))
    tot = sum(c.numpy() for c in cnts)
    lower = sum((cnts[r].numpy() for r in range(rank)), np.zeros(S, dtype=np.int64))
    pixbase = np.cumsum(tot) - tot
    within = np.cumsum(is_pair) - is_pair  # pairs before position j in my sorted slice
    firstpos = np.zeros(S, dtype=np.int64)
    firstpos[ks[first]] = np.nonzero(first)[0]
    refrank = pixbase[ks] + lower[ks] + (within - within[firstpos[ks]])
    # against the unsharded pairing of the oracle (reference order = sensor pixel row-major, then time)
    gcur, gprev = orc.cur, orc.prev
    sel = (gcur >= e_lo) & (gcur < e_hi)
    assert int(tot.sum()) == gcur.size and int(is_pair.sum()) == int(sel.sum())
    mycur, myprev, myrank = vs[is_pair], prev_loc[is_pair], refrank[is_pair]
    assert np.array_equal(gcur[myrank], mycur) and np.array_equal(gprev[myrank], myprev)
    # ---- full evaluation, then restrict to this rank's slice: a measurement belongs to the rank of its current event
    orc.evaluate(sc.quat_init, t0, dt, sc.Gx_init, sc.Gy_init, True)
    st = orc.state
    M = st["cur"].size
    mine = np.nonzero((st["cur"] >= e_lo) & (st["cur"] < e_hi))[0]
    # (1) histogram + cost all-reduce
    W, H = sc.pano_w, sc.pano_h
    hist = torch.from_numpy(np.bincount(st["pix"][mine], minlength=W * H).astype(np.int32))
    dist.all_reduce(hist)
    cost = torch.tensor([0.5 * float(st["ep"][mine] @ st["ep"][mine]), float(mine.size)], dtype=torch.float64)
    dist.all_reduce(cost)
    assert np.array_equal(hist.numpy().reshape(H, W), g["num_ev_map"])
    assert abs(cost[0].item() - float(g["cost_data"])) < 1e-11 * float(g["cost_data"]) and int(cost[1]) == g["ep"].size
    # (2) partial normal equations on the slice with the GLOBAL active set
    sub = {k: (v[mine] if isinstance(v, np.ndarray) and v.shape[:1] == (M,) else v) for k, v in st.items()}
    sub["num"] = hist.numpy().reshape(H, W)
    orc.state = sub
    A11, A12, A22, b1, b2, act = orc.form_normal_eq(n, thres)
    assert np.array_equal(act, g["active"])
    Np = act.size
    for arr in (A11, A22, b1, b2):
        t = torch.from_numpy(arr)
        dist.all_reduce(t)
    if rank == 0:  # the regulariser is added once
        A22r, b2r = orc.apply_l2_reg(A22 * 0, b2 * 0, act, alpha, sc.Gx_init, sc.Gy_init)
    else:
        A22r, b2r = A22 * 0, b2 * 0
    ta, tb = torch.from_numpy(A22r), torch.from_numpy(b2r)
    dist.all_reduce(ta)
    dist.all_reduce(tb)
    A22, b2 = A22 + ta.numpy(), b2 + tb.numpy()
    assert rel(g["A11"], A11) < 1e-11 and rel(g["b1"], b1) < 1e-11
    assert rel(g["A22"], A22) < 1e-11 and rel(g["b2"], b2) < 1e-11
    # (3) A12: local pose windows per pixel, sub-strips to the pixel owners, merge
    A12p = A12.reshape(3 * n, Np, 2)
    touched = np.abs(A12p).reshape(n, 3, Np, 2).sum((1, 3)) > 0  # [pose, pixel]
    lo = np.where(touched.any(0), touched.argmax(0), n)
    hi = np.where(touched.any(0), n - 1 - touched[::-1].argmax(0), -1)
    own = lambda q: (Np * q // world, Np * (q + 1) // world)
    send = []
    for q in range(world):
        a0, a1 = own(q)
        send.append([(int(lo[a]), A12p[3 * lo[a]: 3 * (hi[a] + 1), a, :].copy()) if hi[a] >= lo[a] else (n, None)
                     for a in range(a0, a1)])
    gathered = [None] * world
    dist.all_gather_object(gathered, send)  # gloo has no all_to_all: every rank picks the chunks addressed to it
    a0, a1 = own(rank)
    owned = np.zeros((3 * n, a1 - a0, 2))
    for src in range(world):  # merge in rank order (deterministic)
        for i, (l, blk) in enumerate(gathered[src][rank]):
            if blk is not None:
                owned[3 * l: 3 * l + blk.shape[0], i, :] += blk
    # (3b) the same exchange in the layout of the peer-memory path (emba_b200/csrc/comm.cu): windows all-gathered, every
    # rank derives the SAME [source][owner] volume matrix and from it every owner's receive layout; a sender's
    # sub-strip of pixel a goes to  base[owner][me] + stripoff_me[a] - stripoff_me[first pixel of the owner]  (what
    # k_strip_dst computes and k_pix stores to); the owner's merge finds source s's sub-strip of its i-th pixel at the
    # flattened exclusive scan of the [source][pixel] lengths (what k_own_len + scan_exclusive compute)
    win = torch.from_numpy(np.stack([lo, hi], 1).astype(np.int64))
    wins = [torch.zeros_like(win) for _ in range(world)]
    dist.all_gather(wins, win)
    wins = [w.numpy() for w in wins]
    lens = [np.where(w[:, 1] >= w[:, 0], w[:, 1] - w[:, 0] + 1, 0) for w in wins]           # poses per (source, pixel)
    owner_of = ((np.arange(Np) + 1) * world - 1) // Np                                      # comm.cu: owner_of()
    for q in range(world):
        qa0, qa1 = own(q)
        assert np.all(owner_of[qa0:qa1] == q)
    cnt = np.array([[lens[s_][owner_of == q].sum() for q in range(world)] for s_ in range(world)])  # k_cnt_matrix
    base = np.vstack([np.zeros(world, dtype=np.int64), np.cumsum(cnt, 0)[:-1]])             # base[s][q]: poses before s's chunk
    stripoff = np.concatenate([[0], np.cumsum(lens[rank])])                                  # my local strip offsets
    stores = [[] for _ in range(world)]                                                     # what k_pix would store, per owner
    for a in range(Np):
        if lens[rank][a] == 0:
            continue
        q = owner_of[a]
        dst = base[rank][q] + stripoff[a] - stripoff[own(q)[0]]
        stores[q].append((int(dst), A12p[3 * lo[a]: 3 * (hi[a] + 1), a, :].copy()))
    got = [None] * world
    dist.all_gather_object(got, stores)
    recv = np.full((int(cnt[:, rank].sum()) * 3, 2), np.nan)                                # my receive buffer, 3 rows per pose
    for src in range(world):
        for dst, blk in got[src][rank]:
            assert np.all(np.isnan(recv[3 * dst: 3 * dst + blk.shape[0]]))                  # nobody else wrote here
            recv[3 * dst: 3 * dst + blk.shape[0]] = blk
    assert not np.isnan(recv).any()                                                         # ... and every slot was written
    own_len = np.stack([np.r_[lens[s_][a0:a1], 0] for s_ in range(world)])                  # [source][n_own + 1], tail 0
    own_off = (np.cumsum(own_len.reshape(-1)) - own_len.reshape(-1)).reshape(world, -1)     # flattened exclusive scan
    owned_peer = np.zeros_like(owned)
    for i in range(a1 - a0):
        for s_ in range(world):                                                             # rank order: deterministic
            L = own_len[s_, i]
            if L:
                l = wins[s_][a0 + i, 0]
                owned_peer[3 * l: 3 * (l + L), i, :] += recv[3 * own_off[s_, i]: 3 * (own_off[s_, i] + L)]
    assert np.array_equal(owned_peer, owned)
    # pose windows of different time slices overlap only at slice boundaries
    full = np.zeros((3 * n, Np, 2))
    full[:, a0:a1, :] = owned
    t = torch.from_numpy(full)
    dist.all_reduce(t)  # only to compare against the golden sums
    A12full = t.numpy().reshape(3 * n, 2 * Np)
    assert rel(g["A12_rowsum"], A12full.sum(1)) < 1e-11 and rel(g["A12_colsum"], A12full.sum(0)) < 1e-11
    assert abs(np.linalg.norm(A12full) - float(g["A12_fro"])) < 1e-11 * float(g["A12_fro"])
    # (4) Schur complement from the owners' partial sums, replicated solve, x2 from the owners
    lam = float(g["lam"])
    A11m = A11 + lam * np.diag(np.diag(A11))
    A22m = A22.copy()
    A22m[:, 0, 0] += lam * A22[:, 0, 0]
    A22m[:, 1, 1] += lam * A22[:, 1, 1]
    Cinv = np.linalg.inv(A22m)
    Wo = np.einsum("rpi,pij->rpj", owned, Cinv[a0:a1])
    Sp = torch.from_numpy(np.einsum("rpj,spj->rs", Wo, owned))
    rp = torch.from_numpy(np.einsum("rpj,pj->r", Wo, b2.reshape(Np, 2)[a0:a1]))
    dist.all_reduce(Sp)
    dist.all_reduce(rp)
    S = (A11m - Sp.numpy())[3:, 3:]
    x1 = np.linalg.solve(S, (b1 - rp.numpy())[3:])
    x1f = np.concatenate([np.zeros(3), x1])
    x2 = np.zeros((Np, 2))
    x2[a0:a1] = np.einsum("pij,pj->pi", Cinv[a0:a1], b2.reshape(Np, 2)[a0:a1] - np.einsum("rpj,r->pj", owned, x1f))
    t2 = torch.from_numpy(x2)
    dist.all_reduce(t2)
    assert rel(g["x1"], x1) < 1e-8 and rel(g["x2"], t2.numpy().reshape(-1)) < 1e-8
    # (5) Jacobi-PCG with the pixel-sharded matrix (solve_pcg with several GPUs): vectors replicated, rank 0 adds
    # the A11m block, pixel owners add their A22m blocks and strip products, ONE all-reduce per product; the scalars
    # need no communication. Same iteration count and solution as Eigen's ConjugateGradient in the reference.
    d = 3 * (n - 1)
    own_g = owned[3:]                     # gauge: first pose fixed
    A11g = A11m[3:, 3:]
    bvec = np.concatenate([b1[3:], b2])
    dinv = 1.0 / np.concatenate([np.diag(A11g), np.stack([A22m[:, 0, 0], A22m[:, 1, 1]], 1).reshape(-1)])

    def matvec(v):
        v1, v2 = v[:d], v[d:].reshape(Np, 2)
        y = np.zeros(d + 2 * Np)
        if rank == 0:
            y[:d] += A11g @ v1
        y[:d] += np.einsum("rpj,pj->r", own_g, v2[a0:a1])
        y2 = np.zeros((Np, 2))
        y2[a0:a1] = np.einsum("pij,pj->pi", A22m[a0:a1], v2[a0:a1]) + np.einsum("rpj,r->pj", own_g, v1)
        y[d:] = y2.reshape(-1)
        t = torch.from_numpy(y)
        dist.all_reduce(t)
        return t.numpy()

    x = np.zeros(d + 2 * Np)
    r = bvec.copy()
    rhs2 = float(bvec @ bvec)
    thr = max(1e-12 * rhs2, np.finfo(np.float64).tiny)  # tolerance 1e-6 squared (model.cpp:828-831)
    p_ = dinv * r
    abs_new = float(r @ p_)
    it = 0
    while it < 100:
        tmp = matvec(p_)
        alpha_ = abs_new / float(p_ @ tmp)
        x += alpha_ * p_
        r -= alpha_ * tmp
        if float(r @ r) < thr:
            break
        z = dinv * r
        abs_old, abs_new = abs_new, float(r @ z)
        p_ = z + (abs_new / abs_old) * p_
        it += 1
    assert it == int(g["cg_iters"]), (it, int(g["cg_iters"]))
    assert rel(g["x1_cg"], x[:d]) < 1e-6 and rel(g["x2_cg"], x[d:]) < 1e-6
    out[rank] = 1
    dist.barrier()
    dist.destroy_process_group()


def test_time_sharded_decomposition_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29611, out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]
