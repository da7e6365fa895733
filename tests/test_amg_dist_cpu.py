"""CPU, world_size 2 and 4 over gloo: the DISTRIBUTED host set-up of the smoothed-aggregation hierarchy
(pmgx_amg_setup_dist_h, csrc/amg_setup.cpp) with torch.distributed as the all-gather callback.

Every rank hands the library its owned rows of the oracle's P1 matrix (ghost columns after the owned
ones, the product partitioner's halo lists); the levels are gathered on rank 0 in a global numbering
and checked there: A_{l+1} = P_l^T A_l P_l globally, symmetry, rank-local P that reproduces constants
on interior rows, consistent coarse halo plans, the coarsest inverse, and -- the point of it all -- the
PCG iteration count of the resulting V-cycle (numpy V-cycle of scripts/prototype_sa_amg.py on the
gathered matrices) stays the handful of the single-rank hierarchy."""
import ctypes
import os
import pickle
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PGRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def _worker(rank, world, port, n, out, repl_cap):
    os.environ["PMGX_AMG_REPL_CAP"] = str(repl_cap)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pmg_dolfinx_b200 import api
    from pmg_dolfinx_b200.capi import lib, check, ptr
    import prototype_sa_amg as proto

    Ag, bcg = proto.p1_matrix(n[0]) if n[0] == n[1] == n[2] else (None, None)
    Ag = sp.csr_matrix(Ag)
    mesh = api.BoxMesh(n, PGRID[world], rank)
    s = mesh.space(1)
    no, ng = s.n_owned, s.n_ghost
    g2l = -np.ones(s.n_global, dtype=np.int64)
    g2l[s.l2g] = np.arange(no + ng)
    rows = Ag[s.l2g[:no]].tocoo()
    cols = g2l[rows.col]
    assert (cols >= 0).all(), "an owned row couples to a dof outside owned + ghost"
    Al = sp.csr_matrix((rows.data, (rows.row, cols)), shape=(no, no + ng))
    Al.sort_indices()

    @ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)
    def allgather(_user, mine, nbytes, allp):
        try:
            src = torch.frombuffer((ctypes.c_uint8 * nbytes).from_address(mine), dtype=torch.uint8).clone()
            outs = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(outs, src)
            gathered = torch.cat(outs).numpy()  # keep the array alive across the memmove
            ctypes.memmove(allp, gathered.ctypes.data, nbytes * world)
            return 0
        except Exception as e:  # pragma: no cover
            sys.stderr.write(f"allgather callback failed: {e}\n")
            return 1

    h = ctypes.c_void_p()
    ip, ix = Al.indptr.astype(np.int32), Al.indices.astype(np.int32)
    sr, so, si = (np.ascontiguousarray(a, dtype=np.int32) for a in (s.send_ranks, s.send_offsets, s.send_idx))
    rr, ro, ri = (np.ascontiguousarray(a, dtype=np.int32) for a in (s.recv_ranks, s.recv_offsets, s.recv_idx))
    check(lib.pmgx_amg_setup_dist_h(rank, world, no, ng, ptr(ip), ptr(ix), ptr(Al.data), len(sr), ptr(sr), ptr(so),
                                    ptr(si), len(rr), ptr(rr), ptr(ro), ptr(ri),
                                    ctypes.cast(allgather, ctypes.c_void_p), None, 60, 10, ctypes.addressof(h)))
    levels = []
    for l in range(lib.pmgx_amg_num_levels(h)):
        sz, ds = np.zeros(4, dtype=np.int64), np.zeros(11, dtype=np.int64)
        check(lib.pmgx_amg_level_sizes(h, l, ptr(sz)))
        check(lib.pmgx_amg_level_dist_sizes(h, l, ptr(ds)))
        nrow, nnz, pc, pnnz = (int(v) for v in sz)
        lo, lg, nsn, ns, nrn, nr, dense, nglob, rnnz, repl, n_mine = (int(v) for v in ds)
        assert lo == nrow
        ap, ac, av = np.zeros(nrow + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        pp, pcl, pv = np.zeros(nrow + 1, np.int32), np.zeros(pnnz, np.int32), np.zeros(pnnz)
        lmax = ctypes.c_double()
        check(lib.pmgx_amg_level_get(h, l, ptr(ap), ptr(ac), ptr(av), ptr(pp) if pc else None,
                                     ptr(pcl) if pc else None, ptr(pv) if pc else None, ctypes.addressof(lmax)))
        gs, gr = np.zeros(lg, np.int32), np.zeros(lg, np.int32)
        psr, pso, psi = np.zeros(nsn, np.int32), np.zeros(nsn + 1, np.int32), np.zeros(ns, np.int32)
        prr, pro, pri = np.zeros(nrn, np.int32), np.zeros(nrn + 1, np.int32), np.zeros(nr, np.int32)
        inv = np.zeros(lo * nglob if dense else 0)
        check(lib.pmgx_amg_level_dist_get(h, l, ptr(gs), ptr(gr), ptr(psr), ptr(pso), ptr(psi), ptr(prr), ptr(pro),
                                          ptr(pri), ptr(inv) if dense else None))
        rp, rc, rv = np.zeros((pc + 1) if pc else 0, np.int32), np.zeros(rnnz, np.int32), np.zeros(rnnz)
        if pc:
            check(lib.pmgx_amg_level_get_restriction(h, l, ptr(rp), ptr(rc), ptr(rv)))
        levels.append(dict(n_owned=lo, n_ghost=lg, A=(ap, ac, av), P=(pp, pcl, pv, pc), R=(rp, rc, rv), lmax=lmax.value,
                           ghost_src=gs, ghost_rid=gr, send=(psr, pso, psi), recv=(prr, pro, pri), dense=dense,
                           n_global=nglob, inv=inv, replicated=bool(repl), n_mine=n_mine))
    check(lib.pmgx_amg_destroy(h))
    objs = [None] * world if rank == 0 else None
    dist.gather_object(dict(levels=levels, l2g=s.l2g[:no]), objs, dst=0)
    if rank == 0:
        with open(out, "wb") as f:
            pickle.dump(objs, f)
    dist.barrier()
    dist.destroy_process_group()


def _gcols(per_rank, l):
    """per rank: local column (owned + ghost) -> global id of level l, and the row offsets"""
    if per_rank[0]["levels"][l]["replicated"]:
        # the whole level on every rank in the canonical numbering; rows this rank's restriction produces: n_mine
        ng = per_rank[0]["levels"][l]["n_owned"]
        first = not per_rank[0]["levels"][l - 1]["replicated"]
        no = [r["levels"][l]["n_mine"] for r in per_rank] if first else [ng] + [0] * (len(per_rank) - 1)
        return [np.arange(ng, dtype=np.int64) for _ in per_rank], np.concatenate([[0], np.cumsum(no)]), no
    no = [r["levels"][l]["n_owned"] for r in per_rank]
    off = np.concatenate([[0], np.cumsum(no)])
    out = []
    for q, r in enumerate(per_rank):
        L = r["levels"][l]
        g = np.empty(no[q] + L["n_ghost"], dtype=np.int64)
        g[: no[q]] = off[q] + np.arange(no[q])
        if L["n_ghost"]:
            g[no[q]:] = off[L["ghost_src"]] + L["ghost_rid"]
        out.append(g)
    return out, off, no


def _assemble(per_rank, l):
    """Global A_l, P_l, R_l and row offsets of level l from the ranks' pieces."""
    gc, off, no = _gcols(per_rank, l)
    N = int(off[-1])
    if per_rank[0]["levels"][l]["replicated"]:
        # identical copies on every rank; rank 0's stands for the level
        mats = []
        for r in per_rank:
            ap, ac, av = r["levels"][l]["A"]
            mats.append(sp.csr_matrix((av, ac, ap), shape=(N, N)))
        assert all(abs(m - mats[0]).max() == 0 for m in mats[1:])
        A = mats[0]
        P = R = None
        pp, pcl, pv, pc = per_rank[0]["levels"][l]["P"]
        if pc:
            P = sp.csr_matrix((pv, pcl, pp), shape=(N, pc))
            rp, rc, rv = per_rank[0]["levels"][l]["R"]
            R = sp.csr_matrix((rv, rc, rp), shape=(pc, N))
        return A, P, off, R
    Ar, Ac_, Av = [], [], []
    for q, r in enumerate(per_rank):
        ap, ac, av = r["levels"][l]["A"]
        Ar.append(off[q] + np.repeat(np.arange(no[q]), np.diff(ap)))
        Ac_.append(gc[q][ac])
        Av.append(av)
    A = sp.csr_matrix((np.concatenate(Av), (np.concatenate(Ar), np.concatenate(Ac_))), shape=(N, N))
    P = R = None
    if any(r["levels"][l]["P"][3] for r in per_rank):
        gcc, coff, nc = _gcols(per_rank, l + 1)
        Pr, Pc, Pv, Rr, Rc, Rv = [], [], [], [], [], []
        for q, r in enumerate(per_rank):
            pp, pcl, pv, pc = r["levels"][l]["P"]
            if r["levels"][l + 1]["replicated"]:
                assert pc == r["levels"][l + 1]["n_owned"]           # ... or the canonical numbering of a replicated level
            else:
                assert pc == nc[q] + r["levels"][l + 1]["n_ghost"]   # P's columns: the next level's owned + ghost dofs
            Pr.append(off[q] + np.repeat(np.arange(no[q]), np.diff(pp)))
            Pc.append(gcc[q][pcl])
            Pv.append(pv)
            rp, rc, rv = r["levels"][l]["R"]
            Rr.append(coff[q] + np.repeat(np.arange(nc[q]), np.diff(rp[: nc[q] + 1])))
            Rc.append(gc[q][rc])
            Rv.append(rv)
        P = sp.csr_matrix((np.concatenate(Pv), (np.concatenate(Pr), np.concatenate(Pc))), shape=(N, int(coff[-1])))
        R = sp.csr_matrix((np.concatenate(Rv), (np.concatenate(Rr), np.concatenate(Rc))), shape=(int(coff[-1]), N))
    return A, P, off, R


@pytest.mark.parametrize("repl_cap", [0, 65536], ids=["distributed-to-the-bottom", "small-levels-replicated"])
@pytest.mark.parametrize("world", [2, 4])
def test_distributed_amg_setup_over_gloo(world, repl_cap, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import prototype_sa_amg as proto
    n = (12, 12, 12)
    out = str(tmp_path / "amg.pkl")
    mp.spawn(_worker, args=(world, 29720 + world + (10 if repl_cap else 0), n, out, repl_cap), nprocs=world, join=True)
    per_rank = pickle.load(open(out, "rb"))
    nl = len(per_rank[0]["levels"])
    assert nl >= 2 and all(len(r["levels"]) == nl for r in per_rank)
    assert any(L["replicated"] for L in per_rank[0]["levels"]) == bool(repl_cap)
    assert not per_rank[0]["levels"][0]["replicated"]
    Ag, bcg = proto.p1_matrix(n[0])
    # level 0 in the gathered numbering is a permutation of the oracle matrix
    A0, P0, off0, _ = _assemble(per_rank, 0)
    perm = np.concatenate([r["l2g"] for r in per_rank])
    assert abs(A0 - sp.csr_matrix(Ag)[perm][:, perm]).max() < 1e-14
    levels = []
    for l in range(nl):
        A, P, off, R = _assemble(per_rank, l)
        assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
        lev = dict(A=A, dinv=1.0 / A.diagonal(), lmax=per_rank[0]["levels"][l]["lmax"])
        # every rank holds the same (global) eigenvalue estimate, an upper bound of the true one
        assert all(abs(r["levels"][l]["lmax"] - lev["lmax"]) <= 1e-12 * lev["lmax"] for r in per_rank)
        lam = proto.lam_max(A, lev["dinv"], its=60)
        assert 0.95 * lam <= lev["lmax"] <= 1.4 * lam
        if P is not None:
            lev["P"] = P
            G = (P.T @ A @ P).tocsr()
            An, _, _, _ = _assemble(per_rank, l + 1)
            assert abs(G - An).max() <= 1e-11 * abs(G).max()                 # distributed Galerkin product
            assert abs(R - P.T).max() <= 1e-15                               # R = the owned rows of the global P^T
            # the prolongator is smoothed ACROSS the partition interfaces: it equals the one a single rank
            # would build from the same aggregates, (I - omega D^-1 A) T with T = its unit-entry pattern
            ranks_of_row = np.searchsorted(off, np.arange(A.shape[0]), side="right") - 1
            coff = np.concatenate([[0], np.cumsum([r["levels"][l + 1]["n_owned"] for r in per_rank])])
            ranks_of_col = np.searchsorted(coff, np.arange(P.shape[1]), side="right") - 1
            Pc = P.tocoo()
            if not per_rank[0]["levels"][l]["replicated"] and not per_rank[0]["levels"][l + 1]["replicated"]:
                assert (ranks_of_row[Pc.row] != ranks_of_col[Pc.col]).any()   # P does reach across ranks
            free = np.diff(A.indptr) > 1
            interior = np.abs(A @ np.ones(A.shape[0])) <= 1e-12 * A.diagonal()
            rs = np.asarray(P.sum(axis=1)).ravel()
            assert np.allclose(rs[free & interior], 1.0, atol=1e-12)         # constants are reproduced
            assert not (np.diff(P.indptr)[~free] > 0).any()                  # Dirichlet rows stay out
        # halo plans: what q sends to d is exactly what d expects from q, in order
        for q, r in enumerate(per_rank):
            if r["levels"][l]["replicated"]:
                continue
            L = r["levels"][l]
            sr, so, si = L["send"]
            for k, d in enumerate(sr):
                Ld = per_rank[d]["levels"][l]
                rr, ro, ri = Ld["recv"]
                kk = list(rr).index(q)
                slots = ri[ro[kk]:ro[kk + 1]]
                assert len(slots) == so[k + 1] - so[k]
                assert (Ld["ghost_src"][slots] == q).all()
                assert np.array_equal(Ld["ghost_rid"][slots], si[so[k]:so[k + 1]])
            # every ghost is covered by exactly one receive slot
            assert sorted(L["recv"][2].tolist()) == list(range(L["n_ghost"]))
        levels.append(lev)
    # coarsest level: the ranks' rows of the dense inverse
    last = [r["levels"][-1] for r in per_rank]
    assert all(L["dense"] for L in last)
    Ac, _, off, _ = _assemble(per_rank, nl - 1)
    N = Ac.shape[0]
    inv = np.zeros((N, N))
    if last[0]["replicated"]:
        assert all(np.array_equal(L["inv"], last[0]["inv"]) for L in last)
        inv = last[0]["inv"].reshape(N, N)
    else:
        for q, L in enumerate(last):
            no = L["n_owned"]
            loc2glob = np.concatenate([np.arange(off[q], off[q + 1])] + [np.arange(off[p], off[p + 1]) for p in range(world) if p != q])
            inv[off[q]:off[q + 1], loc2glob] = L["inv"].reshape(no, N)
    assert np.allclose(inv @ Ac.toarray(), np.eye(N), atol=1e-9)
    levels[-1]["dense"] = inv
    # the hierarchy does its job: a handful of PCG iterations, like the single-rank one
    free0 = np.diff(A0.indptr) > 1
    b = np.random.default_rng(1).uniform(-1, 1, A0.shape[0]) * free0
    _, k_amg = proto.pcg(A0, b, lambda r: proto.vcycle(levels, 0, r), 1e-5, 100)
    _, k_jac = proto.pcg(A0, b, lambda r: levels[0]["dinv"] * r, 1e-5, 2000)
    # ... and like a hierarchy that ignores the partition altogether (the single-rank set-up of the same matrix)
    from test_amg_setup import _hierarchy
    single = _hierarchy(Ag, min_coarse=60)
    b_can = np.zeros_like(b)
    b_can[perm] = b
    _, k_single = proto.pcg(Ag, b_can, lambda r: proto.vcycle(single, 0, r), 1e-5, 100)
    assert k_amg <= k_single + 1 and k_amg * 3 < k_jac, (k_amg, k_single, k_jac)
