"""Data-parallel host logic on CPU: the nodal partition / local numbering / halo lists of dedflow_b200.dist, exercised
with world_size 2 over gloo.  Each rank assembles its owned rows with the ORACLE on its local mesh and runs a numpy
restatement of the distributed GMRES (halo exchange + allreduce); the gathered answer must equal the single-domain
oracle solve.  (The CUDA version of the same flow is checked on 2 GPUs by scripts/dist_check.py.)"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from dedflow_b200 import boxmesh  # noqa: E402
from dedflow_b200 import dist as ddist  # noqa: E402

BCS = {0: (1, 1, 1), 2: (0, 1, 0), 3: (0, 0, 1), 4: (0, 0, 0)}


def test_partition_invariants_world3():
    mesh = boxmesh.make_box(6)
    npart = ddist.slab_owner(mesh, 3)
    assert set(np.unique(npart)) == {0, 1, 2}
    lms = [ddist.partition(mesh, npart, r, 3) for r in range(3)]
    # every node owned exactly once; every element local on the ranks owning one of its nodes
    assert sum(lm.n_own for lm in lms) == mesh.num_node
    owned = np.concatenate([lm.nodes_g[:lm.n_own] for lm in lms])
    assert np.array_equal(np.sort(owned), np.arange(mesh.num_node))
    for lm in lms:
        assert np.all(npart[lm.nodes_g[:lm.n_own]] == lm.rank) and np.all(npart[lm.nodes_g[lm.n_own:]] != lm.rank)
        assert np.array_equal(lm.nodes_g[lm.ien], mesh.ien[lm.elems_g])                 # local connectivity maps back
        assert np.array_equal(lm.xg, mesh.xg[lm.nodes_g])
        # interior rows reference no ghost
        touched = np.zeros(lm.num_node, bool)
        ghost_el = (lm.ien >= lm.n_own).any(axis=1)
        touched[np.unique(lm.ien[ghost_el])] = True
        assert not touched[:lm.n_interior].any() and touched[lm.n_interior:lm.n_own].all()
        # halo lists are symmetric: what r sends to q is what q expects from r, in the same (global id) order
        for qi, q in enumerate(lm.neighbors):
            snd = lm.nodes_g[lm.send_nodes[lm.send_offset[qi]:lm.send_offset[qi + 1]]]
            other = lms[q]
            ri = list(other.neighbors).index(lm.rank)
            rcv = other.nodes_g[other.recv_nodes[other.recv_offset[ri]:other.recv_offset[ri + 1]]]
            assert np.array_equal(snd, rcv) and np.all(np.diff(snd) > 0)
        assert lm.recv_offset[-1] == lm.num_node - lm.n_own
    # localize / scatter_owned round trip
    v = np.random.default_rng(0).standard_normal(6 * mesh.num_node)
    out = np.zeros_like(v)
    for lm in lms:
        lm.scatter_owned(lm.localize(v), out)
    assert np.array_equal(out, v)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _halo(lm, x):
    """refresh the ghost (u,p) entries of a 6N_loc-layout numpy vector over gloo"""
    n = lm.num_node
    reqs, bufs = [], []
    for qi, q in enumerate(lm.neighbors):
        sn = lm.send_nodes[lm.send_offset[qi]:lm.send_offset[qi + 1]]
        rn = lm.recv_nodes[lm.recv_offset[qi]:lm.recv_offset[qi + 1]]
        sb = torch.from_numpy(np.concatenate([x[:3 * n].reshape(n, 3)[sn], x[3 * n + sn][:, None]], axis=1).copy())
        rb = torch.zeros(rn.size, 4, dtype=torch.float64)
        reqs.append(dist.isend(sb, int(q)))
        reqs.append(dist.irecv(rb, int(q)))
        bufs.append((rn, rb))
    for r in reqs:
        r.wait()
    for rn, rb in bufs:
        x[:3 * n].reshape(n, 3)[rn] = rb[:, :3].numpy()
        x[3 * n + rn] = rb[:, 3].numpy()


def _make_mesh(m):
    """m > 0: Kuhn box split into z-slabs; m == 0: the unstructured Delaunay cube split by coordinate bisection"""
    if m > 0:
        return boxmesh.make_box(m), ddist.slab_owner
    return boxmesh.delaunay_cube(), ddist.coordinate_owner


def _worker(rank, world, port, m, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    O = pyoracle.get()
    mesh, owner_fn = _make_mesh(m)
    Ng = mesh.num_node
    lm = ddist.partition(mesh, owner_fn(mesh, world), rank, world)
    wg_g, dwg_g = boxmesh.state_random(Ng)
    wg, dwg = lm.localize(wg_g), lm.localize(dwg_g)
    N, no = lm.num_node, lm.n_own
    # ---- local assembly with the oracle: all local elements into the local pattern; owned rows are complete ----
    rp, ci = O.nodal_pattern(N, lm.ien)
    Z = ci.size
    color = np.zeros(lm.num_tet, np.int32)                       # one "batch" in element order (oracle is sequential-safe)
    off, ind = np.array([0, lm.num_tet], np.int32), np.arange(lm.num_tet, dtype=np.int32)
    os.environ["OMP_NUM_THREADS"] = "1"
    F = np.zeros(6 * N)
    blocks = [np.zeros(9 * Z), np.zeros(3 * Z), np.zeros(3 * Z), np.zeros(Z)]
    import ctypes
    O.L.orc_assemble_tet  # noqa: B018  (single color => must run single threaded: no two elements may race)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(1)
    except OSError:
        pass
    O.assemble_tet(N, lm.ien, lm.xg, off, ind, wg, dwg, F=F)
    O.assemble_tet(N, lm.ien, lm.xg, off, ind, wg, dwg, pattern=(rp, ci), blocks=blocks)
    O.assemble_face(lm.f2e, lm.forn, N, lm.ien, lm.xg, color, 1, wg, dwg, F=F)
    O.assemble_face(lm.f2e, lm.forn, N, lm.ien, lm.xg, color, 1, wg, dwg, pattern=(rp, ci), blocks=blocks)
    F[4 * N:] = 0
    for b, t in BCS.items():
        O.dirichlet_vec(lm.bound_nodes[b], np.array(t, np.int32), F)
        O.dirichlet_mat(lm.bound_nodes[b], np.array(t, np.int32), N, (rp, ci), blocks[0], blocks[1])
    for b_ in blocks:                                             # ghost rows: keep the preconditioner setup finite
        pass
    # make ghost-row diagonals harmless for the oracle's pc_setup
    for i in range(no, N):
        k = rp[i] + np.searchsorted(ci[rp[i]:rp[i + 1]], i)
        ln = rp[i + 1] - rp[i]
        for r in range(3):
            blocks[0][rp[i] * 9 + (k - rp[i]) * 3 + r * ln * 3 + r] = 1.0
        blocks[3][k] = 1.0
    d00, d11 = O.pc_setup((rp, ci), blocks)

    def compact(v):
        return np.concatenate([v[:3 * no], v[3 * N:3 * N + no]])

    def expand(c):
        v = np.zeros(6 * N)
        v[:3 * no] = c[:3 * no]
        v[3 * N:3 * N + no] = c[3 * no:]
        return v

    def matvec(v_local):
        _halo(lm, v_local)
        y = np.zeros(6 * N)
        O.fs_amvpby((rp, ci), blocks, 1.0, v_local, 0.0, y)
        return compact(y)

    def gdot(a, b_):
        t = torch.tensor([float(a @ b_)], dtype=torch.float64)
        dist.all_reduce(t)
        return t.item()

    maxit = 120
    r0 = compact(F) - matvec(np.zeros(6 * N))
    beta = np.zeros(maxit + 1)
    beta[0] = np.sqrt(gdot(r0, r0))
    Q = [r0 / beta[0]]
    H = np.zeros((maxit + 1, maxit))
    cs, sn, hist = [], [], [beta[0]]
    it = 0
    while it < maxit:
        z = O.pc_apply(d00, d11, expand(Q[it]))
        w = matvec(z)
        h = np.array([gdot(qj, w) for qj in Q])
        w = w - sum(hj * qj for hj, qj in zip(h, Q))
        nn = np.sqrt(gdot(w, w))
        Q.append(w / nn)
        col = np.concatenate([h, [nn]])
        for i in range(it):
            col[i], col[i + 1] = cs[i] * col[i] + sn[i] * col[i + 1], cs[i] * col[i + 1] - sn[i] * col[i]
        rr = np.hypot(col[it], col[it + 1])
        c, s = col[it] / rr, col[it + 1] / rr
        cs.append(c)
        sn.append(s)
        col[it], col[it + 1] = rr, 0.0
        H[:it + 2, it] = col
        beta[it + 1] = -s * beta[it]
        beta[it] *= c
        hist.append(abs(beta[it + 1]))
        it += 1
        if it % 20 == 0 and (hist[-1] < 1e-12 or hist[-1] < (hist[0] + 1e-16) * 1e-4):
            break
    y = np.linalg.solve(np.triu(H[:it, :it]), beta[:it])
    comb = sum(yj * qj for yj, qj in zip(y, Q[:it]))
    dxl = O.pc_apply(d00, d11, expand(comb))
    out = {"rank": rank, "it": it, "hist": np.array(hist), "nodes_g": lm.nodes_g[:no], "dx_u": dxl[:3 * no], "dx_p": dxl[3 * N:3 * N + no],
           "F_u": F[:3 * no], "F_p": F[3 * N:3 * N + no]}
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_coordinate_owner_partitions_any_mesh():
    """coordinate bisection: balanced, deterministic, usable on the unstructured mesh and on a RENUMBERED box (where the z-slab
    rule, which reads the node id, would scatter every rank over the whole domain); partition invariants hold"""
    mesh = boxmesh.delaunay_cube()
    for world in (2, 3, 8):
        npart = ddist.coordinate_owner(mesh, world)
        cnt = np.bincount(npart, minlength=world)
        assert cnt.max() - cnt.min() <= 1 and np.array_equal(npart, ddist.coordinate_owner(mesh, world))
        lms = [ddist.partition(mesh, npart, r, world) for r in range(world)]
        assert sum(lm.n_own for lm in lms) == mesh.num_node
        for lm in lms:
            assert np.array_equal(lm.nodes_g[lm.ien], mesh.ien[lm.elems_g])
            touched = np.zeros(lm.num_node, bool)
            touched[np.unique(lm.ien[(lm.ien >= lm.n_own).any(axis=1)])] = True
            assert not touched[:lm.n_interior].any() and touched[lm.n_interior:lm.n_own].all()
            for qi, q in enumerate(lm.neighbors):
                snd = lm.nodes_g[lm.send_nodes[lm.send_offset[qi]:lm.send_offset[qi + 1]]]
                other = lms[q]
                ri = list(other.neighbors).index(lm.rank)
                assert np.array_equal(snd, other.nodes_g[other.recv_nodes[other.recv_offset[ri]:other.recv_offset[ri + 1]]])
    box = boxmesh.make_box(6)
    perm = np.random.default_rng(1).permutation(box.num_node)          # renumber the nodes
    inv = np.argsort(perm)
    import copy
    ren = copy.copy(box)
    ren.xg = np.ascontiguousarray(box.xg[perm])
    ren.ien = np.ascontiguousarray(inv[box.ien].astype(np.int32))
    o_box, o_ren = ddist.coordinate_owner(box, 4), ddist.coordinate_owner(ren, 4)
    # same geometric parts whatever the numbering: halo sizes stay those of compact blocks
    halo = lambda mesh_, o: sum(ddist.partition(mesh_, o, r, 4, weak_group=99, bc_groups=()).recv_nodes.size for r in range(4))
    assert np.bincount(o_ren).tolist() == np.bincount(o_box).tolist()
    assert np.array_equal(o_ren, o_box[perm])                      # the same geometric parts
    assert halo(ren, o_ren) == halo(box, o_box)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("m", [6, 0])
def test_two_rank_gloo_solve_matches_single_domain_oracle(oracle, m):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-domain oracle
    from test_gpu_parity import oracle_system
    mesh, _ = _make_mesh(m)
    Ng = mesh.num_node
    wg, dwg = boxmesh.state_random(Ng)
    ref = oracle_system(oracle, mesh, wg, dwg)
    xo, ito, histo = oracle.gmres(ref["pattern"], ref["blocks"], ref["F"])
    dx = np.zeros(6 * Ng)
    Fg = np.zeros(6 * Ng)
    for o in outs:
        g = o["nodes_g"]
        dx[:3 * Ng].reshape(Ng, 3)[g] = o["dx_u"].reshape(-1, 3)
        dx[3 * Ng + g] = o["dx_p"]
        Fg[:3 * Ng].reshape(Ng, 3)[g] = o["F_u"].reshape(-1, 3)
        Fg[3 * Ng + g] = o["F_p"]
        assert o["it"] == ito
        assert np.abs(o["hist"] - histo).max() <= 1e-9 * histo[0]
    assert np.abs(Fg - ref["F"]).max() <= 1e-12 * np.abs(ref["F"]).max()
    assert np.abs(dx[:4 * Ng] - xo[:4 * Ng]).max() <= 1e-9 * np.abs(xo).max()
