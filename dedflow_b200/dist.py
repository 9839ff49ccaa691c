"""Data-parallel driver of the hot path: one process per GPU (torch.distributed for bootstrap only).

Decomposition (the ownership notion of reference src/partition.c:16-77, which is dead code there):
  * every node is owned by exactly one rank (`npart`; default: contiguous z-slabs of the structured box, which is what
    a nodal METIS split of a box approximates);
  * a rank keeps every element that touches an owned node ("ghost elements" are recomputed, so assembly needs no
    communication and stays deterministic) and assembles the rows of its owned nodes only;
  * local node numbering is [interior-owned | boundary-owned | ghost (grouped by owner)];
  * per mat-vec the (u,p) values of boundary-owned nodes go to the neighbours (NCCL send/recv on a side stream,
    overlapped with the interior rows); inner products are summed with ncclAllReduce.

Pure numpy partition logic lives at module level (tested on CPU with gloo, world_size 2); everything touching the GPU
is in DistFlowSystem.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time
from dataclasses import dataclass

import numpy as np


# ----------------------------------------------------------------------------------------------------------------
# partition (numpy, no GPU)
# ----------------------------------------------------------------------------------------------------------------
def slab_owner(mesh, world: int) -> np.ndarray:
    """Owner of every node for a z-slab split of the structured box: planes k are divided as evenly as possible."""
    n1 = mesh.m + 1
    plane_owner = np.zeros(n1, np.int32)
    for r, chunk in enumerate(np.array_split(np.arange(n1), world)):
        plane_owner[chunk] = r
    k = np.arange(mesh.num_node) // (n1 * n1)
    return plane_owner[k]


def coordinate_owner(mesh, world: int) -> np.ndarray:
    """Owner of every node by recursive coordinate bisection of the node coordinates: works for ANY mesh and node numbering
    (unstructured, renumbered), unlike slab_owner.  The node set is cut along its longest extent into two parts whose sizes
    are proportional to the ranks they receive (world need not be a power of two); ties are broken by the other two
    coordinates, then by node id: deterministic, and the same geometric parts whatever the numbering.  Compact parts with small surfaces are what the reference's METIS_PartMeshNodal call is after
    (src/partition.c:16-77; METIS itself is not available here)."""
    xg = np.asarray(mesh.xg, np.float64).reshape(-1, 3)
    owner = np.zeros(xg.shape[0], np.int32)
    stack = [(np.arange(xg.shape[0]), 0, world)]
    while stack:
        ids, r0, nr = stack.pop()
        if nr == 1 or ids.size == 0:
            owner[ids] = r0
            continue
        n1 = nr // 2
        ext = xg[ids].max(axis=0) - xg[ids].min(axis=0)
        axis = int(np.argmax(ext))
        b, c = (axis + 1) % 3, (axis + 2) % 3   # ties on the cut coordinate (grid planes) are broken geometrically, then by id
        order = ids[np.lexsort((ids, xg[ids, c], xg[ids, b], xg[ids, axis]))]
        cut = int(round(ids.size * n1 / nr))
        stack.append((order[:cut], r0, n1))
        stack.append((order[cut:], r0 + n1, nr - n1))
    return owner


OWNERS = {"slab": slab_owner, "rcb": coordinate_owner}


def pick_owner(name: str, mesh, world: int):
    """(name, function).  "auto": z-slabs when the box's planes divide evenly among the ranks (contiguous id ranges, two
    neighbours), coordinate bisection otherwise -- 111 planes over 8 ranks leave 13- and 14-plane slabs, a 7 % imbalance every
    collective then waits out (8 GPUs, 1M tets each: 5.96 ms per step by slabs, 5.86 ms by bisection)."""
    if name == "auto":
        name = "slab" if hasattr(mesh, "m") and (mesh.m + 1) % world == 0 else "rcb"
    return name, OWNERS[name]


@dataclass
class LocalMesh:
    rank: int
    world: int
    num_node_global: int
    num_tet_global: int
    nodes_g: np.ndarray        # [N_loc] global id of every local node, local order
    elems_g: np.ndarray        # [E_loc] global id of every local element (ascending)
    n_interior: int
    n_own: int
    ien: np.ndarray            # [E_loc,4] local node ids
    xg: np.ndarray             # [N_loc,3]
    neighbors: np.ndarray      # [nn]
    send_offset: np.ndarray    # [nn+1]
    send_nodes: np.ndarray     # local ids (owned), ascending global id per neighbour
    recv_offset: np.ndarray
    recv_nodes: np.ndarray     # local ids (ghost), ascending global id per neighbour
    bound_nodes: dict          # boundary group -> local ids of OWNED nodes in the group
    f2e: np.ndarray            # weak-BC faces (group 4): local element ids
    forn: np.ndarray

    @property
    def num_node(self):
        return self.nodes_g.size

    @property
    def num_tet(self):
        return self.elems_g.size

    def localize(self, v_global: np.ndarray) -> np.ndarray:
        """6N-layout global vector -> 6N_loc-layout local vector (owned + ghost entries)."""
        Ng, g = self.num_node_global, self.nodes_g
        out = np.empty(6 * g.size)
        out[:3 * g.size] = v_global[:3 * Ng].reshape(Ng, 3)[g].ravel()
        for s in (3, 4, 5):
            out[s * g.size:(s + 1) * g.size] = v_global[s * Ng + g]
        return out

    def scatter_owned(self, v_local: np.ndarray, v_global: np.ndarray):
        """write the owned entries of a local 6N_loc vector into a global 6N vector"""
        Ng, g, n = self.num_node_global, self.nodes_g[:self.n_own], self.nodes_g.size
        v_global[:3 * Ng].reshape(Ng, 3)[g] = v_local[:3 * n].reshape(n, 3)[:self.n_own]
        for s in (3, 4, 5):
            v_global[s * Ng + g] = v_local[s * n:s * n + self.n_own]


def partition(mesh, npart: np.ndarray, rank: int, world: int, weak_group: int = 4, bc_groups=(0, 2, 3, 4)) -> LocalMesh:
    owned_mask = npart == rank
    owner_e = npart[mesh.ien]                                  # [E,4]
    el_mask = (owner_e == rank).any(axis=1)
    elems_g = np.nonzero(el_mask)[0]
    ien_g = mesh.ien[elems_g]
    own_e = owner_e[elems_g]
    nodes_all = np.unique(ien_g)
    owned_g = nodes_all[owned_mask[nodes_all]]
    ghost_g = nodes_all[~owned_mask[nodes_all]]
    # (my node a, foreign owner q) pairs: a is needed by q  <=>  a shares an element with a q-owned node
    mine = own_e == rank
    pairs = []
    for a in range(4):
        for b in range(4):
            sel = mine[:, a] & ~mine[:, b]
            if sel.any():
                pairs.append(np.stack([own_e[sel, b].astype(np.int64), ien_g[sel, a].astype(np.int64)], axis=1))
    if pairs:
        pr = np.unique(np.concatenate(pairs), axis=0)          # sorted by (q, global id)
    else:
        pr = np.zeros((0, 2), np.int64)
    bnd_g = np.unique(pr[:, 1]).astype(ien_g.dtype)
    interior_g = np.setdiff1d(owned_g, bnd_g, assume_unique=True)
    ghost_owner = npart[ghost_g]
    gorder = np.lexsort((ghost_g, ghost_owner))
    ghost_sorted = ghost_g[gorder]
    nodes_g = np.concatenate([interior_g, bnd_g, ghost_sorted]).astype(np.int64)
    g2l = np.full(mesh.num_node, -1, np.int64)
    g2l[nodes_g] = np.arange(nodes_g.size)
    neighbors = np.unique(np.concatenate([pr[:, 0], ghost_owner[gorder].astype(np.int64)])).astype(np.int32)
    send_off, recv_off, send_nodes, recv_nodes = [0], [0], [], []
    gown_sorted = ghost_owner[gorder]
    for q in neighbors:
        s = pr[pr[:, 0] == q, 1]
        r = ghost_sorted[gown_sorted == q]
        send_nodes.append(g2l[s])
        recv_nodes.append(g2l[r])
        send_off.append(send_off[-1] + s.size)
        recv_off.append(recv_off[-1] + r.size)
    cat = lambda l: (np.concatenate(l) if l else np.zeros(0, np.int64)).astype(np.int32)
    bound_nodes = {}
    for b in bc_groups:
        if b < mesh.num_bound:
            gn = mesh.bound_nodes(b)
            gn = gn[owned_mask[gn]]
            bound_nodes[b] = g2l[gn].astype(np.int32)
    if weak_group < mesh.num_bound:
        f2e_g, forn_g = mesh.bound_faces(weak_group)
        keep = el_mask[f2e_g]
        f2e = np.searchsorted(elems_g, f2e_g[keep]).astype(np.int32)
        forn = forn_g[keep].astype(np.int32)
    else:
        f2e, forn = np.zeros(0, np.int32), np.zeros(0, np.int32)
    return LocalMesh(rank=rank, world=world, num_node_global=mesh.num_node, num_tet_global=mesh.num_tet, nodes_g=nodes_g,
                     elems_g=elems_g, n_interior=interior_g.size, n_own=owned_g.size,
                     ien=np.ascontiguousarray(g2l[ien_g].astype(np.int32)), xg=np.ascontiguousarray(mesh.xg[nodes_g]),
                     neighbors=neighbors, send_offset=np.array(send_off, np.int32), send_nodes=cat(send_nodes),
                     recv_offset=np.array(recv_off, np.int32), recv_nodes=cat(recv_nodes), bound_nodes=bound_nodes,
                     f2e=f2e, forn=forn)


# ----------------------------------------------------------------------------------------------------------------
# GPU side
# ----------------------------------------------------------------------------------------------------------------
class ParallelOps(C.Structure):
    _fields_ = [("n_own", C.c_int), ("n_interior", C.c_int), ("allreduce", C.c_void_p), ("halo_begin", C.c_void_p),
                ("halo_end", C.c_void_p), ("user", C.c_void_p), ("p2p", C.c_void_p), ("halo_begin_aos", C.c_void_p)]


class DistFlowSystem:
    """The hot path of one rank.  Needs torch.distributed initialised (any backend) for the NCCL id broadcast."""

    def __init__(self, lm: LocalMesh, device, bcs=None, max_iter=120, atol=1e-12, rtol=1e-4, peer_memory=True):
        import torch
        import torch.distributed as dist
        from . import api, lib as _lib
        self.torch, self._lib = torch, _lib
        self.L = _lib.load()
        self.lm = lm
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        L = self.L
        self.N, self.E = lm.num_node, lm.num_tet
        self.n_own, self.n_int = lm.n_own, lm.n_interior
        self.ien = torch.from_numpy(lm.ien.reshape(-1)).to(self.dev)
        self.xg = torch.from_numpy(lm.xg.reshape(-1)).to(self.dev)
        self.bcs = dict(api.DEFAULT_BCS if bcs is None else bcs)
        self.bnode = {b: torch.from_numpy(lm.bound_nodes[b]).to(self.dev) for b in self.bcs if b in lm.bound_nodes}
        self.f2e = torch.from_numpy(lm.f2e).to(self.dev)
        self.forn = torch.from_numpy(lm.forn).to(self.dev)
        self.max_iter, self.atol, self.rtol = max_iter, atol, rtol
        p = lambda t: C.c_void_p(t.data_ptr())
        st = self._stream()
        # local pattern (rows of ghost nodes are incomplete and never used)
        self.row_ptr = torch.empty(self.N + 1, dtype=torch.int32, device=self.dev)
        nnz = C.c_int(0)
        _lib.check(L.dfb_pattern_rows(self.N, self.E, p(self.ien), p(self.row_ptr), C.byref(nnz), st), "dfb_pattern_rows")
        self.nnz = nnz.value
        self.col_ind = torch.empty(self.nnz, dtype=torch.int32, device=self.dev)
        _lib.check(L.dfb_pattern_cols(self.N, self.E, p(self.ien), p(self.row_ptr), p(self.col_ind), st), "dfb_pattern_cols")
        Z = self.nnz
        self.A = [torch.zeros(k * Z, dtype=torch.float64, device=self.dev) for k in (9, 3, 3, 1)]
        plan = C.c_void_p()
        _lib.check(L.dfb_plan_create(C.byref(plan), self.N, self.E, p(self.ien), p(self.row_ptr), p(self.col_ind), 0, None,
                                     None, st), "dfb_plan_create")
        self.plan = plan
        _lib.check(L.dfb_plan_set_rows(plan, self.n_own), "dfb_plan_set_rows")
        # communicator
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if lm.rank == 0:
            raw = (C.c_ubyte * 128)()
            _lib.check(L.dfb_comm_unique_id(raw), "dfb_comm_unique_id")
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            idbuf = idbuf.to(self.dev)
        dist.broadcast(idbuf, src=0)
        raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
        comm = C.c_void_p()
        _lib.check(L.dfb_comm_create(C.byref(comm), lm.rank, lm.world, raw), "dfb_comm_create")
        self.comm = comm
        ip = lambda a: np.ascontiguousarray(a, np.int32).ctypes.data_as(C.c_void_p)
        self._halo_keep = [np.ascontiguousarray(a, np.int32) for a in (lm.neighbors, lm.send_offset, lm.send_nodes,
                                                                     lm.recv_offset, lm.recv_nodes)]
        _lib.check(L.dfb_comm_set_halo(comm, self.N, lm.neighbors.size, *[a.ctypes.data_as(C.c_void_p) for a in self._halo_keep]),
                   "dfb_comm_set_halo")
        self.p2p = self._connect_peer_memory(dist) if (peer_memory and lm.world > 1) else None
        ws = C.c_void_p()
        _lib.check(L.dfb_gmres_create(C.byref(ws), self.N, max_iter), "dfb_gmres_create")
        self.gmres = ws
        fn = lambda name: C.cast(getattr(L, name), C.c_void_p).value
        self.ops = ParallelOps(self.n_own, self.n_int, fn("dfb_comm_allreduce"), fn("dfb_comm_halo_begin"),
                               fn("dfb_comm_halo_end"), comm.value, self.p2p, fn("dfb_comm_halo_begin_aos"))
        _lib.check(L.dfb_gmres_set_parallel(ws, C.byref(self.ops)), "dfb_gmres_set_parallel")

    def _connect_peer_memory(self, dist):
        """CUDA IPC over NVLink: every rank maps every other rank's mailbox + z vector; the collectives of a GMRES iteration
        are then fused into the compute kernels (include/dedflow_b200.h, dfb_comm_p2p_*).  Returns the view pointer, or
        None (NCCL path) when the mapping is not possible on this box -- decided collectively so that all ranks agree."""
        torch, L, lm = self.torch, self.L, self.lm
        if lm.world > 8 or os.environ.get("DFB_NO_PEER_MEMORY"):
            return None
        handle = (C.c_ubyte * 64)()
        ok = L.dfb_comm_p2p_alloc(self.comm, handle) == 0
        mine = {"rank": lm.rank, "handle": bytes(handle), "ok": ok, "n_local": lm.num_node, "neighbors": lm.neighbors.tolist(),
                "recv_offset": lm.recv_offset.tolist(), "recv_nodes": lm.recv_nodes.astype(np.int32)}
        everyone = [None] * lm.world
        dist.all_gather_object(everyone, mine)
        everyone.sort(key=lambda d: d["rank"])
        view = None
        if all(d["ok"] for d in everyone):
            handles = b"".join(d["handle"] for d in everyone)
            remote = np.zeros(max(1, lm.send_nodes.size), np.int32)
            nloc = np.zeros(max(1, lm.neighbors.size), np.int32)
            for q, nb in enumerate(lm.neighbors.tolist()):
                d = everyone[nb]
                qq = d["neighbors"].index(lm.rank)              # my position in the neighbour's own neighbour list
                ghosts = d["recv_nodes"][d["recv_offset"][qq]:d["recv_offset"][qq + 1]]
                s0, s1 = int(lm.send_offset[q]), int(lm.send_offset[q + 1])
                assert ghosts.size == s1 - s0, "halo lists of neighbouring ranks disagree"
                remote[s0:s1] = ghosts                          # both sides list the nodes in ascending global id
                nloc[q] = d["n_local"]
            hb = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
            ok = L.dfb_comm_p2p_connect(self.comm, hb, remote.ctypes.data_as(C.c_void_p), nloc.ctypes.data_as(C.c_void_p)) == 0
            if ok:
                view = L.dfb_comm_p2p_view(self.comm)
        flag = torch.tensor([1 if view else 0], device=self.dev if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if lm.rank == 0:
                print("dedflow_b200: peer-memory mode unavailable (%s); using the NCCL path" %
                      L.dfb_last_error().decode(errors="replace"), file=sys.stderr)
            return None
        return view

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def assemble_system(self, wg, dwg, F=None, J=False):
        L, N, st, _lib = self.L, self.N, self._stream(), self._lib
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        A = self.A if J else [None] * 4
        _lib.check(L.dfb_assemble_tet(self.plan, p(self.xg), p(wg), p(dwg), p(F), *[p(a) for a in A], 1, 1, st), "dfb_assemble_tet")
        if self.f2e.numel():
            _lib.check(L.dfb_assemble_face(self.plan, self.f2e.numel(), p(self.f2e), p(self.forn), p(self.xg), p(wg), p(dwg), p(F),
                                           *[p(a) for a in A], st), "dfb_assemble_face")
        if F is not None:
            F[4 * N:].zero_()
        for b, types in self.bcs.items():
            if b not in self.bnode or self.bnode[b].numel() == 0:
                continue
            t = (C.c_int * 3)(*types)
            if F is not None:
                _lib.check(L.dfb_dirichlet_vec(self.bnode[b].numel(), p(self.bnode[b]), 3, t, p(F), st), "dfb_dirichlet_vec")
            if J:
                _lib.check(L.dfb_dirichlet_mat(self.bnode[b].numel(), p(self.bnode[b]), 3, t, N, p(self.row_ptr), p(self.col_ind),
                                               p(self.A[0]), p(self.A[1]), st), "dfb_dirichlet_mat")

    def krylov_solve(self, dx, F):
        iters = C.c_int(0)
        hist = np.zeros(self.max_iter + 1)
        p = lambda t: C.c_void_p(t.data_ptr())
        self._lib.check(self.L.dfb_gmres_solve(self.gmres, self.N, p(self.row_ptr), p(self.col_ind), *[p(a) for a in self.A], p(dx),
                                               p(F), self.atol, self.rtol, C.byref(iters), hist.ctypes.data_as(C.c_void_p),
                                               self._stream()), "dfb_gmres_solve")
        return iters.value, hist[:iters.value + 1]

    # ------------------------------------------------------------------ the driver around the path
    def solve_flow_system(self, wgold, dwgold, dwg, maxit=4, tol=0.5e-3):
        """SolveFlowSystem (main.c:77-283) on the local (owned + ghost) entries.  The pointwise updates keep the ghost entries
        consistent because dx's ghosts are refreshed at the end of every solve; block norms run over the owned nodes and are
        summed over ranks.  Returns [(rnorm[4], gmres iterations)] (identical on all ranks)."""
        import torch.distributed as dist
        torch, L, N, st, _lib = self.torch, self.L, self.N, self._stream(), self._lib
        p = lambda t: C.c_void_p(t.data_ptr())
        if getattr(self, "_newton_ws", None) is None:
            mk = lambda n=6 * N: torch.zeros(n, dtype=torch.float64, device=self.dev)
            self._newton_ws = (mk(), mk(), mk(), mk(), mk(4))
        wgalpha, dwgalpha, F, dx, ss = self._newton_ws

        def stage():
            _lib.check(L.dfb_genalpha_stage(N, p(wgold), p(dwgold), p(dwg), p(wgalpha), p(dwgalpha), st), "dfb_genalpha_stage")

        def residual():
            self.assemble_system(wgalpha, dwgalpha, F=F)
            _lib.check(L.dfb_block_sumsq(N, self.n_own, p(F), p(ss), st), "dfb_block_sumsq")
            dist.all_reduce(ss)
            return np.sqrt(ss.cpu().numpy())

        stage()
        r0 = residual()
        hist = [(r0.copy(), 0)]
        r0 = r0 + 1e-16
        it, converged = 0, False
        while not converged and it < maxit:
            self.assemble_system(wgalpha, dwgalpha, J=True)
            dx.zero_()
            its, _ = self.krylov_solve(dx, F)
            _lib.check(L.dfb_newton_update(N, p(dx), p(dwg), st), "dfb_newton_update")
            stage()
            r = residual()
            hist.append((r.copy(), its))
            converged = bool(np.all(r < tol * r0))
            it += 1
        return hist

    def time_step(self, wgold, dwgold, dwg, **kw):
        """one pass of the time loop of main.c:537-565"""
        L, N, st, _lib = self.L, self.N, self._stream(), self._lib
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(L.dfb_genalpha_predict(N, p(dwg), st), "dfb_genalpha_predict")
        hist = self.solve_flow_system(wgold, dwgold, dwg, **kw)
        _lib.check(L.dfb_genalpha_correct(N, p(wgold), p(dwgold), p(dwg), st), "dfb_genalpha_correct")
        return hist

    def matvec_owned(self, x, y):
        """y[owned rows, compact 4*n_own] = A x (ghosts of x refreshed first) -- used by the parity script."""
        p = lambda t: C.c_void_p(t.data_ptr())
        self._lib.check(self.L.dfb_comm_halo(self.comm, p(x), self._stream()), "dfb_comm_halo")
        self._lib.check(self.L.dfb_spmv_fs(self.N, p(self.row_ptr), p(self.col_ind), *[p(a) for a in self.A], 1.0, p(x), 0.0, p(y),
                                           self._stream()), "dfb_spmv_fs")

    def close(self):
        if getattr(self, "gmres", None):
            self.L.dfb_gmres_destroy(self.gmres)
            self.gmres = None
        if getattr(self, "plan", None):
            self.L.dfb_plan_destroy(self.plan)
            self.plan = None
        if getattr(self, "comm", None):
            self.L.dfb_comm_destroy(self.comm)
            self.comm = None


def weak_scaling_m(m1: int, world: int) -> int:
    """cells per direction so that the element count per GPU stays that of an m1^3 box"""
    return int(round(m1 * world ** (1.0 / 3.0)))


def gather_owned(lm, v_local, dist, torch):
    """global 6N vector (on every rank) from the owned entries of the ranks' local vectors"""
    g = np.zeros(6 * lm.num_node_global)
    lm.scatter_owned(v_local, g)
    t = torch.from_numpy(g).cuda()
    dist.all_reduce(t)
    return t.cpu().numpy()


def bench_parity(dist, torch, rank, world, local_rank, owner_fn=None, m=20, mesh=None, label=None):
    """The embedded parity check of `bench.py --gpus N`: BASELINE configs[0] (m=20, 48,000 tets) split over ALL N ranks, the
    same collectives mode as the timed run (peer memory when available), one assembly + solve, gathered and compared on rank 0
    with the single-domain CPU oracle.  Bars: F <= 1e-12, dx and residual history <= 1e-10, identical iteration count."""
    from . import boxmesh
    mesh = mesh if mesh is not None else boxmesh.make_box(m)
    Ng = mesh.num_node
    npart = (owner_fn or slab_owner)(mesh, world)
    lm = partition(mesh, npart, rank, world)
    wg_g, dwg_g = boxmesh.state_random(Ng)
    fs = DistFlowSystem(lm, f"cuda:{local_rank}")
    N = fs.N
    d_wg, d_dwg = torch.from_numpy(lm.localize(wg_g)).cuda(), torch.from_numpy(lm.localize(dwg_g)).cuda()
    F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    dx = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    fs.assemble_system(d_wg, d_dwg, F=F)
    fs.assemble_system(d_wg, d_dwg, J=True)
    it, hist = fs.krylov_solve(dx, F)
    torch.cuda.synchronize()
    Fg = gather_owned(lm, F.cpu().numpy(), dist, torch)
    xg = gather_owned(lm, dx.cpu().numpy(), dist, torch)
    ghost = float(np.abs(lm.localize(xg)[:4 * N] - dx.cpu().numpy()[:4 * N]).max())   # ghosts of the solution after the final halo
    out = {"mesh": f"{label or f'Kuhn box m={m}'} ({mesh.num_tet} tets) over {world} ranks, "
                   f"{(owner_fn or slab_owner).__name__}", "collectives": "peer-memory" if fs.p2p else "nccl",
           "checker": "oracle/oracle.c (CPU, single domain)"}
    ok = True
    if rank == 0:
        from oracle import pyoracle
        ref = pyoracle.get().reference_step(mesh, wg_g, dwg_g)
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        out["F_rel"] = rel(Fg, ref["F"])
        out["dx_rel"] = rel(xg[:4 * Ng], ref["dx"][:4 * Ng])
        out["iters"] = int(it)
        out["iters_equal"] = bool(it == ref["iters"])
        out["hist_rel"] = float(np.abs(hist - ref["hist"]).max() / ref["hist"][0]) if out["iters_equal"] else float("inf")
        out["ghost_abs"] = ghost
        ok = (out["F_rel"] <= 1e-12 and out["dx_rel"] <= 1e-10 and out["hist_rel"] <= 1e-10 and out["iters_equal"] and
              ghost <= 1e-12 * float(np.abs(ref["dx"]).max()))
        out["ok"] = ok
    fs.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    out["ok"] = bool(flag.item())
    return out


def _bench_case(B, args, dist, torch, dlib, mesh, rank, world, local_rank, fixed_its, steps, warmup, full, owner_fn=None):
    """one mesh over all ranks: setup, warm-up, timed steps (barrier + synchronize on both sides, max over ranks)"""
    from . import boxmesh
    t0 = time.time()
    npart = (owner_fn or slab_owner)(mesh, world)
    lm = partition(mesh, npart, rank, world)
    Ng, Eg = mesh.num_node, mesh.num_tet
    wg_g, dwg_g = boxmesh.state_random(Ng)
    wg, dwg = lm.localize(wg_g), lm.localize(dwg_g)
    del wg_g, dwg_g
    fs = (DistFlowSystem(lm, f"cuda:{local_rank}") if fixed_its is None else
          DistFlowSystem(lm, f"cuda:{local_rank}", max_iter=fixed_its, atol=0.0, rtol=0.0))
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    N = fs.N
    h_wg, h_dwg = torch.from_numpy(wg).pin_memory(), torch.from_numpy(dwg).pin_memory()
    h_dx = torch.zeros(6 * N, dtype=torch.float64).pin_memory()
    d_wg, d_dwg = h_wg.cuda(), h_dwg.cuda()
    F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    dx = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    state = {}

    def step():
        fs.assemble_system(d_wg, d_dwg, F=F)
        fs.assemble_system(d_wg, d_dwg, J=True)
        dx.zero_()
        state["iters"], state["hist"] = fs.krylov_solve(dx, F)

    copy_stream = torch.cuda.Stream()
    ev_wg, ev_dwg = torch.cuda.Event(), torch.cuda.Event()

    def step_e2e():
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            d_wg[:3 * N].copy_(h_wg[:3 * N], non_blocking=True)      # the velocities: all the Jacobian reads
            ev_wg.record(copy_stream)
            d_wg[3 * N:].copy_(h_wg[3 * N:], non_blocking=True)
            d_dwg.copy_(h_dwg, non_blocking=True)
            ev_dwg.record(copy_stream)
        cur.wait_event(ev_wg)
        fs.assemble_system(d_wg, d_dwg, J=True)
        cur.wait_event(ev_dwg)
        fs.assemble_system(d_wg, d_dwg, F=F)
        dx.zero_()
        state["iters"], state["hist"] = fs.krylov_solve(dx, F)
        h_dx.copy_(dx, non_blocking=True)
        torch.cuda.synchronize()

    def timed(fn, k):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)          # max over ranks
        return t.item() / k

    sampler = B.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # priming (first-use work + >= 1 s of load; a fixed count so that all ranks agree), then exactly `warmup` untimed steps
    prime = 2 if not full else max(2, int(np.ceil(1.0 / 0.006 / 8)))
    for _ in range(prime + warmup):
        step()
    l0 = dlib.launch_count()
    ms = timed(step, steps)
    launches = dlib.launch_count() - l0
    res = {"ms": ms, "launches": int(launches), "setup_s": setup_s, "N_local": N, "Ng": Ng, "Eg": Eg, "prime": prime,
           "p2p": bool(fs.p2p)}
    if full:
        res["ms_e2e"] = timed(step_e2e, steps)
    res["tF"] = timed(lambda: fs.assemble_system(d_wg, d_dwg, F=F), 3)
    res["tJ"] = timed(lambda: fs.assemble_system(d_wg, d_dwg, J=True), 3)

    def solve():
        dx.zero_()
        state["iters"], state["hist"] = fs.krylov_solve(dx, F)
    res["t_solve"] = timed(solve, 3)
    res["iters"], res["hist"] = int(state["iters"]), state["hist"]
    # per-kernel CUDA-event times of one more solve on every rank (diagnostic): the in-solve mat-vec (halo wait included) is the
    # roofline kernel, the waits of the fused collectives show up in the kernels that poll for them
    dlib.set_option("DFB_PROFILE", -1)
    solve()
    prof = dlib.solve_profile(fs.gmres)
    dlib.set_option("DFB_PROFILE", 0)
    keys = sorted(prof)
    t = torch.tensor([prof[k]["avg_us"] for k in keys], device="cuda", dtype=torch.float64)
    tmax, tmin = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    res["solve_kernels"] = {k: {"launches": prof[k]["launches"], "avg_us_max_rank": float(a), "avg_us_min_rank": float(b)}
                            for k, a, b in zip(keys, tmax.tolist(), tmin.tolist())}
    res["clocks"] = sampler.stop() if rank == 0 else None
    Zt = torch.tensor([float(fs.row_ptr[fs.n_own].item())], device="cuda", dtype=torch.float64)   # owned nonzeros
    dist.all_reduce(Zt)
    res["Zg"] = int(Zt.item())
    fs.close()
    del fs, d_wg, d_dwg, F, dx
    torch.cuda.empty_cache()
    return res


def _strong_block(B, args, dist, torch, dlib, m, rank, world, local_rank, steps):
    """BASELINE configs[2] / [3]: the m^3 box split over all ranks with the reference's stopping rule, followed by the SAME mesh
    on rank 0's GPU alone (the other ranks wait) -- the 1-GPU denominator, measured in the same run at the same clocks."""
    from . import boxmesh
    mesh = boxmesh.make_box(m)
    r = _bench_case(B, args, dist, torch, dlib, mesh, rank, world, local_rank, None, steps, 1, False)
    one = None
    if rank == 0:
        try:
            one = B.strong_one_gpu(m, local_rank, steps=steps, mesh=mesh)
        except Exception as e:
            one = {"error": repr(e)[:200]}
    del mesh
    dist.barrier()
    if rank != 0:
        return None
    blk = {"m": m, "elems": r["Eg"], "nodes": r["Ng"], "n_gpus": world, "steps": steps, "ms_per_step": r["ms"],
           "elems_per_s": r["Eg"] / (r["ms"] * 1e-3), "gmres_iters": r["iters"], "assemble_F_ms": r["tF"], "assemble_J_ms": r["tJ"],
           "solve_ms": r["t_solve"], "setup_s": r["setup_s"], "clocks": r["clocks"], "solve_kernels": r["solve_kernels"],
           "one_gpu": one}
    if one and "ms_per_step" in one:
        blk["efficiency_vs_one_gpu_same_run"] = one["ms_per_step"] / (world * r["ms"])
    return blk


def bench_main(args, rank, world, local_rank, B=None):
    """bench.py --gpus N (N > 1).  In this order: (1) the embedded parity check (m=20 over all N ranks against the CPU oracle; a
    missed bar fails the run), (2) the timed weak-scaling case (~the configs[1] element count per GPU), (3) the strong-scaling
    blocks on the north star's meshes (16M tets; 64M at N = 8) with their 1-GPU denominators.  B = the bench module (it owns
    the protected stdout the JSON line goes to)."""
    import torch
    import torch.distributed as dist
    from . import boxmesh, lib as dlib
    if B is None:
        import bench as B
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if getattr(args, "timesteps", 0) > 0:
        m = args.m if args.fixed_m else weak_scaling_m(args.m, world)
        t0 = time.time()
        mesh = boxmesh.make_box(m)
        lm = partition(mesh, slab_owner(mesh, world), rank, world)
        Ng, Eg = mesh.num_node, mesh.num_tet
        del mesh
        fs = DistFlowSystem(lm, f"cuda:{local_rank}")
        torch.cuda.synchronize()
        ret = B.run_timesteps(args, fs, lm, lm.localize, world, time.time() - t0, dist=dist, rank=rank, Eg=Eg, Ng=Ng)
        dist.destroy_process_group()
        return ret
    parity = None
    if not getattr(args, "no_parity", False):
        parity = bench_parity(dist, torch, rank, world, local_rank)
        # the same on an UNSTRUCTURED mesh (ragged rows, arbitrary numbering) split by coordinate bisection
        try:
            dmesh = boxmesh.delaunay_cube(2400, 120)
        except ImportError:
            dmesh = None
        if dmesh is not None:
            p2 = bench_parity(dist, torch, rank, world, local_rank, owner_fn=coordinate_owner, mesh=dmesh, label="Delaunay cube")
            parity = {"box": parity, "delaunay": p2, "ok": bool(parity["ok"] and p2["ok"])}
        if not parity["ok"]:
            if rank == 0:
                print("bench.py: embedded parity check FAILED: " + json.dumps(parity), file=sys.stderr)
            dist.destroy_process_group()
            raise SystemExit(1)
    m = args.m if args.fixed_m else weak_scaling_m(args.m, world)
    # Weak scaling keeps the work per GPU fixed: the mesh grows with N, and so would the iteration count the reference's
    # stopping rule needs (40 at 1M tets, 60 at 4-8M).  The solve is therefore pinned to the 40 iterations configs[1] needs
    # on one GPU; strong scaling (--fixed-m and the strong blocks) runs the reference's stopping rule unchanged.
    fixed_its = None if args.fixed_m else 40
    mesh = boxmesh.make_box(m)
    owner_name, owner_fn = pick_owner(getattr(args, "owner", "auto"), mesh, world)
    r = _bench_case(B, args, dist, torch, dlib, mesh, rank, world, local_rank, fixed_its, args.steps, args.warmup, True,
                    owner_fn=owner_fn)
    del mesh
    strong = {}
    if not args.fixed_m and not getattr(args, "no_strong", False) and args.m == 55:
        strong["strong_16M"] = _strong_block(B, args, dist, torch, dlib, 139, rank, world, local_rank, 3)
        if world >= 8:
            strong["strong_64M"] = _strong_block(B, args, dist, torch, dlib, 220, rank, world, local_rank, 2)
    if rank == 0:
        hbm, hbm_src = B.measured_hbm_peak()
        Ng, Eg, Zg, N = r["Ng"], r["Eg"], r["Zg"], r["N_local"]
        ab = B.algorithmic_bytes(Ng, Eg, Zg)
        its = r["iters"]
        # roofline kernel: the mat-vec AS IT RUNS INSIDE THE SOLVE (peer-memory mode: one launch whose boundary-row blocks wait
        # for the neighbours' halo stores), slowest rank, all ranks' bytes
        t_spmv = r["solve_kernels"].get("spmv", {}).get("avg_us_max_rank", float("nan")) * 1e-3
        spmv_gbs = ab["spmv"] / (t_spmv * 1e-3) / 1e9
        line = {
            "metric": B.METRIC, "value": Eg / (r["ms"] * 1e-3), "unit": B.UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "strong" if args.fixed_m else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Kuhn box m={m}: {Eg} tets, {Ng} nodes over {world} GPUs ({owner_name} node ownership, ghost elements "
                                   f"recomputed; halo + all-reduces {'fused into the Krylov kernels over NVLink peer memory' if r['p2p'] else 'by NCCL'}); "
                                   f"step = AssembleSystem(F)+AssembleSystem(J)+KrylovSolve ({its} GMRES iterations"
                                   f"{', pinned: weak scaling keeps per-GPU work fixed' if fixed_its else ', reference stopping rule'}), state B",
                       "elements_per_gpu": Eg / world, "collectives": "peer-memory" if r["p2p"] else "nccl", "prime_steps": r["prime"],
                       "l2": "per-GPU working set exceeds the 126 MB L2; no explicit flush"},
            "clocks": r["clocks"],
            "parity": parity,
            "e2e": {"value": Eg / (r["ms_e2e"] * 1e-3), "unit": B.UNIT, "ms_per_step": r["ms_e2e"], "h2d_bytes_per_step": 2 * 6 * N * 8 * world,
                    "d2h_bytes_per_step": 6 * N * 8 * world},
            "gpu_launches": r["launches"],
            "roofline": {"kernel": "k_spmv_fs<8,peer> as launched inside the solve (halo wait included), slowest rank", "bound": "hbm",
                         "achieved": spmv_gbs, "peak": hbm * world, "unit": "GB/s", "frac": spmv_gbs / (hbm * world),
                         "traffic": None, "traffic_note": "ncu is a one-GPU tool here (never run on a multi-rank command); the per-launch "
                                                           "DRAM bytes of the same kernel at N=1 are in the N=1 line",
                         "peak_source": hbm_src + f" x {world} GPUs", "bytes_per_launch": ab["spmv"], "ms_per_launch": t_spmv},
            "breakdown": {"assemble_F_ms": r["tF"], "assemble_J_ms": r["tJ"], "assemble_elems_per_s": Eg / ((r["tF"] + r["tJ"]) * 1e-3),
                          "spmv_ms": t_spmv, "spmv_gbs": spmv_gbs, "spmv_pct_hbm": 100 * spmv_gbs / (hbm * world),
                          "solve_s_per_step": r["t_solve"] * 1e-3, "gmres_iters": its, "setup_s": r["setup_s"],
                          "final_residual": float(r["hist"][-1]), "initial_residual": float(r["hist"][0]),
                          "solve_kernels": r["solve_kernels"]},
        }
        line.update({k: v for k, v in strong.items() if v is not None})
        B.emit(line)
    dist.destroy_process_group()
