"""One-GPU probe of the Jacobian assembly on the LOCAL meshes of a multi-GPU run: partitions the weak-scaling box of `world` ranks
with both ownership functions and times k_pairJ on the local mesh of a few ranks (plan statistics with DFB_VERBOSE=1).
usage: python scripts/j_partition_probe.py [world] [rank ...]"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from dedflow_b200 import boxmesh, dist as ddist, lib as dlib  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ranks = [int(a) for a in sys.argv[2:]] or [0, 3]
m = ddist.weak_scaling_m(55, world)
mesh = boxmesh.make_box(m)
L = dlib.load()
dlib.set_option("DFB_VERBOSE", 1)
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for name in ("slab", "rcb"):
    npart = ddist.OWNERS[name](mesh, world)
    for rank in ranks:
        lm = ddist.partition(mesh, npart, rank, world)
        N, E = lm.num_node, lm.num_tet
        ien = torch.from_numpy(lm.ien.reshape(-1)).cuda()
        xg = torch.from_numpy(lm.xg.reshape(-1)).cuda()
        row_ptr = torch.empty(N + 1, dtype=torch.int32, device="cuda")
        nnz = C.c_int(0)
        dlib.check(L.dfb_pattern_rows(N, E, P(ien), P(row_ptr), C.byref(nnz), st), "rows")
        col_ind = torch.empty(nnz.value, dtype=torch.int32, device="cuda")
        dlib.check(L.dfb_pattern_cols(N, E, P(ien), P(row_ptr), P(col_ind), st), "cols")
        A = [torch.zeros(k * nnz.value, dtype=torch.float64, device="cuda") for k in (9, 3, 3, 1)]
        plan = C.c_void_p()
        dlib.check(L.dfb_plan_create(C.byref(plan), N, E, P(ien), P(row_ptr), P(col_ind), 0, None, None, st), "plan")
        dlib.check(L.dfb_plan_set_rows(plan, lm.n_own), "set_rows")
        wg, dwg = (torch.from_numpy(a).cuda() for a in boxmesh.state_random(N))
        F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
        callJ = lambda: L.dfb_assemble_tet(plan, P(xg), P(wg), P(dwg), None, P(A[0]), P(A[1]), P(A[2]), P(A[3]), 1, 1, st)
        callF = lambda: L.dfb_assemble_tet(plan, P(xg), P(wg), P(dwg), P(F), None, None, None, None, 1, 1, st)
        out = {}
        for label, call in (("J", callJ), ("F", callF)):
            assert call() == 0, L.dfb_last_error()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); call(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            out[label] = float(np.median(ts)) * 1e3
        print(f"[probe] {name:4s} rank {rank}: {lm.n_own} owned / {N} local nodes, {E} local tets, {lm.neighbors.size} neighbours: "
              f"J {out['J']:.1f} us, F {out['F']:.1f} us", flush=True)
        L.dfb_plan_destroy(plan)
        del A, ien, xg, row_ptr, col_ind
