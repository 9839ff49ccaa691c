"""Block-Jacobi (the reference's preconditioner) against the opt-in two-level Schur-complement preconditioner on time-step
systems: GMRES iterations and solve time for the first Newton system of time step `nstep + 1`, and whole time steps.
usage: python scripts/pc_ab.py m [nsteps] [agg_cells,cheb_degree[,check_every] ...]   (check_every: DFB_GMRES_CHECK, default 20)"""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from dedflow_b200 import api, boxmesh, lib as dlib  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 55
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mesh = boxmesh.make_box(m)
N = mesh.num_node
print(f"m={m}: {mesh.num_tet} tets, {N} nodes", flush=True)
cfgs = [("jacobi", 4, 10, 20)] + [("schur2", int(c[0]), int(c[1]), int(c[2]) if len(c) > 2 else 20)
                                   for c in (a.split(",") for a in (sys.argv[3:] or ["4,10"]))]
for kind, agg_cells, degree, check in cfgs:
    dlib.set_option("DFB_GMRES_CHECK", check)
    fs = api.FlowSystem(mesh, with_colors=False)
    t0 = time.time()
    fs.set_preconditioner(kind, agg_cells=agg_cells, cheb_degree=degree)
    torch.cuda.synchronize()
    t_create = time.time() - t0
    state = [torch.from_numpy(a.copy()).cuda() for a in boxmesh.state_initial(mesh)]
    fs.time_step(*[s.clone() for s in state])          # warm-up (plans, graphs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    hists = [fs.time_step(*state) for _ in range(nsteps)]
    b.record()
    torch.cuda.synchronize()
    newton = sum(len(h) - 1 for h in hists)
    gm = [it for h in hists for _, it in h[1:]]
    last = hists[-1][-1][0]
    print(f"{kind:7s} agg {agg_cells} deg {degree} check {check}: {nsteps} time steps in {a.elapsed_time(b):9.2f} ms, {newton} Newton iterations, GMRES iterations per solve {gm}, "
          f"last Newton norms {np.array2string(last, precision=3)}, preconditioner create {t_create:.2f} s", flush=True)
    fs.close()
