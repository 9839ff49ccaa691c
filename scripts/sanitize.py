"""Small end-to-end pass of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python scripts/sanitize.py
Unstructured mesh (ragged rows) + box mesh, all assembly variants, SpMV, GMRES, Newton kernels, drop-in entry points."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import delaunay_mesh  # noqa: E402
from dedflow_b200 import api, boxmesh, lib as dlib  # noqa: E402

for mesh in (delaunay_mesh(120, 12), boxmesh.make_box(5)):
    N = mesh.num_node
    fs = api.FlowSystem(mesh)
    wg, dwg = (torch.from_numpy(a).cuda() for a in boxmesh.state_random(N))
    F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    for mode in ("gather", "atomic", "colored"):
        fs.assemble_system(wg, dwg, F=F, mode=mode)
        fs.assemble_system(wg, dwg, J=True, mode=mode)
    import os
    for variant in ("pull", "fused", "pairs"):       # the three atomic-free Jacobian kernels (read per call)
        dlib.set_option("DFB_J_VARIANT", variant)
        fs.assemble_system(wg, dwg, J=True, mode="gather")
    dlib.set_option("DFB_J_VARIANT", "pairs")
    for variant in ("scratch", "pipe", "patch"):     # the residual kernels
        dlib.set_option("DFB_F_VARIANT", variant)
        fs.assemble_system(wg, dwg, F=F, mode="gather")
    x = torch.randn(6 * N, dtype=torch.float64, device="cuda")
    y = torch.zeros_like(x)
    fs.matrix_matvec(x, y)
    dx = torch.zeros_like(x)
    it, hist = fs.krylov_solve(dx, F)
    for key, val in (("DFB_GMRES_CHECK", 5), ("DFB_GIVENS_DEFER", 0), ("DFB_GRAPH", 0)):   # the other solver paths
        dlib.set_option(key, val)
        fs.krylov_solve(dx, F)                        # (x0 != 0 from here on: r0 runs its mat-vec)
    dlib.set_option("DFB_GMRES_CHECK", 20); dlib.set_option("DFB_GIVENS_DEFER", 1); dlib.set_option("DFB_GRAPH", 1)
    fs.set_preconditioner("schur2")
    dx.zero_()
    fs.krylov_solve(dx, F)
    fs.set_preconditioner("jacobi")
    state = [torch.from_numpy(a.copy()).cuda() for a in boxmesh.state_initial(mesh)]
    h = fs.time_step(*state)
    torch.cuda.synchronize()
    print("ok", N, mesh.num_tet, it, len(h))
    fs.close()
