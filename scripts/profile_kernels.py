"""Launch each hot kernel a few times on the BASELINE configs[1] mesh (for `ncu`): J gather, F gather, SpMV, and a
20-iteration GMRES (multi-dot / update at growing column counts)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dedflow_b200 import api, boxmesh  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mode = sys.argv[2] if len(sys.argv) > 2 else "gather"
mesh = boxmesh.make_box(m)
N = mesh.num_node
maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 20
fs = api.FlowSystem(mesh, max_iter=maxit, atol=0.0, rtol=0.0)
wg, dwg = boxmesh.state_random(N)
d_wg, d_dwg = torch.from_numpy(wg).cuda(), torch.from_numpy(dwg).cuda()
F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
for _ in range(2):
    fs.assemble_system(d_wg, d_dwg, F=F, mode=mode)
    fs.assemble_system(d_wg, d_dwg, J=True, mode=mode)
x = torch.randn(6 * N, dtype=torch.float64, device="cuda")
y = torch.zeros_like(x)
for _ in range(2):
    fs.matrix_matvec(x, y)
dx = torch.zeros_like(F)
it, hist = fs.krylov_solve(dx, F)
torch.cuda.synchronize()
print("profiled", m, mode, it, hist[-1] / hist[0])
