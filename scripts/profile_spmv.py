"""A few launches of every SpMV variant on the BASELINE configs[1] mesh (for ncu)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dedflow_b200 import api, boxmesh, lib as dlib  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 55
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2]
mesh = boxmesh.make_box(m)
N = mesh.num_node
fs = api.FlowSystem(mesh)
wg, dwg = (torch.from_numpy(a).cuda() for a in boxmesh.state_random(N))
fs.assemble_system(wg, dwg, J=True)
x = torch.randn(6 * N, dtype=torch.float64, device="cuda")
y = torch.zeros_like(x)
for v in variants:
    dlib.set_option("DFB_SPMV_TMA", v)
    for _ in range(3):
        fs.matrix_matvec(x, y)
torch.cuda.synchronize()
print("done")
