"""A few launches of the residual-assembly variants on the BASELINE configs[1] mesh (for ncu)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dedflow_b200 import api, boxmesh, lib as dlib  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mesh = boxmesh.make_box(m)
N = mesh.num_node
P = lambda t: C.c_void_p(t.data_ptr())
wg, dwg = (torch.from_numpy(a).cuda() for a in boxmesh.state_random(N))
fs = api.FlowSystem(mesh)
F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
for opts in ({"DFB_F_VARIANT": "patch", "DFB_F_PATCH_CTAS": "2"},):
    for k, v in opts.items():
        dlib.set_option(k, v)
    for _ in range(3):
        assert fs.L.dfb_assemble_tet(fs.plan, P(fs.xg), P(wg), P(dwg), P(F), None, None, None, None, 1, 1, fs._stream()) == 0
torch.cuda.synchronize()
print("done")
