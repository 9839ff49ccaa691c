"""Launch the Jacobian assembly a few times on the BASELINE configs[1] mesh (for `ncu -k regex:k_pairJ`)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dedflow_b200 import api, boxmesh  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mesh = boxmesh.make_box(m)
N = mesh.num_node
wg, dwg = boxmesh.state_random(N)
d_wg, d_dwg = torch.from_numpy(wg).cuda(), torch.from_numpy(dwg).cuda()
fs = api.FlowSystem(mesh)
P = lambda t: C.c_void_p(t.data_ptr())
for _ in range(4):
    assert fs.L.dfb_assemble_tet(fs.plan, P(fs.xg), P(d_wg), P(d_dwg), None, P(fs.A00), P(fs.A01), P(fs.A10), P(fs.A11), 1, 1, fs._stream()) == 0
torch.cuda.synchronize()
print("done")
