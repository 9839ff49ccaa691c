"""Multi-GPU parity: run under torchrun with N GPUs.  Every rank assembles + solves its slab; the gathered F and dx are
compared on rank 0 with the single-domain CPU oracle (m small) -- tolerance 1e-10 relative (summation order differs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dist_check.py 12
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from dedflow_b200 import boxmesh, dist as ddist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("m", type=int, nargs="?", default=12)
args = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
mesh = boxmesh.make_box(args.m)
Ng = mesh.num_node
lm = ddist.partition(mesh, ddist.slab_owner(mesh, world), rank, world)
wg_g, dwg_g = boxmesh.state_random(Ng)
fs = ddist.DistFlowSystem(lm, f"cuda:{lr}")
N = fs.N
d_wg = torch.from_numpy(lm.localize(wg_g)).cuda()
d_dwg = torch.from_numpy(lm.localize(dwg_g)).cuda()
F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
dx = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
fs.assemble_system(d_wg, d_dwg, F=F)
fs.assemble_system(d_wg, d_dwg, J=True)
it, hist = fs.krylov_solve(dx, F)
torch.cuda.synchronize()
Fg = np.zeros(6 * Ng)
dxg = np.zeros(6 * Ng)
lm.scatter_owned(F.cpu().numpy(), Fg)
lm.scatter_owned(dx.cpu().numpy(), dxg)
tF, tx = torch.from_numpy(Fg).cuda(), torch.from_numpy(dxg).cuda()
dist.all_reduce(tF)
dist.all_reduce(tx)
# ghost consistency of the solution after the final halo
full = lm.localize(tx.cpu().numpy())
ghost_err = float(np.abs(full[:4 * N] - dx.cpu().numpy()[:4 * N]).max())
ok = True
if rank == 0:
    from oracle import pyoracle
    from test_gpu_parity import oracle_system
    O = pyoracle.get()
    ref = oracle_system(O, mesh, wg_g, dwg_g)
    xo, ito, histo = O.gmres(ref["pattern"], ref["blocks"], ref["F"])
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    eF = rel(tF.cpu().numpy(), ref["F"])
    ex = rel(tx.cpu().numpy()[:4 * Ng], xo[:4 * Ng])
    eh = float(np.abs(hist - histo).max() / histo[0]) if it == ito else float("inf")
    ok = eF <= 1e-12 and ex <= 1e-10 and eh <= 1e-10 and it == ito and ghost_err <= 1e-12 * np.abs(xo).max()
    print(f"dist_check world={world} m={args.m}: iters {it} (oracle {ito})  F rel {eF:.2e}  dx rel {ex:.2e}  hist rel {eh:.2e}  "
          f"ghost {ghost_err:.2e}  -> {'OK' if ok else 'FAIL'}", flush=True)
# ---- two time steps (predictor, Newton with reassembly, corrector) against the single-domain oracle driver ----
wg0 = [a.copy() for a in boxmesh.state_initial(mesh)]
d = [torch.from_numpy(lm.localize(a)).cuda() for a in wg0]
ok2 = True
ctx = None
if rank == 0:
    ctx = O.driver_setup(mesh)
for step in range(2):
    gh = fs.time_step(*d)
    gathered = []
    for v in d:
        g = np.zeros(6 * Ng)
        lm.scatter_owned(v.cpu().numpy(), g)
        t = torch.from_numpy(g).cuda()
        dist.all_reduce(t)
        gathered.append(t.cpu().numpy())
    if rank == 0:
        oh = O.time_step(ctx, *wg0)
        same = len(gh) == len(oh) and all(gi == oi for (_, gi), (_, oi) in zip(gh, oh))
        en = max(float(np.abs(gr - orr).max()) for (gr, _), (orr, _) in zip(gh, oh)) / oh[0][0].max() if same else float("inf")
        es = max(float(np.abs(g[:4 * Ng] - w[:4 * Ng]).max() / max(np.abs(w[:4 * Ng]).max(), 1e-12)) for g, w in zip(gathered, wg0))
        ok2 = ok2 and same and en <= 1e-8 and es <= 1e-8
        print(f"dist_check time step {step + 1}: Newton its {len(gh) - 1} (oracle {len(oh) - 1})  norms rel {en:.2e}  state rel {es:.2e}  "
              f"-> {'OK' if ok2 else 'FAIL'}", flush=True)
ok = ok and ok2
fs.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
