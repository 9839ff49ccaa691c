"""Per-launch kernel times of one data-parallel solve (weak-scaling mesh of bench.py), rank 0's view.  Run under torchrun.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 scripts/dist_profile.py"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from dedflow_b200 import boxmesh, dist as ddist, lib as dlib  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
m = ddist.weak_scaling_m(55, world)
mesh = boxmesh.make_box(m)
lm = ddist.partition(mesh, ddist.slab_owner(mesh, world), rank, world)
wg_g, dwg_g = boxmesh.state_random(mesh.num_node)
fs = ddist.DistFlowSystem(lm, f"cuda:{lr}", max_iter=40, atol=0.0, rtol=0.0)
N = fs.N
d_wg, d_dwg = torch.from_numpy(lm.localize(wg_g)).cuda(), torch.from_numpy(lm.localize(dwg_g)).cuda()
F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
dx = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
fs.assemble_system(d_wg, d_dwg, F=F)
fs.assemble_system(d_wg, d_dwg, J=True)


def timed(label, opts):
    """median event time of 7 graph-replayed solves under `opts` (rank 0 prints)"""
    for k, v in opts.items():
        dlib.set_option(k, v)
    ts = []
    for i in range(10):
        dx.zero_()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _, hist = fs.krylov_solve(dx, F)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    if rank == 0:
        print(f"[variant] {label:28s} solve {sorted(ts)[len(ts) // 2]:7.3f} ms (min {min(ts):7.3f})  final residual {hist[-1]:.6e}", flush=True)


for label, opts in (("flag from the update", {"DFB_HALO_DEFER": "0"}), ("flag from the next mat-vec", {"DFB_HALO_DEFER": "1"}),
                    ("flag from the update", {"DFB_HALO_DEFER": "0"}), ("flag from the next mat-vec", {"DFB_HALO_DEFER": "1"})):
    timed(label, opts)
if len(sys.argv) > 1:
    dlib.set_option("DFB_HALO_DEFER", sys.argv[1])
torch.cuda.synchronize()
dist.barrier()
dlib.set_option("DFB_PROFILE", 2 if rank == 0 else -1)
dx.zero_()
fs.krylov_solve(dx, F)
torch.cuda.synchronize()
dist.barrier()
fs.close()
dist.destroy_process_group()
