"""A/B of the Krylov solve at BASELINE configs[1] under environment settings, one GPU.
usage: python scripts/solve_ab.py m "ENV=val,ENV2=val" "ENV=val" ...   (each argument is one configuration; "-" = defaults)"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dedflow_b200 import api, boxmesh, lib as dlib  # noqa: E402

m = int(sys.argv[1])
cfgs = sys.argv[2:] or ["-"]
mesh = boxmesh.make_box(m)
N = mesh.num_node
wg, dwg = boxmesh.state_random(N)
d_wg, d_dwg = torch.from_numpy(wg).cuda(), torch.from_numpy(dwg).cuda()
fs = api.FlowSystem(mesh)
F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
fs.assemble_system(d_wg, d_dwg, F=F)
fs.assemble_system(d_wg, d_dwg, J=True)
dx = torch.zeros_like(F)
base = None
for cfg in cfgs:
    keys = []
    if cfg != "-":
        for kv in cfg.split(","):
            k, v = kv.split("=")
            dlib.set_option(k, v)
            keys.append(k)
    ts = []
    for i in range(8):
        dx.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        it, hist = fs.krylov_solve(dx, F)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    sol = dx.clone()
    if base is None:
        base = sol
    d = float((sol - base).abs().max() / base.abs().max())
    print(f"{cfg:40s} solve {np.median(ts):7.3f} ms  ({it} its, {np.median(ts) / it * 1e3:6.1f} us/it)  diff vs first {d:.1e}", flush=True)
    # (options stay as set: list every switch explicitly in each configuration)
fs.close()
