"""TMA-ring SpMV (k_spmv_tma) against the register-staged kernel (k_spmv_fs) and timing of both, one GPU.
usage: python scripts/spmv_check.py [m ...]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import delaunay_mesh, shuffled_mesh  # noqa: E402
from dedflow_b200 import api, boxmesh, lib as dlib  # noqa: E402


def run(mesh, name, reps=0):
    N = mesh.num_node
    fs = api.FlowSystem(mesh)
    wg, dwg = (torch.from_numpy(a).cuda() for a in boxmesh.state_random(N))
    fs.assemble_system(wg, dwg, J=True)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(6 * N, dtype=torch.float64, device="cuda", generator=g)
    out = {}
    for tma in (0, 1, 2):
        dlib.set_option("DFB_SPMV_TMA", tma)
        y = torch.full((6 * N,), 3.0, dtype=torch.float64, device="cuda")
        fs.matrix_amvpby(1.0, x, 0.0, y)
        y2 = torch.full((6 * N,), 3.0, dtype=torch.float64, device="cuda")
        fs.matrix_amvpby(-0.5, x, 2.0, y2)
        torch.cuda.synchronize()
        out[tma] = (y.clone(), y2.clone())
        if reps:
            flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device="cuda")
            ts = []
            for _ in range(reps):
                flush.add_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fs.matrix_amvpby(1.0, x, 0.0, y); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            tb = []
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                fs.matrix_amvpby(1.0, x, 0.0, y)
            b.record(); torch.cuda.synchronize()
            print(f"  {name} tma={tma}: flushed-L2 {np.median(ts) * 1e3:7.1f} us   back-to-back {a.elapsed_time(b) / 20 * 1e3:7.1f} us", flush=True)
    for v in (1, 2):
        d1 = float((out[v][0] - out[0][0])[:4 * N].abs().max() / out[0][0][:4 * N].abs().max())
        d2 = float((out[v][1] - out[0][1])[:4 * N].abs().max() / out[0][1][:4 * N].abs().max())
        tail_ok = bool((out[v][0][4 * N:] == 3.0).all()) and bool((out[v][1][4 * N:] == 3.0).all())
        print(f"{name}: N={N} tma={v} vs register kernel rel diff {d1:.2e} / {d2:.2e} (beta != 0), tail untouched {tail_ok}", flush=True)
        assert d1 < 1e-13 and d2 < 1e-13 and tail_ok
    fs.close()


if __name__ == "__main__":
    run(boxmesh.make_box(1), "box m=1")
    run(boxmesh.make_box(3), "box m=3")
    run(shuffled_mesh(7), "shuffled m=7")
    run(delaunay_mesh(), "delaunay")
    run(boxmesh.make_box(20), "box m=20", reps=10)
    for m in [int(a) for a in sys.argv[1:]] or [55]:
        run(boxmesh.make_box(m), f"box m={m}", reps=20)
