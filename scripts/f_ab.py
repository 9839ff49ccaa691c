"""A/B of the residual (F) assembly variants on one GPU: correctness against the scratch variant + kernel times.
usage: python scripts/f_ab.py [m] """
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import delaunay_mesh, shuffled_mesh  # noqa: E402
from dedflow_b200 import api, boxmesh, lib as dlib  # noqa: E402

P = lambda t: C.c_void_p(t.data_ptr())
CFGS = [("scratch", {"DFB_F_VARIANT": "scratch"}), ("patch 3 CTAs/SM", {"DFB_F_VARIANT": "patch", "DFB_F_PATCH_CTAS": "3"}),
        ("patch 2 CTAs/SM", {"DFB_F_VARIANT": "patch", "DFB_F_PATCH_CTAS": "2"}), ("pipe", {"DFB_F_VARIANT": "pipe"})]


def run(mesh, name, reps=0):
    N = mesh.num_node
    wg, dwg = (torch.from_numpy(a).cuda() for a in boxmesh.state_random(N))
    ref = None
    for cname, opts in CFGS:
        for k, v in opts.items():
            dlib.set_option(k, v)
        fs = api.FlowSystem(mesh)
        st = fs._stream()
        F = torch.full((6 * N,), 7.0, dtype=torch.float64, device="cuda")
        call = lambda ow=1: fs.L.dfb_assemble_tet(fs.plan, P(fs.xg), P(wg), P(dwg), P(F), None, None, None, None, 1, ow, st)
        assert call() == 0, fs.L.dfb_last_error()
        torch.cuda.synchronize()
        got = F.clone()
        assert call(0) == 0                    # accumulate on top: 2x
        d_acc = float((F - 2.0 * got).abs().max() / got.abs().max())
        if ref is None:
            ref, d = got, 0.0
        else:
            d = float((got - ref).abs().max() / ref.abs().max())
        line = f"{name:14s} {cname:16s} rel diff vs scratch {d:.2e}  accumulate {d_acc:.1e}"
        if reps:
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); call(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            line += f"   F kernels {np.median(ts) * 1e3:7.1f} us   plan {fs.L.dfb_plan_bytes(fs.plan) / 1e6:.0f} MB"
        print(line, flush=True)
        assert d <= 1e-13 and d_acc <= 1e-15
        fs.close()


if __name__ == "__main__":
    dlib.set_option("DFB_VERBOSE", "1")
    run(boxmesh.make_box(1), "box m=1")
    run(boxmesh.make_box(4), "box m=4")
    run(shuffled_mesh(9), "shuffled m=9")
    run(delaunay_mesh(), "delaunay")
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 55
    run(boxmesh.make_box(m), f"box m={m}", reps=20)
