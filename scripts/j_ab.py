"""A/B of the Jacobian assembly variants on one GPU: kernel time (CUDA events, median of 20) and agreement with the pull variant.
usage: python scripts/j_ab.py [m] [shuffle]"""
import ctypes as C
import os
os.environ["DFB_VERBOSE"] = "1"
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dedflow_b200 import api, boxmesh, lib as dlib  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mesh = boxmesh.make_box(m)
if len(sys.argv) > 2 and sys.argv[2] == "shuffle":      # random node and element numbering (timing only: boundary lists go stale)
    rng = np.random.default_rng(1)
    perm = rng.permutation(mesh.num_node).astype(np.int32)
    xg = np.empty_like(mesh.xg.reshape(-1, 3))
    xg[perm] = mesh.xg.reshape(-1, 3)
    mesh.xg = np.ascontiguousarray(xg.reshape(mesh.xg.shape))
    mesh.ien = np.ascontiguousarray(perm[mesh.ien][rng.permutation(mesh.num_tet)])
N, E = mesh.num_node, mesh.num_tet
wg, dwg = boxmesh.state_random(N)
d_wg, d_dwg = torch.from_numpy(wg).cuda(), torch.from_numpy(dwg).cuda()
P = lambda t: C.c_void_p(t.data_ptr())
ref = None
shuffle = len(sys.argv) > 2 and sys.argv[2] == "shuffle"
CFGS = [("pairs 96x4", {"DFB_J_VARIANT": "pairs", "DFB_J_PAIR_ORDER": "morton", "DFB_J_PAIR_ROWS": "8", "DFB_J_PAIR_NT": "96"}),
        ("pairs 128x3", {"DFB_J_VARIANT": "pairs", "DFB_J_PAIR_ORDER": "morton", "DFB_J_PAIR_ROWS": "8", "DFB_J_PAIR_NT": "128"}),
        ("pairs R16 224x2", {"DFB_J_VARIANT": "pairs", "DFB_J_PAIR_ORDER": "morton", "DFB_J_PAIR_ROWS": "16", "DFB_J_PAIR_NT": "224"}),
        ("pairs R16 192x2", {"DFB_J_VARIANT": "pairs", "DFB_J_PAIR_ORDER": "morton", "DFB_J_PAIR_ROWS": "16", "DFB_J_PAIR_NT": "192"})]
OLD_CFGS = [("pull", {"DFB_J_VARIANT": "pull"}),
        ("pairs morton", {"DFB_J_VARIANT": "pairs", "DFB_J_PAIR_ORDER": "morton"}),
        ("pairs natural", {"DFB_J_VARIANT": "pairs", "DFB_J_PAIR_ORDER": "natural"}),
        ("pairs R=16", {"DFB_J_VARIANT": "pairs", "DFB_J_PAIR_ORDER": "morton", "DFB_J_PAIR_ROWS": "16"}),
        ("fused", {"DFB_J_VARIANT": "fused"})]
for name, env in CFGS:
    for k_, v_ in env.items():
        dlib.set_option(k_, v_)
    fs = api.FlowSystem(mesh)
    st = fs._stream()
    call = lambda: fs.L.dfb_assemble_tet(fs.plan, P(fs.xg), P(d_wg), P(d_dwg), None, P(fs.A00), P(fs.A01), P(fs.A10), P(fs.A11), 1, 1, st)
    for _ in range(3):
        assert call() == 0, fs.L.dfb_last_error()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    vals = [x.clone() for x in fs.blocks()]
    if ref is None:
        ref = vals
        d = 0.0
    else:
        d = max(float((x - y).abs().max() / y.abs().max()) for x, y in zip(vals, ref))
    Fv = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    callF = lambda: fs.L.dfb_assemble_tet(fs.plan, P(fs.xg), P(d_wg), P(d_dwg), P(Fv), None, None, None, None, 1, 1, st)
    tf = []
    for k in range(13):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); assert callF() == 0; b.record()
        torch.cuda.synchronize()
        if k >= 3:
            tf.append(a.elapsed_time(b))
    print(f"   F kernels {np.median(tf) * 1e3:8.1f} us", flush=True)
    print(f"{name:12s} m={m} E={E}: J kernels {np.median(ts) * 1e3:8.1f} us   max rel diff vs first {d:.2e}   plan {fs.L.dfb_plan_bytes(fs.plan) / 1e6:.0f} MB", flush=True)
    fs.close()
