"""Run the reference's own CUDA build (oracle/_ref) on a GPU box, compare the CPU oracle and the B200 path with it,
and write golden fixtures.

    python oracle/ref/run_ref.py --m 6 --golden tests/golden/ref_m6.npz      (full arrays, small mesh)
    python oracle/ref/run_ref.py --m 20 --out gpurun_out/ref_m20.npz --time  (summary + timings)

TEST INFRASTRUCTURE ONLY.  /root/reference is not needed at run time: the library was built beforehand.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from dedflow_b200 import boxmesh  # noqa: E402
from oracle import pyoracle  # noqa: E402
from oracle.ref import reflib  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def ev_time(fn, reps):
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=6)
    ap.add_argument("--shuffle", action="store_true")
    ap.add_argument("--delaunay", action="store_true", help="unstructured mesh (tests/conftest.py::delaunay_mesh); the mesh is stored in the golden file")
    ap.add_argument("--state", default="B")
    ap.add_argument("--golden", default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--ours", action="store_true", help="also compare the B200 path")
    ap.add_argument("--no-d1-patch", action="store_true")
    args = ap.parse_args()

    if args.delaunay:
        from conftest import delaunay_mesh
        mesh = delaunay_mesh()
        args.m = 0
    elif args.shuffle:
        from conftest import shuffled_mesh
        mesh = shuffled_mesh(args.m)
    else:
        mesh = boxmesh.make_box(args.m)
    N, E = mesh.num_node, mesh.num_tet
    wg, dwg = boxmesh.state_random(N) if args.state == "B" else boxmesh.state_default(mesh)
    R = reflib.RefProblem(mesh, patch_d1=not args.no_d1_patch)
    report = {"m": args.m, "N": N, "E": E, "state": args.state, "shuffle": args.shuffle, "d1_patched": R.patched_d1}

    pats = R.patterns()
    color, boff, bind, nc = R.color_batches()
    d_wg, d_dwg = torch.from_numpy(wg).cuda(), torch.from_numpy(dwg).cuda()
    F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    R.assemble(d_wg, d_dwg, F_t=F)
    R.assemble(d_wg, d_dwg, J=True)
    torch.cuda.synchronize()
    Fh = F.cpu().numpy()
    blocks = R.block_vals()
    rng = np.random.default_rng(3)
    x = rng.standard_normal(6 * N)
    y = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    R.matvec(torch.from_numpy(x).cuda(), y)
    torch.cuda.synchronize()
    yh = y.cpu().numpy()
    # GMRES: full solve + truncated runs to obtain the residual at every 20th iteration to full precision
    sols, hists = {}, {}
    for k in (20, 40, 60, 80, 100, 120):
        dx = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
        h = R.solve(dx, F, maxit=k, atol=0.0, rtol=0.0)
        sols[k] = dx.cpu().numpy()
        hists[k] = h
    dx = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    printed = R.solve(dx, F)
    dxh = dx.cpu().numpy()
    report["ref_gmres_printed"] = printed

    # ---------------- oracle vs reference ----------------
    O = pyoracle.get()
    rp, ci = O.nodal_pattern(N, mesh.ien)
    cmp = {}
    cmp["row_ptr"] = bool(np.array_equal(rp, pats["1x1"][0]))
    cmp["col_ind"] = bool(np.array_equal(ci, pats["1x1"][1]))
    for name, (br, bc) in {"3x3": (3, 3), "3x1": (3, 1), "1x3": (1, 3)}.items():
        orp, oci = O.expand_block(rp, ci, br, bc, fix_last=R.patched_d1)
        cmp[f"row_ptr_{name}"] = bool(np.array_equal(orp, pats[name][0]))
        cmp[f"col_ind_{name}"] = bool(np.array_equal(oci, pats[name][1]))
    w = O.weights(pyoracle.curand_host_u32(E))
    ocolor, orounds, ties = O.color_jpl(N, mesh.ien, w)
    ooff, oind = O.color_batches(ocolor)
    cmp["ties"] = ties
    cmp["num_color"] = [int(nc), int(orounds)]
    cmp["color"] = bool(np.array_equal(ocolor, color))
    cmp["batch_offset"] = bool(np.array_equal(ooff, boff))
    cmp["batch_ind"] = bool(np.array_equal(oind, bind))
    Z = ci.size
    oF = np.zeros(6 * N)
    ob = [np.zeros(9 * Z), np.zeros(3 * Z), np.zeros(3 * Z), np.zeros(Z)]
    O.assemble_tet(N, mesh.ien, mesh.xg, ooff, oind, wg, dwg, F=oF)
    O.assemble_tet(N, mesh.ien, mesh.xg, ooff, oind, wg, dwg, pattern=(rp, ci), blocks=ob)
    f2e, forn = mesh.bound_faces(4)
    O.assemble_face(f2e, forn, N, mesh.ien, mesh.xg, ocolor, orounds, wg, dwg, F=oF)
    O.assemble_face(f2e, forn, N, mesh.ien, mesh.xg, ocolor, orounds, wg, dwg, pattern=(rp, ci), blocks=ob)
    oF[4 * N:] = 0
    for b, t in {0: (1, 1, 1), 2: (0, 1, 0), 3: (0, 0, 1), 4: (0, 0, 0)}.items():
        O.dirichlet_vec(mesh.bound_nodes(b), np.array(t, np.int32), oF)
        O.dirichlet_mat(mesh.bound_nodes(b), np.array(t, np.int32), N, (rp, ci), ob[0], ob[1])
    cmp["F_rel"] = rel(oF, Fh)
    for nme, a, b in zip(("A00", "A01", "A10", "A11"), ob, blocks):
        cmp[f"{nme}_rel"] = rel(a, b)
    oy = np.zeros(6 * N)
    O.fs_amvpby((rp, ci), blocks, 1.0, x, 0.0, oy)
    cmp["matvec_rel"] = rel(oy[:4 * N], yh[:4 * N])
    cmp["matvec_tail_untouched"] = bool(np.all(yh[4 * N:] == 0))
    ox, oit, ohist = O.gmres((rp, ci), blocks, Fh)
    cmp["gmres_iters"] = [int(oit), int(printed[-1][0]) if printed else -1]
    cmp["gmres_x_rel"] = rel(ox[:4 * N], dxh[:4 * N])
    # residual at 20,40,..: true residual of the truncated reference solves vs the oracle's |beta_k|
    res_true = {}
    for k, s in sols.items():
        r = np.zeros(6 * N)
        O.fs_amvpby((rp, ci), blocks, 1.0, s, 0.0, r)
        res_true[k] = float(np.linalg.norm(Fh[:4 * N] - r[:4 * N]))
    ox120, _, oh120 = O.gmres((rp, ci), blocks, Fh, atol=0.0, rtol=0.0)
    cmp["gmres_hist_rel_to_r0"] = {str(k): abs(res_true[k] - oh120[k]) / oh120[0] for k in res_true}
    cmp["gmres_printed_vs_oracle"] = [[it, v, float(ohist[it]) if it < len(ohist) else None] for it, v in printed]
    report["oracle_vs_reference"] = cmp

    # ---------------- ours vs reference ----------------
    if args.ours:
        from dedflow_b200 import api
        fs = api.FlowSystem(mesh)
        oc = {}
        oc["row_ptr"] = bool(np.array_equal(fs.row_ptr.cpu().numpy(), pats["1x1"][0]))
        oc["col_ind"] = bool(np.array_equal(fs.col_ind.cpu().numpy(), pats["1x1"][1]))
        for name, (br, bc) in {"3x3": (3, 3), "3x1": (3, 1), "1x3": (1, 3)}.items():
            nrp, nci = fs.csr_attr_create_block(br, bc)
            n = pats[name][0].size - (0 if R.patched_d1 else 1)
            oc[f"row_ptr_{name}"] = bool(np.array_equal(nrp.cpu().numpy()[:n], pats[name][0][:n]))
            oc[f"col_ind_{name}"] = bool(np.array_equal(nci.cpu().numpy(), pats[name][1]))
        oc["color"] = bool(np.array_equal(fs.color.cpu().numpy(), color))
        oc["batch_offset"] = bool(np.array_equal(fs.batch_offset, boff))
        oc["batch_ind"] = bool(np.array_equal(fs.batch_ind.cpu().numpy(), bind))
        for mode in ("gather", "atomic", "colored"):
            F2 = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
            fs.assemble_system(d_wg, d_dwg, F=F2, mode=mode)
            fs.assemble_system(d_wg, d_dwg, J=True, mode=mode)
            oc[f"F_rel_{mode}"] = rel(F2.cpu().numpy(), Fh)
            for nme, a, b in zip(("A00", "A01", "A10", "A11"), fs.blocks(), blocks):
                oc[f"{nme}_rel_{mode}"] = rel(a.cpu().numpy(), b)
        for a, b in zip(fs.blocks(), blocks):
            a.copy_(torch.from_numpy(b))
        y2 = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
        fs.matrix_matvec(torch.from_numpy(x).cuda(), y2)
        oc["matvec_rel"] = rel(y2.cpu().numpy()[:4 * N], yh[:4 * N])
        dx2 = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
        it2, hist2 = fs.krylov_solve(dx2, F)
        oc["gmres_iters"] = int(it2)
        oc["gmres_x_rel"] = rel(dx2.cpu().numpy()[:4 * N], dxh[:4 * N])
        fs.atol = fs.rtol = 0.0
        dx3 = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
        it3, hist3 = fs.krylov_solve(dx3, F)
        oc["gmres_hist_rel_to_r0"] = {str(k): abs(res_true[k] - hist3[k]) / hist3[0] for k in res_true}
        oc["gmres_x120_rel"] = rel(dx3.cpu().numpy()[:4 * N], sols[120][:4 * N])
        report["ours_vs_reference"] = oc

    # ---------------- timings of the reference's entry points ----------------
    if args.time:
        tm = {}
        tm["assemble_F_ms"] = ev_time(lambda: R.assemble(d_wg, d_dwg, F_t=F), 5)
        tm["assemble_J_ms"] = ev_time(lambda: R.assemble(d_wg, d_dwg, J=True), 5)
        xs = torch.from_numpy(x).cuda()
        tm["matvec_ms"] = ev_time(lambda: R.matvec(xs, y), 20)

        def solve():
            d = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
            R.solve(d, F)
        R.assemble(d_wg, d_dwg, F_t=F)
        tm["krylov_solve_ms"] = ev_time(solve, 3)
        t0 = time.time()
        reflib.RefProblem.color_batches(R)
        tm["coloring_wall_ms"] = (time.time() - t0) * 1e3
        report["reference_cuda_timings"] = tm

    print(json.dumps(report, indent=1))
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).with_suffix(".json").write_text(json.dumps(report, indent=1))
    if args.golden:
        Path(args.golden).parent.mkdir(parents=True, exist_ok=True)
        np.savez_compressed(
            args.golden, m=args.m, shuffle=args.shuffle, state=args.state, d1_patched=R.patched_d1,
            row_ptr=pats["1x1"][0], col_ind=pats["1x1"][1],
            row_ptr_3x3=pats["3x3"][0], col_ind_3x3=pats["3x3"][1], row_ptr_3x1=pats["3x1"][0], col_ind_3x1=pats["3x1"][1],
            row_ptr_1x3=pats["1x3"][0], col_ind_1x3=pats["1x3"][1],
            color=color, batch_offset=boff, batch_ind=bind,
            F=Fh, A00=blocks[0], A01=blocks[1], A10=blocks[2], A11=blocks[3], x=x, y=yh, dx=dxh,
            res_iters=np.array(sorted(res_true)), res_true=np.array([res_true[k] for k in sorted(res_true)]),
            dx120=sols[120], printed=np.array(printed, dtype=np.float64).reshape(-1, 2),
            **({"mesh_xg": mesh.xg, "mesh_ien": mesh.ien, "mesh_bound_node_offset": mesh.bound_node_offset,
                "mesh_bound_node": mesh.bound_node, "mesh_bound_elem_offset": mesh.bound_elem_offset,
                "mesh_bound_f2e": mesh.bound_f2e, "mesh_bound_forn": mesh.bound_forn, "mesh_bound_ien": mesh.bound_ien}
               if args.delaunay else {}))


if __name__ == "__main__":
    main()
