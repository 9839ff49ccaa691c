"""CPU-side tests (pytest -m "not gpu"): the oracle against structural properties and known answers, the element
arithmetic of the kernels (probed on the host) against the oracle, and the C-ABI surface of the library."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, delaunay_mesh, shuffled_mesh
from dedflow_b200 import boxmesh
from oracle import pyoracle


def test_box_mesh_counts():
    m = boxmesh.make_box(20)
    assert m.num_tet == 48000 and m.num_node == 9261          # BASELINE config 1
    x = m.xg[m.ien]
    J = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0], x[:, 3] - x[:, 0]], axis=2)
    det = np.linalg.det(J)
    assert det.min() > 0 and abs(det.sum() / 6 - 1.0) < 1e-12  # positively oriented, fills the unit cube
    for b in range(6):
        assert len(m.bound_nodes(b)) == 21 * 21 and len(m.bound_faces(b)[0]) == 2 * 20 * 20


def test_curand_known_answer():
    # cuRAND host XORWOW, seed 1234 -- first draws and the weights derived from them
    raw = pyoracle.curand_host_u32(8)
    assert raw.tolist() == [624778773, 3522650202, 2363946744, 1266286439, 3928747533, 3732235839, 1382638835, 3362343509]
    w = pyoracle.get().weights(raw)
    assert w.tolist() == [624778773, 301424733, 216463098, 192544616, 707522064, 511010370, 308897012, 141118040]


def test_pattern_properties(oracle):
    m = boxmesh.make_box(5)
    rp, ci = oracle.nodal_pattern(m.num_node, m.ien)
    # independent construction with numpy sets
    adj = [set([i]) for i in range(m.num_node)]
    for e in m.ien:
        for a in e:
            adj[a].update(e.tolist())
    for i in range(m.num_node):
        assert ci[rp[i]:rp[i + 1]].tolist() == sorted(adj[i])
    assert np.diff(rp).max() == 15
    for br, bc in ((3, 3), (3, 1), (1, 3)):
        nrp, nci = oracle.expand_block(rp, ci, br, bc)
        assert nrp[-1] == ci.size * br * bc and np.all(np.diff(nrp) >= 0)
        # scalar row i*br+j holds columns col*bc+l, k outer / l inner
        i, j = 17, br - 1
        row = nci[nrp[i * br + j]:nrp[i * br + j + 1]]
        want = (ci[rp[i]:rp[i + 1]][:, None] * bc + np.arange(bc)[None, :]).ravel()
        assert row.tolist() == want.tolist()
        nrp0, _ = oracle.expand_block(rp, ci, br, bc, fix_last=False)
        assert nrp0[-1] == 0 and np.array_equal(nrp0[:-1], nrp[:-1])      # defect D1


def test_coloring_valid_and_is_dag_depth(oracle):
    m = boxmesh.make_box(6)
    w = oracle.weights(pyoracle.curand_host_u32(m.num_tet))
    color, rounds, ties = oracle.color_jpl(m.num_node, m.ien, w)
    assert ties == 0 and color.min() == 0 and color.max() + 1 == rounds
    for c in range(rounds):
        nodes = m.ien[color == c].ravel()
        assert len(np.unique(nodes)) == len(nodes)
    # JPL on tie-free weights == longest-path depth in the weight-oriented conflict graph
    rp, cidx = oracle.v2e(m.num_node, m.ien)
    order = np.argsort(-w.astype(np.int64))
    depth = np.full(m.num_tet, -1)
    for e in order:
        nb = np.unique(np.concatenate([cidx[rp[n]:rp[n + 1]] for n in m.ien[e]]))
        nb = nb[(nb != e) & (w[nb] > w[e])]
        depth[e] = 0 if nb.size == 0 else depth[nb].max() + 1
    assert np.array_equal(depth, color)
    off, ind = oracle.color_batches(color)
    assert off[-1] == m.num_tet
    for c in range(rounds):
        seg = ind[off[c]:off[c + 1]]
        assert np.all(color[seg] == c) and np.all(np.diff(seg) > 0)


def _assemble_oracle(O, m, wg, dwg):
    N = m.num_node
    rp, ci = O.nodal_pattern(N, m.ien)
    Z = ci.size
    w = O.weights(pyoracle.curand_host_u32(m.num_tet))
    color, nc, _ = O.color_jpl(N, m.ien, w)
    off, ind = O.color_batches(color)
    F = np.zeros(6 * N)
    blocks = [np.zeros(9 * Z), np.zeros(3 * Z), np.zeros(3 * Z), np.zeros(Z)]
    O.assemble_tet(N, m.ien, m.xg, off, ind, wg, dwg, F=F)
    O.assemble_tet(N, m.ien, m.xg, off, ind, wg, dwg, pattern=(rp, ci), blocks=blocks)
    f2e, forn = m.bound_faces(4)
    O.assemble_face(f2e, forn, N, m.ien, m.xg, color, nc, wg, dwg, F=F)
    O.assemble_face(f2e, forn, N, m.ien, m.xg, color, nc, wg, dwg, pattern=(rp, ci), blocks=blocks)
    F[4 * N:] = 0
    for b, t in {0: (1, 1, 1), 2: (0, 1, 0), 3: (0, 0, 1), 4: (0, 0, 0)}.items():
        O.dirichlet_vec(m.bound_nodes(b), np.array(t, np.int32), F)
        O.dirichlet_mat(m.bound_nodes(b), np.array(t, np.int32), N, (rp, ci), blocks[0], blocks[1])
    return (rp, ci), F, blocks


def test_oracle_assembly_matches_dense_reassembly(oracle):
    """The CSR scatter + Dirichlet of the oracle against a dense re-assembly of the per-element checkpoints."""
    m = shuffled_mesh(3)
    N = m.num_node
    wg, dwg = boxmesh.state_random(N)
    (rp, ci), F, blocks = _assemble_oracle(oracle, m, wg, dwg)
    eF, eJ, _, _ = oracle.tet_elements(N, m.ien, m.xg, wg, dwg)
    D = np.zeros((4 * N, 4 * N))
    Fd = np.zeros(6 * N)

    def dof(n, i):
        return n * 3 + i if i < 3 else 3 * N + n

    def add(nodes, ef, ej):
        for a in range(4):
            for i in range(3):
                Fd[nodes[a] * 3 + i] += ef[a, i]
            Fd[3 * N + nodes[a]] += ef[a, 3]
            for b in range(4):
                for i in range(4):
                    for j in range(4):
                        D[dof(nodes[a], i), dof(nodes[b], j)] += ej[a, b, i, j]
    for e in range(m.num_tet):
        add(m.ien[e], eF[e], eJ[e])
    f2e, forn = m.bound_faces(4)
    fF, fJ = oracle.face_elements(f2e, forn, N, m.ien, m.xg, wg, dwg)
    for k, e in enumerate(f2e):
        add(m.ien[e], fF[k], fJ[k])
    for b, t in {0: (1, 1, 1), 2: (0, 1, 0), 3: (0, 0, 1)}.items():
        for n in m.bound_nodes(b):
            for ic in range(3):
                if t[ic]:
                    Fd[n * 3 + ic] = 0
                    D[n * 3 + ic, :] = 0
                    D[n * 3 + ic, n * 3 + ic] = 1
    assert np.abs(F - Fd).max() <= 1e-13 * np.abs(Fd).max()
    x = np.random.default_rng(0).standard_normal(6 * N)
    y = np.zeros(6 * N)
    oracle.fs_amvpby((rp, ci), blocks, 1.0, x, 0.0, y)
    yd = D @ x[:4 * N]
    assert np.abs(y[:4 * N] - yd).max() <= 1e-12 * np.abs(yd).max()
    assert np.all(y[4 * N:] == 0)                                   # defect D4: rows [4N,6N) untouched


def test_oracle_gmres_consistent(oracle):
    m = boxmesh.make_box(6)
    N = m.num_node
    wg, dwg = boxmesh.state_random(N)
    pat, F, blocks = _assemble_oracle(oracle, m, wg, dwg)
    x, it, hist = oracle.gmres(pat, blocks, F)
    assert it % 20 == 0 and it <= 120                                # defect D10
    y = np.zeros(6 * N)
    oracle.fs_amvpby(pat, blocks, 1.0, x, 0.0, y)
    true_res = np.linalg.norm(F[:4 * N] - y[:4 * N])
    assert abs(true_res - hist[-1]) <= 1e-8 * hist[0]
    assert np.all(np.diff(hist) <= 1e-12 * hist[0])                  # GMRES residuals are non-increasing
    # block-Jacobi applies (B^-1)^T (defect D3)
    d00, d11 = oracle.pc_setup(pat, blocks)
    rp, ci = pat
    i = 100
    k = rp[i] + np.searchsorted(ci[rp[i]:rp[i + 1]], i)
    ln = rp[i + 1] - rp[i]
    B = np.array([[blocks[0][rp[i] * 9 + (k - rp[i]) * 3 + r * ln * 3 + c] for c in range(3)] for r in range(3)])
    v = np.zeros(6 * N)
    v[3 * i:3 * i + 3] = [1.0, 2.0, 3.0]
    out = oracle.pc_apply(d00, d11, v)
    assert np.allclose(out[3 * i:3 * i + 3], np.linalg.inv(B).T @ [1.0, 2.0, 3.0], rtol=1e-10)


@pytest.fixture(scope="module")
def probe():
    src = ROOT / "tests" / "cpu_probe" / "probe.cpp"
    so = ROOT / "tests" / "cpu_probe" / "libprobe.so"
    hdr = ROOT / "dedflow_b200" / "csrc" / "elem_math.cuh"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(so), str(src)], check=True)
    return C.CDLL(str(so))


@pytest.mark.parametrize("kind", ["box", "delaunay"])
def test_kernel_element_math_matches_oracle(oracle, probe, kind):
    """The hoisted Jacobian / residual / face arithmetic the CUDA kernels use (elem_math.cuh, evaluated on the host)
    against the reference-ordered oracle: max|d| / max|A| <= 1e-12 per sub-block.  Box: every local vertex
    order; Delaunay: irregular geometry (slivers, element sizes over two orders of magnitude)."""
    m = shuffled_mesh(6) if kind == "box" else delaunay_mesh(150, 15)
    N, E = m.num_node, m.num_tet
    for wg, dwg in (boxmesh.state_random(N), boxmesh.state_default(m)):
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        eF = np.zeros((E, 4, 6))
        eJ = np.zeros((E, 4, 4, 4, 4))
        probe.probe_tet(E, N, vp(m.ien), vp(m.xg), vp(wg), vp(dwg), vp(eF), vp(eJ))
        oF, oJ, _, _ = oracle.tet_elements(N, m.ien, m.xg, wg, dwg)
        assert np.abs(eF - oF).max() <= 1e-12 * np.abs(oF).max()
        eP = np.zeros((E, 4, 4, 4, 4))     # the node-pair formulation of k_pairJ (element record -> pair accumulators)
        probe.probe_tet_pairs(E, N, vp(m.ien), vp(m.xg), vp(wg), vp(dwg), vp(eP))
        for i0, i1, j0, j1 in ((0, 3, 0, 3), (0, 3, 3, 4), (3, 4, 0, 3), (3, 4, 3, 4)):
            a, b = eJ[..., i0:i1, j0:j1], oJ[..., i0:i1, j0:j1]
            assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()
            assert np.abs(eP[..., i0:i1, j0:j1] - b).max() <= 1e-12 * np.abs(b).max()
        for b in range(6):
            f2e, forn = m.bound_faces(b)
            assert set(np.unique(forn)) == {0, 1, 2, 3}
            fF = np.zeros((len(f2e), 4, 6))
            fJ = np.zeros((len(f2e), 4, 4, 4, 4))
            probe.probe_face(len(f2e), vp(f2e), vp(forn), N, vp(m.ien), vp(m.xg), vp(wg), vp(dwg), vp(fF), vp(fJ))
            gF, gJ = oracle.face_elements(f2e, forn, N, m.ien, m.xg, wg, dwg)
            assert np.abs(fF - gF).max() <= 1e-12 * max(np.abs(gF).max(), 1e-300)
            assert np.abs(fJ - gJ[..., :4, :4]).max() <= 1e-12 * np.abs(gJ).max()


def test_oracle_driver_time_step(oracle):
    """numpy restatement of main.c's driver (SolveFlowSystem + one pass of the time loop): slot rules of the alpha-level
    states, Newton residuals decrease, phi/T residual blocks are discarded, pressure slot of wgold is never touched."""
    mesh = boxmesh.make_box(4)
    N = mesh.num_node
    rng = np.random.default_rng(0)
    wgold, dwgold, dwg = (rng.standard_normal(6 * N) for _ in range(3))
    wga, dwga = oracle.alpha_states(N, wgold, dwgold, dwg)
    assert np.all(wga[3 * N:4 * N] == 0) and np.array_equal(dwga[3 * N:4 * N], dwg[3 * N:4 * N])
    am, af = oracle.K_ALPHAM, oracle.K_ALPHAF
    assert np.allclose(dwga[:3 * N], (1 - am) * dwgold[:3 * N] + am * dwg[:3 * N], rtol=1e-14)
    assert np.allclose(wga[4 * N:], wgold[4 * N:] + oracle.K_DT * af * ((1 - oracle.K_GAMMA) * dwgold[4 * N:] + oracle.K_GAMMA * dwg[4 * N:]), rtol=1e-13)
    ctx = oracle.driver_setup(mesh)
    state = [a.copy() for a in boxmesh.state_initial(mesh)]
    p_before = state[0][3 * N:4 * N].copy()
    hist = oracle.time_step(ctx, *state)
    norms = np.array([h[0] for h in hist])
    assert 2 <= len(hist) <= 5 and all(h[1] % 20 == 0 and h[1] > 0 for h in hist[1:])
    assert norms[-1, 0] < 1e-2 * norms[0, 0]                       # Newton converges on the momentum block
    assert np.all(norms[:, 2:] == 0)                               # phi / T residuals are zeroed (main.c:63-66)
    assert np.array_equal(state[0][3 * N:4 * N], p_before)         # the corrector skips the pressure slot of wgold
    assert np.array_equal(state[1], state[2])                      # dwgold = dwg


def _declared(header):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    text = re.sub(r"typedef[^;{]*\(\s*\*[^;]*;", "", text)          # function-pointer typedefs
    text = re.sub(r"(struct|enum)\s+\w+\s*\{[^{}]*\}", "", text)    # struct / enum bodies
    return {d for d in re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text) if not d.startswith("(")}


def test_abi_library_exports_every_declared_symbol():
    """the shared libraries load on a GPU-less host and export exactly what include/*.h declares: libdedflow_b200.so the hot path
    (dedflow_b200.h, dedflow_compat.h), libdedflow_h5flat.so the file interface (dedflow_h5flat.h)"""
    import ctypes
    import __graft_entry__
    from dedflow_b200 import _build, lib
    _build.build()
    L = lib.load()
    assert L.dfb_version() >= 100
    declared = _declared(ROOT / "include" / "dedflow_b200.h") | _declared(ROOT / "include" / "dedflow_compat.h")
    missing = [d for d in sorted(declared) if not hasattr(L, d)]
    assert not missing, f"declared but not exported: {missing}"
    assert set(lib.exported_symbols()) <= declared
    __graft_entry__.build_h5flat()
    H = ctypes.CDLL(str(ROOT / "dedflow_b200" / "libdedflow_h5flat.so"))
    h5 = _declared(ROOT / "include" / "dedflow_h5flat.h")
    assert len(h5) == 20 and not [d for d in sorted(h5) if not hasattr(H, d)]
    assert {p.name for p in (ROOT / "include").glob("*.h")} == {"dedflow_b200.h", "dedflow_compat.h", "dedflow_h5flat.h"}


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dedflow_b200 import api, lib
    with pytest.raises(lib.DfbError):
        api.FlowSystem(boxmesh.make_box(2))


def test_option_table_accepts_the_documented_switches_and_rejects_others():
    """dfb_set_option is host-only: every key the header documents is accepted, an unknown one is an argument error"""
    from dedflow_b200 import _build, lib
    _build.build()
    lib.load()
    header = (ROOT / "include" / "dedflow_b200.h").read_text()
    block = header[header.index("Variant switches"):header.index("int dfb_set_option")]
    keys = sorted(set(re.findall(r"DFB_[A-Z0-9_]+", block)))
    assert {"DFB_GMRES_CHECK", "DFB_HALO_DEFER", "DFB_GIVENS_DEFER", "DFB_F_VARIANT", "DFB_J_PAIR_NT"} <= set(keys)
    defaults = {"DFB_J_VARIANT": "pairs", "DFB_J_PAIR_ORDER": "morton", "DFB_F_VARIANT": "patch", "DFB_ASSEMBLE_MODE": "gather",
                "DFB_PC": "jacobi", "DFB_J_PAIR_ROWS": "8", "DFB_J_PAIR_NT": "96", "DFB_F_PATCH_CTAS": "2", "DFB_SPMV_G": "8",
                "DFB_GMRES_CHECK": "20", "DFB_PC_AGG": "4", "DFB_PC_DEGREE": "10", "DFB_HALO_DEFER": "1", "DFB_GIVENS_DEFER": "1",
                "DFB_GRAPH": "1"}
    for k in keys:
        lib.set_option(k, defaults.get(k, "0"))            # (sets every switch to its default)
    with pytest.raises(lib.DfbError):
        lib.set_option("DFB_NO_SUCH_SWITCH", "1")


def test_bench_ownership_policy():
    """bench.py --owner auto: slabs when the box's planes divide evenly among the ranks, coordinate bisection otherwise"""
    from dedflow_b200 import dist as ddist
    for world, m, want in ((2, 69, "slab"), (4, 87, "slab"), (8, 110, "rcb"), (3, 8, "slab"), (3, 9, "rcb")):
        mesh = boxmesh.make_box(m) if m < 20 else type("M", (), {"m": m})()
        name, fn = ddist.pick_owner("auto", mesh, world)
        assert name == want and fn is ddist.OWNERS[want]
    mesh = boxmesh.make_box(9)
    for name in ("slab", "rcb"):
        npart = ddist.pick_owner(name, mesh, 3)[1](mesh, 3)
        counts = np.bincount(npart, minlength=3)
        assert counts.sum() == mesh.num_node and counts.min() > 0
        if name == "rcb":
            assert counts.max() - counts.min() <= 1         # balanced to a node


def test_every_parsed_switch_is_also_read_from_the_environment():
    """setup.cu: the keys apply_option() parses and the keys options() reads from the environment at load are the same set
    (a switch missing from the second list works through dfb_set_option but silently ignores the environment)"""
    src = (ROOT / "dedflow_b200" / "csrc" / "setup.cu").read_text()
    parsed = set(re.findall(r'is\("(DFB_[A-Z0-9_]+)"\)', src))
    env_list = src[src.index("static const char* keys[]"):]
    env_list = set(re.findall(r'"(DFB_[A-Z0-9_]+)"', env_list[:env_list.index("};")]))
    assert parsed == env_list, (sorted(parsed - env_list), sorted(env_list - parsed))


def test_pc2_restatement_is_a_better_preconditioner_than_block_jacobi(oracle):
    """oracle/pc2_oracle.py (the checker of the opt-in preconditioner, tests/test_gpu_pc2.py) on the first Newton system of a
    time step: its right-preconditioned GMRES meets the tolerance in fewer iterations than the reference's block-Jacobi, the
    returned vector has the residual the history claims, and one application is linear."""
    import scipy.sparse as sp
    from oracle import pc2_oracle
    mesh = boxmesh.make_box(10)
    N = mesh.num_node
    ctx = oracle.driver_setup(mesh)
    wgold, dwgold, dwg = boxmesh.state_initial(mesh)
    fac = (oracle.K_GAMMA - 1.0) / oracle.K_GAMMA
    dwg[:3 * N] *= fac
    dwg[4 * N:] *= fac
    wga, dwga = oracle.alpha_states(N, wgold, dwgold, dwg)
    F = oracle.assemble_system(ctx, wga, dwga, want_F=True)
    blocks = oracle.assemble_system(ctx, wga, dwga, want_J=True)
    _, it_jacobi, hist_jacobi = oracle.gmres(ctx["pattern"], blocks, F)
    R = pc2_oracle.Pc2Oracle(oracle, mesh, ctx["pattern"], blocks, agg_cells=4, cheb_degree=10)
    x, it, hist = R.gmres(F)
    assert it % 20 == 0 and it <= it_jacobi and hist[-1] <= 1e-4 * hist[0]
    assert hist[20] < hist_jacobi[20]                       # and it is ahead at the first test already
    A = sp.bmat([[R.A00, R.A01], [R.A10, R.A11]]).tocsr()
    res = np.linalg.norm(F[:4 * N] - A @ x[:4 * N])
    assert abs(res - hist[-1]) <= 1e-8 * hist[0]
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal(6 * N), rng.standard_normal(6 * N)
    lin = R.apply(2.0 * a - 3.0 * b)[:4 * N] - (2.0 * R.apply(a) - 3.0 * R.apply(b))[:4 * N]
    assert np.abs(lin).max() <= 1e-9 * np.abs(R.apply(a)[:4 * N]).max()
