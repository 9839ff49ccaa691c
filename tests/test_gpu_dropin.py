"""Drop-in test (pytest -m gpu): the reference's own host code that stays in a DEDFlow build (Mesh.c, MeshData.c, common.c,
alloc.c -- compiled unmodified by oracle/ref/Makefile into oracle/_ref/libdedflow_hybrid.so) linked against the PRODUCT
library, driven by the very same ctypes driver (oracle/ref/reflib.py) that drives the reference's own CUDA build.  Every
struct the driver pokes into (Mesh3D, CSRAttr, Matrix, MatrixFS, MatrixCSR, Dirichlet, Krylov) is the reference's layout;
every hot-path entry point (CSRAttrCreate ... KrylovSolve, include/dedflow_compat.h) resolves to libdedflow_b200.so.
Results are checked against the CPU oracle with the same bars as tests/test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

from conftest import shuffled_mesh
from dedflow_b200 import boxmesh
from oracle import pyoracle
from test_gpu_parity import TOL_ASM, TOL_SOLVE, oracle_system, rel

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def reflib():
    from oracle.ref import reflib as r
    if not r.hybrid_available():
        pytest.fail("oracle/_ref/libdedflow_hybrid.so is missing: run __graft_entry__.build() where /root/reference exists")
    return r


def hybrid(reflib, mesh):
    # patch_d1=False: the product writes the final row_ptr entry of the blocked patterns itself (defect D1 fixed)
    return reflib.RefProblem(mesh, patch_d1=False, so_path=reflib.HYBRID_SO)


def test_hybrid_resolves_hot_path_into_product_library(reflib):
    """the hybrid library itself only holds Mesh*/common/alloc: the hot-path symbols come from libdedflow_b200.so"""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", str(reflib.HYBRID_SO)], capture_output=True, text=True).stdout
    defined = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    assert "Mesh3DCreate" in defined and "Mesh3DGenerateColorBatch" in defined
    for sym in ("CSRAttrCreate", "ColorMeshTet", "AssembleSystemTet", "KrylovSolve", "MatrixMatVec", "DirichletApplyMat"):
        assert sym not in defined
    prod = subprocess.run(["nm", "-D", "--defined-only", str(reflib.HYBRID_SO.parents[2] / "dedflow_b200" / "libdedflow_b200.so")],
                          capture_output=True, text=True).stdout
    for sym in ("CSRAttrCreate", "ColorMeshTet", "AssembleSystemTet", "KrylovSolve", "MatrixMatVec", "DirichletApplyMat"):
        assert f" T {sym}" in prod


@pytest.mark.parametrize("m,shuffle", [(2, True), (6, True), (20, False)])
def test_dropin_setup_bit_exact(reflib, oracle, m, shuffle):
    mesh = shuffled_mesh(m) if shuffle else boxmesh.make_box(m)
    R = hybrid(reflib, mesh)
    rp, ci = oracle.nodal_pattern(mesh.num_node, mesh.ien)
    pats = R.patterns()
    assert np.array_equal(pats["1x1"][0], rp) and np.array_equal(pats["1x1"][1], ci)
    for name, (br, bc) in {"3x3": (3, 3), "3x1": (3, 1), "1x3": (1, 3)}.items():
        orp, oci = oracle.expand_block(rp, ci, br, bc, fix_last=True)
        assert np.array_equal(pats[name][0], orp) and np.array_equal(pats[name][1], oci)
    color, off, ind, nc = R.color_batches()          # reference Mesh3DGenerateColorBatch -> our ColorMeshTet / GetMaxColor /
    w = oracle.weights(pyoracle.curand_host_u32(mesh.num_tet))   # CountValueColorLegacy / FindValueColor
    ocolor, rounds, ties = oracle.color_jpl(mesh.num_node, mesh.ien, w)
    assert ties == 0 and nc == rounds and np.array_equal(color, ocolor)
    ooff, oind = oracle.color_batches(ocolor)
    assert np.array_equal(off, ooff) and np.array_equal(ind, oind)


@pytest.mark.parametrize("m,shuffle", [(3, True), (8, True), (20, False)])
def test_dropin_assemble_solve(reflib, oracle, m, shuffle):
    mesh = shuffled_mesh(m) if shuffle else boxmesh.make_box(m)
    N = mesh.num_node
    wg, dwg = boxmesh.state_random(N)
    ref = oracle_system(oracle, mesh, wg, dwg)
    R = hybrid(reflib, mesh)
    R.color_batches()
    L = C.CDLL(str(reflib.HYBRID_SO.parents[2] / "dedflow_b200" / "libdedflow_b200.so"))
    d_wg, d_dwg = torch.from_numpy(wg).cuda(), torch.from_numpy(dwg).cuda()
    F = torch.full((6 * N,), 7.0, dtype=torch.float64, device="cuda")
    R.assemble(d_wg, d_dwg, F_t=F)                   # main.c:31-75 through the drop-in entry points
    R.assemble(d_wg, d_dwg, J=True)
    Fh = F.cpu().numpy()
    assert rel(Fh[:3 * N], ref["F"][:3 * N]) <= TOL_ASM
    assert np.abs(Fh[3 * N:4 * N] - ref["F"][3 * N:4 * N]).max() <= TOL_ASM * np.abs(ref["F"]).max()
    assert np.all(Fh[4 * N:] == 0)
    for got, want, name in zip(R.block_vals(), ref["blocks"], ("A00", "A01", "A10", "A11")):
        assert rel(got, want) <= TOL_ASM, name
    # MatrixMatVec
    rng = np.random.default_rng(5)
    x = rng.standard_normal(6 * N)
    y0 = rng.standard_normal(6 * N)
    yo = y0.copy()
    oracle.fs_amvpby(ref["pattern"], ref["blocks"], 1.0, x, 0.0, yo)
    dy = torch.from_numpy(y0.copy()).cuda()
    R.matvec(torch.from_numpy(x).cuda(), dy)
    got = dy.cpu().numpy()
    assert rel(got[:4 * N], yo[:4 * N]) <= 1e-12 and np.array_equal(got[4 * N:], y0[4 * N:])
    # KrylovSolve: iterations, printed residual log (reference format), solution
    xo, ito, histo = oracle.gmres(ref["pattern"], ref["blocks"], ref["F"])
    dx = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    printed = R.solve(dx, F)
    assert [p[0] for p in printed] == list(range(0, ito + 1, 20))
    for it, val in printed:
        assert abs(val - histo[it]) <= 6e-5 * histo[it] + 1e-300          # 5 printed digits
    L.dfb_compat_last_history.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    hist = np.zeros(121)
    assert L.dfb_compat_last_history(R.ksp, hist.ctypes.data, 121) == ito
    assert np.abs(hist[:ito + 1] - histo).max() <= TOL_SOLVE * histo[0]
    assert rel(dx.cpu().numpy()[:4 * N], xo[:4 * N]) <= TOL_SOLVE


def test_dropin_generic_matrix_ops(reflib, oracle):
    """MatrixZeroRow / MatrixGetDiag / MatrixAMVPBY on a single CSR block, PCJacobi / PCNone through the exported entry
    points, and MatrixAddElemValueBlockedBatched with the reference's (lda=6, stride=36) element layout."""
    mesh = shuffled_mesh(4)
    N, E = mesh.num_node, mesh.num_tet
    R = hybrid(reflib, mesh)
    L = R.L
    rp, ci = oracle.nodal_pattern(N, mesh.ien)
    Z = ci.size
    rng = np.random.default_rng(9)
    # --- element scatter: every (e,a,b) 6x6 block -> the four sub-blocks, colour by colour -----------------------
    color, off, ind, nc = R.color_batches()
    elemJ = rng.standard_normal((E, 16, 36))
    d_elemJ = torch.from_numpy(elemJ).cuda()
    L.MatrixZero(R.J)
    L.MatrixAddElemValueBlockedBatched.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                                   C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    d_ien = R.mesh.contents.device.contents.ien
    d_batch = R.mesh.contents.batch_ind
    for c in range(nc):
        n = int(off[c + 1] - off[c])
        sub = torch.from_numpy(elemJ[ind[off[c]:off[c + 1]]].copy()).cuda()   # batch-local layout, like the reference's elem_J
        L.MatrixAddElemValueBlockedBatched(C.cast(R.J, C.c_void_p), 4, n, C.c_void_p(d_batch + 4 * int(off[c])), C.c_void_p(d_ien),
                                           6, 6, C.c_void_p(sub.data_ptr()), 6, 36, None)
        torch.cuda.synchronize()
    want = [np.zeros(9 * Z), np.zeros(3 * Z), np.zeros(3 * Z), np.zeros(Z)]
    for e in range(E):
        for a in range(4):
            ra = mesh.ien[e, a]
            s, ln = rp[ra], rp[ra + 1] - rp[ra]
            for b in range(4):
                k = int(np.searchsorted(ci[s:s + ln], mesh.ien[e, b]))
                blk = elemJ[e, a * 4 + b].reshape(6, 6)
                for ii in range(3):
                    for jj in range(3):
                        want[0][s * 9 + ii * 3 * ln + k * 3 + jj] += blk[ii, jj]
                    want[1][s * 3 + ii * ln + k] += blk[ii, 3]
                    want[2][s * 3 + k * 3 + ii] += blk[3, ii]
                want[3][s + k] += blk[3, 3]
    for got, w in zip(R.block_vals(), want):
        assert rel(got, w) <= 1e-13
    want = R.block_vals()      # from here on compare bit-for-bit against what is stored (summation order differs from numpy's)
    # --- single CSR block: amvpby, get_diag, zero_row ---------------------------------------------------------------
    A00 = R.fs.mat[0]
    L.MatrixAMVPBY.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p]
    x = rng.standard_normal(3 * N)
    y0 = rng.standard_normal(3 * N)
    dy = torch.from_numpy(y0.copy()).cuda()
    dxv = torch.from_numpy(x).cuda()
    L.MatrixAMVPBY(C.cast(A00, C.c_void_p), 0.5, C.c_void_p(dxv.data_ptr()), -2.0, C.c_void_p(dy.data_ptr()))
    orp, oci = oracle.expand_block(rp, ci, 3, 3, fix_last=True)
    import scipy.sparse as sp
    M = sp.csr_matrix((want[0], oci, orp), shape=(3 * N, 3 * N))
    assert rel(dy.cpu().numpy(), 0.5 * (M @ x) - 2.0 * y0) <= 1e-12
    L.MatrixGetDiag.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    d1 = torch.zeros(3 * N, dtype=torch.float64, device="cuda")
    L.MatrixGetDiag(C.cast(A00, C.c_void_p), C.c_void_p(d1.data_ptr()), 1)
    assert np.array_equal(d1.cpu().numpy(), M.diagonal())
    d3 = torch.zeros(9 * N, dtype=torch.float64, device="cuda")
    L.MatrixGetDiag(C.cast(A00, C.c_void_p), C.c_void_p(d3.data_ptr()), 3)
    Md = M.toarray()
    blocks = np.stack([Md[3 * i:3 * i + 3, 3 * i:3 * i + 3] for i in range(N)])
    assert np.array_equal(d3.cpu().numpy().reshape(N, 3, 3), blocks)            # row-major blocks (matrix_impl.cu:663-672)
    # PCJacobi(bs=3): applies (B^-1)^T (defect D3); PCJacobi(bs=1); PCNone
    L.PCCreateJacobi.restype = C.c_void_p
    L.PCCreateJacobi.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    L.PCSetup.argtypes = [C.c_void_p]
    L.PCApply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.PCDestroy.argtypes = [C.c_void_p]
    pc = L.PCCreateJacobi(C.cast(A00, C.c_void_p), 3, None)
    L.PCSetup(pc)
    out = torch.zeros(3 * N, dtype=torch.float64, device="cuda")
    L.PCApply(pc, C.c_void_p(dxv.data_ptr()), C.c_void_p(out.data_ptr()))
    wantp = np.einsum("nrc,nr->nc", np.linalg.inv(blocks), x.reshape(N, 3)).ravel()   # (B^-1)^T x
    assert rel(out.cpu().numpy(), wantp) <= 1e-10
    L.PCDestroy(pc)
    # zero_row through the FS matrix: velocity rows of three nodes become unit rows (A00) / zero rows (A01), columns kept
    rows = np.array([0 * 3 + 1, 5 * 3 + 0, 7 * 3 + 2], np.int32)
    d_rows = torch.from_numpy(rows).cuda()
    L.MatrixZeroRow.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_double]
    L.MatrixZeroRow(C.cast(R.J, C.c_void_p), 3, C.c_void_p(d_rows.data_ptr()), 0, 1.0)
    got = R.block_vals()
    M2 = sp.csr_matrix((got[0], oci, orp), shape=(3 * N, 3 * N)).toarray()
    want2 = Md.copy()
    want2[rows] = 0.0
    want2[rows, rows] = 1.0
    assert np.array_equal(M2, want2)
    o31 = oracle.expand_block(rp, ci, 3, 1, fix_last=True)
    B01 = sp.csr_matrix((got[1], o31[1], o31[0]), shape=(3 * N, N)).toarray()
    W01 = sp.csr_matrix((want[1], o31[1], o31[0]), shape=(3 * N, N)).toarray()
    W01[rows] = 0.0
    assert np.array_equal(B01, W01)
    assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])    # pressure rows untouched (defect D8)


def test_dropin_vector_helpers(reflib):
    """VecAXPY / VecPointwiseMult / VecPointwiseDiv / VecPointwiseInv (reference vec.h:7-10, vec.cu:14-70) through the exported
    drop-in symbols: same argument order (VecAXPY(a, x, y, n): y = a*x + y; Mult/Div(x, y, z, n): z = x op y; Inv in place),
    bit-exact against numpy (one rounding per entry), in-place aliasing, n not a multiple of the block size, n = 0."""
    L = C.CDLL(str(reflib.HYBRID_SO.parents[2] / "dedflow_b200" / "libdedflow_b200.so"))
    P = lambda t: C.c_void_p(t.data_ptr())
    L.VecAXPY.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_int32]
    L.VecPointwiseMult.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    L.VecPointwiseDiv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    L.VecPointwiseInv.argtypes = [C.c_void_p, C.c_int32]
    rng = np.random.default_rng(21)
    for n in (1, 255, 256, 257, 100003):
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        y[np.abs(y) < 1e-3] = 1.0
        dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y.copy()).cuda()
        dz = torch.full((n + 8,), -7.25, dtype=torch.float64, device="cuda")          # guard behind the end
        L.VecPointwiseMult(P(dx), P(dy), P(dz), n)
        assert np.array_equal(dz.cpu().numpy()[:n], x * y) and bool((dz[n:] == -7.25).all())
        L.VecPointwiseDiv(P(dx), P(dy), P(dz), n)
        assert np.array_equal(dz.cpu().numpy()[:n], x / y) and bool((dz[n:] == -7.25).all())
        L.VecAXPY(-0.375, P(dx), P(dy), n)                                              # y = a*x + y
        got = dy.cpu().numpy()
        want_fma = np.array([np.float64(np.longdouble(-0.375) * np.longdouble(a) + np.longdouble(b)) for a, b in zip(x[:64], y[:64])])
        assert np.abs(got - (-0.375 * x + y)).max() <= 2.3e-16 * np.abs(got).max()    # fused or unfused multiply-add
        assert np.array_equal(got[:64], want_fma) or np.array_equal(got[:64], (-0.375 * x + y)[:64])
        inv = torch.from_numpy(y.copy()).cuda()
        L.VecPointwiseInv(P(inv), n)
        assert np.array_equal(inv.cpu().numpy(), 1.0 / y)
        L.VecPointwiseMult(P(dx), P(dx), P(dx), n)                                      # aliasing: x = x*x in place
        assert np.array_equal(dx.cpu().numpy(), x * x)
    before = dz.clone()
    L.VecPointwiseMult(P(dx), P(dy), P(dz), 0)
    L.VecAXPY(1.0, P(dx), P(dz), 0)
    L.VecPointwiseInv(P(dz), 0)
    torch.cuda.synchronize()
    assert torch.equal(before, dz)
