"""ctypes front end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of oracle.c.  Imported by tests/, by
``__graft_entry__.smoke()`` and by ``bench.py``'s cpu_baseline / ``--impl reference`` legs; never by
the product package ``dedflow_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "liboracle.so"

I32P = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
F64P = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
U32P = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> Path:
    """Compile oracle.c -> oracle/_build/liboracle.so (gcc, OpenMP, no FMA contraction)."""
    src = HERE / "oracle.c"
    if LIB_PATH.exists() and not force and LIB_PATH.stat().st_mtime >= src.stat().st_mtime:
        return LIB_PATH
    LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["gcc", "-O2", "-march=native", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden",
           "-Wall", "-o", str(LIB_PATH), str(src), "-lm"]
    if os.environ.get("ORACLE_PORTABLE", "1") == "1":
        cmd.remove("-march=native")      # the .so travels to the GPU box: keep it portable
    subprocess.run(cmd, check=True)
    return LIB_PATH


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self):
        build()
        L = C.CDLL(str(LIB_PATH))
        self.L = L
        L.orc_nodal_pattern.restype = C.c_int64
        L.orc_nodal_pattern.argtypes = [C.c_int32, C.c_int32, I32P, I32P, I32P]
        L.orc_expand_block.restype = None
        L.orc_expand_block.argtypes = [C.c_int32, I32P, I32P, C.c_int32, C.c_int32, I32P, I32P, C.c_int]
        L.orc_weights_from_u32.restype = None
        L.orc_weights_from_u32.argtypes = [C.c_int64, U32P, I32P]
        L.orc_v2e.restype = None
        L.orc_v2e.argtypes = [C.c_int32, C.c_int32, I32P, I32P, I32P]
        L.orc_color_jpl.restype = C.c_int32
        L.orc_color_jpl.argtypes = [C.c_int32, C.c_int32, I32P, I32P, C.c_int32, I32P, C.POINTER(C.c_int64)]
        L.orc_color_batches.restype = None
        L.orc_color_batches.argtypes = [C.c_int32, I32P, C.c_int32, I32P, I32P]
        L.orc_tet_elements.restype = None
        L.orc_tet_elements.argtypes = [C.c_int32, C.c_void_p, C.c_int32, I32P, F64P, F64P, F64P] + [C.c_void_p] * 4
        L.orc_assemble_tet.restype = None
        L.orc_assemble_tet.argtypes = [C.c_int32, I32P, F64P, C.c_int32, I32P, I32P, F64P, F64P, C.c_void_p,
                                       C.c_void_p, C.c_void_p] + [C.c_void_p] * 4
        L.orc_face_elements.restype = None
        L.orc_face_elements.argtypes = [C.c_int32, I32P, I32P, C.c_int32, I32P, F64P, F64P, F64P, C.c_void_p, C.c_void_p]
        L.orc_assemble_face.restype = None
        L.orc_assemble_face.argtypes = [C.c_int32, I32P, I32P, C.c_int32, I32P, F64P, I32P, C.c_int32, F64P, F64P,
                                        C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_void_p] * 4
        L.orc_dirichlet_vec.restype = None
        L.orc_dirichlet_vec.argtypes = [C.c_int32, I32P, C.c_int32, I32P, F64P]
        L.orc_dirichlet_mat.restype = None
        L.orc_dirichlet_mat.argtypes = [C.c_int32, I32P, C.c_int32, I32P, C.c_int32, I32P, I32P, F64P, F64P]
        L.orc_fs_amvpby.restype = None
        L.orc_fs_amvpby.argtypes = [C.c_int32, I32P, I32P, F64P, F64P, F64P, F64P, C.c_double, F64P, C.c_double, F64P]
        L.orc_pc_setup.restype = None
        L.orc_pc_setup.argtypes = [C.c_int32, I32P, I32P, F64P, F64P, F64P, F64P]
        L.orc_pc_apply.restype = None
        L.orc_pc_apply.argtypes = [C.c_int32, F64P, F64P, F64P, F64P]
        L.orc_gmres.restype = C.c_int32
        L.orc_gmres.argtypes = [C.c_int32, I32P, I32P, F64P, F64P, F64P, F64P, C.c_int32, C.c_double, C.c_double, F64P,
                                F64P, F64P]
        L.orc_num_threads.restype = C.c_int

    # ---- a1/a2 pattern -----------------------------------------------------------------
    def nodal_pattern(self, num_node, ien):
        ien = np.ascontiguousarray(ien, dtype=np.int32).reshape(-1)
        row_ptr = np.zeros(num_node + 1, np.int32)
        cap = np.zeros(num_node * 64, np.int32)
        nnz = self.L.orc_nodal_pattern(num_node, ien.size // 4, ien, row_ptr, cap)
        if nnz < 0:
            raise OverflowError("row longer than PREALLOC_SIZE=64 (reference csr.c:64 asserts)")
        return row_ptr, cap[:nnz].copy()

    def expand_block(self, row_ptr, col_ind, br, bc, fix_last=True):
        n = row_ptr.size - 1
        nnz = int(row_ptr[-1])
        nrp = np.zeros(n * br + 1, np.int32)
        nci = np.zeros(nnz * br * bc, np.int32)
        self.L.orc_expand_block(n, row_ptr, col_ind, br, bc, nrp, nci, int(fix_last))
        return nrp, nci

    # ---- a3/a4 coloring ----------------------------------------------------------------
    def weights(self, raw_u32):
        raw = np.ascontiguousarray(raw_u32, dtype=np.uint32)
        w = np.zeros(raw.size, np.int32)
        self.L.orc_weights_from_u32(raw.size, raw, w)
        return w

    def v2e(self, num_node, ien):
        ien = np.ascontiguousarray(ien, dtype=np.int32).reshape(-1)
        rp = np.zeros(num_node + 1, np.int32)
        ci = np.zeros(ien.size, np.int32)
        self.L.orc_v2e(num_node, ien.size // 4, ien, rp, ci)
        return rp, ci

    def color_jpl(self, num_node, ien, weight, max_color=256):
        ien = np.ascontiguousarray(ien, dtype=np.int32).reshape(-1)
        color = np.zeros(ien.size // 4, np.int32)
        ties = C.c_int64(0)
        rounds = self.L.orc_color_jpl(ien.size // 4, num_node, ien, np.ascontiguousarray(weight, np.int32), max_color,
                                      color, C.byref(ties))
        return color, int(rounds), int(ties.value)

    def color_batches(self, color):
        color = np.ascontiguousarray(color, np.int32)
        nc = int(color.max()) + 1
        off = np.zeros(nc + 1, np.int32)
        ind = np.zeros(color.size, np.int32)
        self.L.orc_color_batches(color.size, color, nc, off, ind)
        return off, ind

    # ---- a6-a8 assembly ----------------------------------------------------------------
    def tet_elements(self, num_node, ien, xg, wg, dwg, elem_ids=None, want_F=True, want_J=True, want_geom=False):
        ien = np.ascontiguousarray(ien, dtype=np.int32).reshape(-1)
        xg = np.ascontiguousarray(xg, dtype=np.float64).reshape(-1)
        if elem_ids is not None:
            elem_ids = np.ascontiguousarray(elem_ids, np.int32)
            n = elem_ids.size
        else:
            n = ien.size // 4
        eF = np.zeros((n, 4, 6)) if want_F else None
        eJ = np.zeros((n, 4, 4, 6, 6)) if want_J else None
        met = np.zeros((n, 10)) if want_geom else None
        shg = np.zeros((n, 4, 3)) if want_geom else None
        self.L.orc_tet_elements(n, _opt(elem_ids), num_node, ien, xg, wg, dwg, _opt(eF), _opt(eJ), _opt(met), _opt(shg))
        return eF, eJ, met, shg

    def assemble_tet(self, num_node, ien, xg, batch_offset, batch_ind, wg, dwg, F=None, pattern=None, blocks=None):
        ien = np.ascontiguousarray(ien, dtype=np.int32).reshape(-1)
        xg = np.ascontiguousarray(xg, dtype=np.float64).reshape(-1)
        rp, ci = pattern if pattern is not None else (None, None)
        b = blocks if blocks is not None else (None,) * 4
        self.L.orc_assemble_tet(num_node, ien, xg, batch_offset.size - 1, batch_offset, batch_ind, wg, dwg, _opt(F),
                                _opt(rp), _opt(ci), *[_opt(a) for a in b])

    def face_elements(self, f2e, forn, num_node, ien, xg, wg, dwg, want_F=True, want_J=True):
        ien = np.ascontiguousarray(ien, dtype=np.int32).reshape(-1)
        xg = np.ascontiguousarray(xg, dtype=np.float64).reshape(-1)
        n = f2e.size
        eF = np.zeros((n, 4, 6)) if want_F else None
        eJ = np.zeros((n, 4, 4, 6, 6)) if want_J else None
        self.L.orc_face_elements(n, np.ascontiguousarray(f2e, np.int32), np.ascontiguousarray(forn, np.int32), num_node,
                                 ien, xg, wg, dwg, _opt(eF), _opt(eJ))
        return eF, eJ

    def assemble_face(self, f2e, forn, num_node, ien, xg, color, num_color, wg, dwg, F=None, pattern=None, blocks=None):
        ien = np.ascontiguousarray(ien, dtype=np.int32).reshape(-1)
        xg = np.ascontiguousarray(xg, dtype=np.float64).reshape(-1)
        rp, ci = pattern if pattern is not None else (None, None)
        b = blocks if blocks is not None else (None,) * 4
        self.L.orc_assemble_face(f2e.size, np.ascontiguousarray(f2e, np.int32), np.ascontiguousarray(forn, np.int32),
                                 num_node, ien, xg, np.ascontiguousarray(color, np.int32), num_color, wg, dwg, _opt(F),
                                 _opt(rp), _opt(ci), *[_opt(a) for a in b])

    # ---- a9 Dirichlet ------------------------------------------------------------------
    def dirichlet_vec(self, bnode, bctype, b, shape=3):
        self.L.orc_dirichlet_vec(bnode.size, np.ascontiguousarray(bnode, np.int32), shape,
                                 np.ascontiguousarray(bctype, np.int32), b)

    def dirichlet_mat(self, bnode, bctype, num_node, pattern, A00, A01, shape=3):
        rp, ci = pattern
        self.L.orc_dirichlet_mat(bnode.size, np.ascontiguousarray(bnode, np.int32), shape,
                                 np.ascontiguousarray(bctype, np.int32), num_node, rp, ci, A00, A01)

    # ---- a10-a12 solve -----------------------------------------------------------------
    def fs_amvpby(self, pattern, blocks, alpha, x, beta, y):
        rp, ci = pattern
        self.L.orc_fs_amvpby(rp.size - 1, rp, ci, *blocks, alpha, x, beta, y)

    def pc_setup(self, pattern, blocks):
        rp, ci = pattern
        n = rp.size - 1
        d00 = np.zeros(9 * n)
        d11 = np.zeros(n)
        self.L.orc_pc_setup(n, rp, ci, blocks[0], blocks[3], d00, d11)
        return d00, d11

    def pc_apply(self, d00, d11, x):
        y = np.zeros_like(x)
        self.L.orc_pc_apply(d11.size, d00, d11, x, y)
        return y

    def gmres(self, pattern, blocks, b, x0=None, maxit=120, atol=1e-12, rtol=1e-4):
        rp, ci = pattern
        n = rp.size - 1
        x = np.zeros(6 * n) if x0 is None else np.array(x0, dtype=np.float64)
        hist = np.zeros(maxit + 1)
        it = self.L.orc_gmres(n, rp, ci, *blocks, maxit, atol, rtol, x, np.ascontiguousarray(b, np.float64), hist)
        return x, int(it), hist[:it + 1]

    # ------------------------------------------------------------------------------------------------------------
    # the driver around the path: numpy restatement of reference src/main.c:31-75 (AssembleSystem), :77-283
    # (SolveFlowSystem) and :537-565 (one pass of the time loop).  SURVEY.md section 8(f), rank 1.
    # ------------------------------------------------------------------------------------------------------------
    K_RHOC, K_DT = 0.5, 5e-2
    K_ALPHAM = (3.0 - K_RHOC) / (1.0 + K_RHOC)
    K_ALPHAF = 1.0 / (1.0 + K_RHOC)
    K_GAMMA = 0.5 + K_ALPHAM - K_ALPHAF
    BCS = {0: (1, 1, 1), 2: (0, 1, 0), 3: (0, 0, 1), 4: (0, 0, 0)}          # main.c:454-476

    def driver_setup(self, mesh):
        """what main() builds once (main.c:377-411): pattern, colors, batches"""
        N = mesh.num_node
        rp, ci = self.nodal_pattern(N, mesh.ien)
        w = self.weights(curand_host_u32(mesh.num_tet))
        color, nc, ties = self.color_jpl(N, mesh.ien, w)
        off, ind = self.color_batches(color)
        return dict(mesh=mesh, pattern=(rp, ci), color=color, nc=nc, off=off, ind=ind)

    def assemble_system(self, ctx, wgalpha, dwgalpha, want_F=False, want_J=False):
        """AssembleSystem (main.c:31-75): zero, tets, faces of boundary 4, zero the phi/T residual, Dirichlet"""
        mesh = ctx["mesh"]
        N = mesh.num_node
        rp, ci = ctx["pattern"]
        Z = ci.size
        F = np.zeros(6 * N) if want_F else None
        blocks = [np.zeros(9 * Z), np.zeros(3 * Z), np.zeros(3 * Z), np.zeros(Z)] if want_J else None
        kw = dict(F=F) if want_F else dict(pattern=(rp, ci), blocks=blocks)
        self.assemble_tet(N, mesh.ien, mesh.xg, ctx["off"], ctx["ind"], wgalpha, dwgalpha, **kw)
        if mesh.num_bound > 4:
            f2e, forn = mesh.bound_faces(4)
            self.assemble_face(f2e, forn, N, mesh.ien, mesh.xg, ctx["color"], ctx["nc"], wgalpha, dwgalpha, **kw)
        if want_F:
            F[4 * N:] = 0.0
        for b, t in self.BCS.items():
            bt = np.array(t, np.int32)
            if want_F:
                self.dirichlet_vec(mesh.bound_nodes(b), bt, F)
            if want_J:
                self.dirichlet_mat(mesh.bound_nodes(b), bt, N, (rp, ci), blocks[0], blocks[1])
        return F if want_F else blocks

    def alpha_states(self, N, wgold, dwgold, dwg):
        """main.c:107-118: the axpy sequences exactly as written (two accumulating steps each)"""
        f1 = (1.0 - self.K_ALPHAM, self.K_ALPHAM)
        f2 = (self.K_DT * self.K_ALPHAF * (1.0 - self.K_GAMMA), self.K_DT * self.K_ALPHAF * self.K_GAMMA)
        dwgalpha = np.zeros(6 * N)
        dwgalpha += f1[0] * dwgold
        dwgalpha += f1[1] * dwg
        dwgalpha[3 * N:4 * N] = dwg[3 * N:4 * N]
        wgalpha = wgold.copy()
        wgalpha += f2[0] * dwgold
        wgalpha += f2[1] * dwg
        wgalpha[3 * N:4 * N] = 0.0
        return wgalpha, dwgalpha

    @staticmethod
    def block_norms(N, F):
        return np.array([np.linalg.norm(F[:3 * N]), np.linalg.norm(F[3 * N:4 * N]), np.linalg.norm(F[4 * N:5 * N]),
                         np.linalg.norm(F[5 * N:])])

    def solve_flow_system(self, ctx, wgold, dwgold, dwg, maxit=4, tol=0.5e-3):
        """SolveFlowSystem (main.c:77-283); dwg is updated in place; returns [(rnorm[4], gmres its)]"""
        N = ctx["mesh"].num_node
        wga, dwga = self.alpha_states(N, wgold, dwgold, dwg)
        F = self.assemble_system(ctx, wga, dwga, want_F=True)
        r0 = self.block_norms(N, F)
        hist = [(r0.copy(), 0)]
        r0 = r0 + 1e-16
        it, converged = 0, False
        while not converged and it < maxit:
            blocks = self.assemble_system(ctx, wga, dwga, want_J=True)
            dx, its, _ = self.gmres(ctx["pattern"], blocks, F)
            dwg -= dx
            wga, dwga = self.alpha_states(N, wgold, dwgold, dwg)
            F = self.assemble_system(ctx, wga, dwga, want_F=True)
            r = self.block_norms(N, F)
            hist.append((r.copy(), its))
            converged = bool(np.all(r < tol * r0))
            it += 1
        return hist

    def time_step(self, ctx, wgold, dwgold, dwg, **kw):
        """one pass of the time loop, main.c:537-565; the three vectors are updated in place"""
        N = ctx["mesh"].num_node
        fac = (self.K_GAMMA - 1.0) / self.K_GAMMA
        dwg[:3 * N] *= fac
        dwg[4 * N:] *= fac
        hist = self.solve_flow_system(ctx, wgold, dwgold, dwg, **kw)
        c0, c1 = self.K_DT * (1.0 - self.K_GAMMA), self.K_DT * self.K_GAMMA
        for sl in (slice(0, 3 * N), slice(4 * N, 6 * N)):
            wgold[sl] += c0 * dwgold[sl]
            wgold[sl] += c1 * dwg[sl]
        dwgold[:] = dwg
        return hist

    def num_threads(self):
        return int(self.L.orc_num_threads())

    def reference_step(self, mesh, wg, dwg, solve=True):
        """One step of the hot path as main.c:31-75 + :215-221 runs it (assemble F and J with weak-BC faces and Dirichlet rows,
        then the GMRES solve): the checker of the bench-embedded parity test and of tests/test_gpu_parity.py."""
        ctx = self.driver_setup(mesh)
        N = mesh.num_node
        rp, ci = ctx["pattern"]
        Z = ci.size
        F = np.zeros(6 * N)
        blocks = [np.zeros(9 * Z), np.zeros(3 * Z), np.zeros(3 * Z), np.zeros(Z)]
        self.assemble_tet(N, mesh.ien, mesh.xg, ctx["off"], ctx["ind"], wg, dwg, F=F)
        self.assemble_tet(N, mesh.ien, mesh.xg, ctx["off"], ctx["ind"], wg, dwg, pattern=(rp, ci), blocks=blocks)
        if mesh.num_bound > 4:
            f2e, forn = mesh.bound_faces(4)
            self.assemble_face(f2e, forn, N, mesh.ien, mesh.xg, ctx["color"], ctx["nc"], wg, dwg, F=F)
            self.assemble_face(f2e, forn, N, mesh.ien, mesh.xg, ctx["color"], ctx["nc"], wg, dwg, pattern=(rp, ci), blocks=blocks)
        F[4 * N:] = 0
        for b, t in self.BCS.items():
            if b < mesh.num_bound:
                self.dirichlet_vec(mesh.bound_nodes(b), np.array(t, np.int32), F)
                self.dirichlet_mat(mesh.bound_nodes(b), np.array(t, np.int32), N, (rp, ci), blocks[0], blocks[1])
        out = dict(pattern=(rp, ci), F=F, blocks=blocks)
        if solve:
            out["dx"], out["iters"], out["hist"] = self.gmres((rp, ci), blocks, F)
        return out


# ----------------------------------------------------------------------------------------
# cuRAND host generator (libcurand is a CPU library for the *Host generator): the reference draws its
# coloring weights with CURAND_RNG_PSEUDO_DEFAULT (XORWOW), seed 1234 (color_impl.cu:225-237).
# ----------------------------------------------------------------------------------------
_CURAND = None


def curand_host_u32(n: int, seed: int = 1234) -> np.ndarray:
    global _CURAND
    if _CURAND is None:
        for name in ("libcurand.so.10", "/usr/local/cuda/lib64/libcurand.so.10", "libcurand.so"):
            try:
                _CURAND = C.CDLL(name)
                break
            except OSError:
                continue
        if _CURAND is None:
            raise OSError("libcurand not found")
    gen = C.c_void_p()
    CURAND_RNG_PSEUDO_DEFAULT = 100
    st = _CURAND.curandCreateGeneratorHost(C.byref(gen), CURAND_RNG_PSEUDO_DEFAULT)
    assert st == 0, f"curandCreateGeneratorHost -> {st}"
    st = _CURAND.curandSetPseudoRandomGeneratorSeed(gen, C.c_ulonglong(seed))
    assert st == 0
    out = np.zeros(n, np.uint32)
    st = _CURAND.curandGenerate(gen, out.ctypes.data_as(C.c_void_p), C.c_size_t(n))
    assert st == 0, f"curandGenerate -> {st}"
    _CURAND.curandDestroyGenerator(gen)
    return out


_ORACLE = None


def get() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
