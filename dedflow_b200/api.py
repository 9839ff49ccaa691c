"""Host-side driver of the core C ABI, mirroring what reference src/main.c does around the hot path.

The method names follow the reference entry points they stand for:

    FlowSystem.csr_attr_create()         CSRAttrCreate + CSRAttrCreateBlock x3      (main.c:377-380)
    FlowSystem.generate_color_batch()    Mesh3DGenerateColorBatch                   (main.c:411, Mesh.c:165-206)
    FlowSystem.assemble_system(F=, J=)   AssembleSystem                             (main.c:31-75)
    FlowSystem.matrix_matvec(x, y)       MatrixMatVec                               (matrix.c:521-524)
    FlowSystem.krylov_solve(dx, F)       KrylovSolve(ksp, J, dx, F)                 (main.c:217, krylov.c:386-456)

torch is used for device memory and streams only (plumbing); every computation is a call into
libdedflow_b200.so.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as _lib

MODE = {"auto": 0, "gather": 1, "atomic": 2, "colored": 3}

# boundary conditions of the reference driver (main.c:454-476): boundary id -> BCType per velocity component
DEFAULT_BCS = {0: (1, 1, 1), 2: (0, 1, 0), 3: (0, 0, 1), 4: (0, 0, 0)}
WEAK_BC_GROUP = 4   # assemble.cu:1825-1828 (defect D13)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class FlowSystem:
    def __init__(self, mesh, device="cuda:0", bcs=None, max_iter=120, atol=1e-12, rtol=1e-4, with_colors=True):
        if not torch.cuda.is_available():
            raise _lib.DfbError("no CUDA device: dedflow_b200 has no CPU fallback")
        self.L = _lib.load()
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.mesh = mesh
        self.N, self.E = mesh.num_node, mesh.num_tet
        self.ien = torch.from_numpy(np.ascontiguousarray(mesh.ien.reshape(-1))).to(self.dev)
        self.xg = torch.from_numpy(np.ascontiguousarray(mesh.xg.reshape(-1))).to(self.dev)
        self.bcs = dict(DEFAULT_BCS if bcs is None else bcs)
        self.bnode = {b: torch.from_numpy(np.ascontiguousarray(mesh.bound_nodes(b))).to(self.dev) for b in self.bcs}
        f2e, forn = mesh.bound_faces(WEAK_BC_GROUP) if mesh.num_bound > WEAK_BC_GROUP else (np.zeros(0, np.int32),) * 2
        self.f2e = torch.from_numpy(np.ascontiguousarray(f2e)).to(self.dev)
        self.forn = torch.from_numpy(np.ascontiguousarray(forn)).to(self.dev)
        self.max_iter, self.atol, self.rtol = max_iter, atol, rtol
        self.plan = None
        self.gmres = None
        self.color = None
        self.batch_offset = None
        self.batch_ind = None
        self.csr_attr_create()
        if with_colors:
            self.generate_color_batch()
        self._make_plan()

    # ------------------------------------------------------------------ setup
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def csr_attr_create(self):
        L, N, E = self.L, self.N, self.E
        self.row_ptr = torch.empty(N + 1, dtype=torch.int32, device=self.dev)
        nnz = C.c_int(0)
        _lib.check(L.dfb_pattern_rows(N, E, _p(self.ien), _p(self.row_ptr), C.byref(nnz), self._stream()), "dfb_pattern_rows")
        self.nnz = nnz.value
        self.col_ind = torch.empty(self.nnz, dtype=torch.int32, device=self.dev)
        _lib.check(L.dfb_pattern_cols(N, E, _p(self.ien), _p(self.row_ptr), _p(self.col_ind), self._stream()), "dfb_pattern_cols")
        Z = self.nnz
        # the four sub-block value arrays of the field-split matrix (main.c:385-391)
        self.A00 = torch.zeros(9 * Z, dtype=torch.float64, device=self.dev)
        self.A01 = torch.zeros(3 * Z, dtype=torch.float64, device=self.dev)
        self.A10 = torch.zeros(3 * Z, dtype=torch.float64, device=self.dev)
        self.A11 = torch.zeros(Z, dtype=torch.float64, device=self.dev)

    def csr_attr_create_block(self, br, bc):
        """CSRAttrCreateBlock(spy1x1, br, bc) -> (row_ptr, col_ind) device tensors."""
        N, Z = self.N, self.nnz
        nrp = torch.empty(N * br + 1, dtype=torch.int32, device=self.dev)
        nci = torch.empty(Z * br * bc, dtype=torch.int32, device=self.dev)
        _lib.check(self.L.dfb_pattern_expand(N, _p(self.row_ptr), _p(self.col_ind), br, bc, _p(nrp), _p(nci), self._stream()),
                   "dfb_pattern_expand")
        return nrp, nci

    def generate_color_batch(self, seed=1234, max_color=256, weights=None):
        L, N, E = self.L, self.N, self.E
        if weights is None:
            self.weight = torch.empty(E, dtype=torch.int32, device=self.dev)
            _lib.check(L.dfb_color_weights(E, seed, _p(self.weight), self._stream()), "dfb_color_weights")
        else:
            self.weight = torch.as_tensor(weights, dtype=torch.int32).to(self.dev)
        self.color = torch.empty(E, dtype=torch.int32, device=self.dev)
        nc = C.c_int(0)
        _lib.check(L.dfb_color_jpl(N, E, _p(self.ien), _p(self.weight), max_color, _p(self.color), C.byref(nc), self._stream()),
                   "dfb_color_jpl")
        self.num_color = nc.value
        self.batch_offset = np.zeros(self.num_color + 1, np.int32)
        self.batch_ind = torch.empty(E, dtype=torch.int32, device=self.dev)
        _lib.check(L.dfb_color_batches(E, _p(self.color), self.num_color, self.batch_offset.ctypes.data_as(C.c_void_p),
                                       _p(self.batch_ind), self._stream()), "dfb_color_batches")

    def _make_plan(self):
        if self.plan is not None:
            self.L.dfb_plan_destroy(self.plan)
        plan = C.c_void_p()
        nb = 0 if self.batch_offset is None else self.batch_offset.size - 1
        _lib.check(self.L.dfb_plan_create(C.byref(plan), self.N, self.E, _p(self.ien), _p(self.row_ptr), _p(self.col_ind), nb,
                                          None if nb == 0 else self.batch_offset.ctypes.data_as(C.c_void_p),
                                          None if nb == 0 else _p(self.batch_ind), self._stream()), "dfb_plan_create")
        self.plan = plan

    # ------------------------------------------------------------------ hot path
    def blocks(self):
        return self.A00, self.A01, self.A10, self.A11

    def assemble_system(self, wgalpha, dwgalpha, F=None, J=False, mode="auto", faces=True, dirichlet=True):
        """AssembleSystem(mesh, wgalpha, dwgalpha, F|NULL, J|NULL, bcs, nbc) of main.c:31-75."""
        L, N, st = self.L, self.N, self._stream()
        m = MODE[mode]
        gather = m in (0, 1)
        A = self.blocks() if J else (None,) * 4
        if not gather:                      # reference flow: zero, then accumulate (main.c:44-49)
            if F is not None:
                F.zero_()
            if J:
                for a in A:
                    a.zero_()
        _lib.check(L.dfb_assemble_tet(self.plan, _p(self.xg), _p(wgalpha), _p(dwgalpha), _p(F), *[_p(a) for a in A], m,
                                      1 if gather else 0, st), "dfb_assemble_tet")
        if faces and self.f2e.numel():
            _lib.check(L.dfb_assemble_face(self.plan, self.f2e.numel(), _p(self.f2e), _p(self.forn), _p(self.xg), _p(wgalpha),
                                           _p(dwgalpha), _p(F), *[_p(a) for a in A], st), "dfb_assemble_face")
        if F is not None:
            F[4 * N:].zero_()               # main.c:63-66: phi / T residuals are discarded
        if dirichlet:
            for b, types in self.bcs.items():
                t = (C.c_int * 3)(*types)
                if F is not None:
                    _lib.check(L.dfb_dirichlet_vec(self.bnode[b].numel(), _p(self.bnode[b]), 3, t, _p(F), st), "dfb_dirichlet_vec")
                if J:
                    _lib.check(L.dfb_dirichlet_mat(self.bnode[b].numel(), _p(self.bnode[b]), 3, t, N, _p(self.row_ptr),
                                                   _p(self.col_ind), _p(self.A00), _p(self.A01), st), "dfb_dirichlet_mat")

    def matrix_amvpby(self, alpha, x, beta, y):
        _lib.check(self.L.dfb_spmv_fs(self.N, _p(self.row_ptr), _p(self.col_ind), *[_p(a) for a in self.blocks()], alpha, _p(x),
                                      beta, _p(y), self._stream()), "dfb_spmv_fs")

    def matrix_matvec(self, x, y):
        self.matrix_amvpby(1.0, x, 0.0, y)

    def pc_setup(self):
        self.dinv00 = torch.empty(9 * self.N, dtype=torch.float64, device=self.dev)
        self.dinv11 = torch.empty(self.N, dtype=torch.float64, device=self.dev)
        _lib.check(self.L.dfb_pc_setup(self.N, _p(self.row_ptr), _p(self.col_ind), _p(self.A00), _p(self.A11), _p(self.dinv00),
                                       _p(self.dinv11), self._stream()), "dfb_pc_setup")

    def pc_apply(self, x, y):
        _lib.check(self.L.dfb_pc_apply(self.N, _p(self.dinv00), _p(self.dinv11), _p(x), _p(y), self._stream()), "dfb_pc_apply")

    def _workspace(self):
        if self.gmres is None:
            ws = C.c_void_p()
            _lib.check(self.L.dfb_gmres_create(C.byref(ws), self.N, self.max_iter), "dfb_gmres_create")
            self.gmres = ws
        return self.gmres

    def set_preconditioner(self, kind="jacobi", agg_cells=4, cheb_degree=10):
        """"jacobi": the reference's block-Jacobi (default, krylov.c:439-452).  "schur2": the opt-in two-level Schur-complement
        preconditioner (the slot the reference reserves for AMGX on the pressure block, pc.c:160-235; csrc/pc2.cu)."""
        ws = self._workspace()
        if kind == "jacobi":
            _lib.check(self.L.dfb_gmres_set_pc2(ws, None), "dfb_gmres_set_pc2")
            return
        if kind != "schur2":
            raise ValueError(kind)
        if getattr(self, "pc2", None) is None:
            pc = C.c_void_p()
            _lib.check(self.L.dfb_pc2_create(C.byref(pc), self.N, _p(self.row_ptr), _p(self.col_ind), _p(self.xg), agg_cells,
                                             cheb_degree, self._stream()), "dfb_pc2_create")
            self.pc2 = pc
        _lib.check(self.L.dfb_gmres_set_pc2(ws, self.pc2), "dfb_gmres_set_pc2")

    def krylov_solve(self, dx, F):
        """KrylovSolve(ksp, J, dx, F).  Returns (iterations, residual history |beta_k|, k = 0..iterations)."""
        self._workspace()
        iters = C.c_int(0)
        hist = np.zeros(self.max_iter + 1)
        _lib.check(self.L.dfb_gmres_solve(self.gmres, self.N, _p(self.row_ptr), _p(self.col_ind), *[_p(a) for a in self.blocks()],
                                          _p(dx), _p(F), self.atol, self.rtol, C.byref(iters), hist.ctypes.data_as(C.c_void_p),
                                          self._stream()), "dfb_gmres_solve")
        return iters.value, hist[:iters.value + 1]

    # ------------------------------------------------------------------ the driver around the path
    def solve_flow_system(self, wgold, dwgold, dwg, maxit=4, tol=0.5e-3, mode="auto"):
        """SolveFlowSystem of main.c:77-283: Newton iteration on the alpha-level states.  dwg is updated in place.
        Returns a list of (rnorm[4], gmres_iterations) per Newton iteration, entry 0 = initial norms."""
        L, N, st = self.L, self.N, self._stream()
        if getattr(self, "_newton_ws", None) is None:
            mk = lambda: torch.zeros(6 * N, dtype=torch.float64, device=self.dev)
            self._newton_ws = (mk(), mk(), mk(), mk())
        wgalpha, dwgalpha, F, dx = self._newton_ws
        norms = (C.c_double * 4)()

        def stage():
            _lib.check(L.dfb_genalpha_stage(N, _p(wgold), _p(dwgold), _p(dwg), _p(wgalpha), _p(dwgalpha), st), "dfb_genalpha_stage")

        def residual():
            self.assemble_system(wgalpha, dwgalpha, F=F, mode=mode)
            _lib.check(L.dfb_block_norms(N, _p(F), norms, st), "dfb_block_norms")
            return np.array(norms[:])

        stage()
        r0 = residual()
        hist = [(r0.copy(), 0)]
        r0 = r0 + 1e-16                                   # main.c:153-156
        it, converged = 0, False
        while not converged and it < maxit:
            self.assemble_system(wgalpha, dwgalpha, J=True, mode=mode)
            dx.zero_()
            its, _ = self.krylov_solve(dx, F)
            _lib.check(L.dfb_newton_update(N, _p(dx), _p(dwg), st), "dfb_newton_update")     # main.c:226
            stage()                                                                          # main.c:232-246
            r = residual()
            hist.append((r.copy(), its))
            converged = bool(np.all(r < tol * r0))                                           # main.c:271-276
            it += 1
        return hist

    def time_step(self, wgold, dwgold, dwg, **kw):
        """One pass of the time loop of main.c:537-565: predictor, Newton solve, corrector.  All three vectors are updated."""
        L, N, st = self.L, self.N, self._stream()
        _lib.check(L.dfb_genalpha_predict(N, _p(dwg), st), "dfb_genalpha_predict")
        hist = self.solve_flow_system(wgold, dwgold, dwg, **kw)
        _lib.check(L.dfb_genalpha_correct(N, _p(wgold), _p(dwgold), _p(dwg), st), "dfb_genalpha_correct")
        return hist

    def close(self):
        if self.plan is not None:
            self.L.dfb_plan_destroy(self.plan)
            self.plan = None
        if self.gmres is not None:
            self.L.dfb_gmres_destroy(self.gmres)
            self.gmres = None
        if getattr(self, "pc2", None) is not None:
            self.L.dfb_pc2_destroy(self.pc2)
            self.pc2 = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
