"""ctypes driver of oracle/_ref/libdedflow_ref.so -- the reference's OWN CUDA implementation of the hot path,
built unmodified from /root/reference/src by oracle/ref/Makefile.

TEST INFRASTRUCTURE ONLY.  Used to (a) pin the CPU oracle and the B200 kernels against the real reference on a
GPU box, (b) generate tests/golden/ref_*.npz (oracle/ref/run_ref.py), (c) time the reference beside ours
(bench.py --impl reference).  Never imported by the product package.

The driver mirrors what reference src/main.c does around the hot path (main.c:31-75 AssembleSystem,
main.c:377-411 setup, main.c:454-476 boundary conditions, main.c:211-219 solve) with the reference's own
structs (Mesh.h:14-45, MeshData.h:10-19, csr.h:12-20, matrix.h:63-103, dirichlet.h:19-27).
Needs torch (device buffers) and a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SO = HERE.parent / "_ref" / "libdedflow_ref.so"
# the reference's remaining host code (Mesh.c MeshData.c common.c alloc.c) linked against the PRODUCT library:
# the same driver below then exercises dedflow_b200's drop-in entry points (include/dedflow_compat.h)
HYBRID_SO = HERE.parent / "_ref" / "libdedflow_hybrid.so"

i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


class Mesh3DData(C.Structure):
    _fields_ = [("is_host", C.c_int32), ("num_node", C.c_int32), ("num_tet", C.c_int32), ("num_prism", C.c_int32),
                ("num_hex", C.c_int32), ("xg", C.c_void_p), ("ien", C.c_void_p)]


class Mesh3D(C.Structure):
    _fields_ = [("num_node", C.c_int32), ("num_tet", C.c_int32), ("num_prism", C.c_int32), ("num_hex", C.c_int32),
                ("host", C.POINTER(Mesh3DData)), ("device", C.POINTER(Mesh3DData)),
                ("num_bound", C.c_int32),
                ("bound_fid", C.c_void_p), ("bound_node_offset", C.c_void_p), ("bound_node", C.c_void_p),
                ("bound_elem_offset", C.c_void_p), ("bound_ien", C.c_void_p), ("bound_f2e", C.c_void_p),
                ("bound_forn", C.c_void_p),
                ("num_batch", C.c_int32), ("batch_offset", C.c_void_p), ("batch_ind", C.c_void_p),
                ("num_color", C.c_int32), ("color", C.c_void_p)]


class CSRAttr(C.Structure):
    pass


CSRAttr._fields_ = [("num_row", C.c_int32), ("num_col", C.c_int32), ("nnz", C.c_int32), ("row_ptr", C.c_void_p),
                    ("col_ind", C.c_void_p), ("parent", C.POINTER(CSRAttr))]


class Matrix(C.Structure):
    _fields_ = [("size", C.c_int32 * 2), ("type", C.c_int), ("data", C.c_void_p), ("stream_ref", C.c_void_p),
                ("op", C.c_void_p * 15)]


class MatrixCSR(C.Structure):
    _fields_ = [("external_attr", C.c_int32), ("attr", C.POINTER(CSRAttr)), ("val", C.c_void_p), ("descr", C.c_void_p),
                ("buffer_size", C.c_int32), ("buffer", C.c_void_p)]


class MatrixFS(C.Structure):
    _fields_ = [("n_offset", C.c_int32), ("offset", C.c_void_p), ("d_offset", C.c_void_p), ("stream", C.c_void_p),
                ("spy1x1", C.POINTER(CSRAttr)), ("d_matval", C.c_void_p), ("mat", C.POINTER(C.POINTER(Matrix)))]


class Dirichlet(C.Structure):
    _fields_ = [("mesh", C.c_void_p), ("face_ind", C.c_int32), ("shape", C.c_int32), ("buffer_size", C.c_size_t),
                ("buffer", C.c_void_p)]          # followed by BCType bctype[shape]


def available() -> bool:
    return REF_SO.exists()


def hybrid_available() -> bool:
    return HYBRID_SO.exists()


class CaptureStdout:
    """Capture what C code printf()s to fd 1 (the reference reports its GMRES residuals only there)."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.tmp = tempfile.TemporaryFile(mode="w+b")
        self.saved = os.dup(1)
        os.dup2(self.tmp.fileno(), 1)
        self.text = ""
        return self

    def __exit__(self, *a):
        self.lib_flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        self.tmp.seek(0)
        self.text = self.tmp.read().decode(errors="replace")
        self.tmp.close()

    @staticmethod
    def lib_flush():
        C.CDLL(None).fflush(None)


class RefProblem:
    """One mesh set up exactly as reference main.c does, driven through the reference's own entry points."""

    def __init__(self, mesh, patch_d1: bool = True, quiet: bool = True, so_path=None):
        import torch
        self.torch = torch
        self.quiet = quiet
        L = C.CDLL(str(so_path or REF_SO), mode=C.RTLD_LOCAL)
        self.L = L
        self.mesh_np = mesh
        N, E = mesh.num_node, mesh.num_tet
        self.N, self.E = N, E
        dev = torch.device("cuda", 0)
        self.dev = dev
        L.Mesh3DCreate.restype = C.POINTER(Mesh3D)
        L.Mesh3DCreate.argtypes = [C.c_int32] * 4
        L.CSRAttrCreate.restype = C.POINTER(CSRAttr)
        L.CSRAttrCreate.argtypes = [C.POINTER(Mesh3D)]
        L.CSRAttrCreateBlock.restype = C.POINTER(CSRAttr)
        L.CSRAttrCreateBlock.argtypes = [C.POINTER(CSRAttr), C.c_int32, C.c_int32]
        L.MatrixCreateTypeFS.restype = C.POINTER(Matrix)
        L.MatrixCreateTypeFS.argtypes = [C.c_int32, i32p, C.c_void_p]
        L.MatrixCreateTypeCSR.restype = C.POINTER(Matrix)
        L.MatrixCreateTypeCSR.argtypes = [C.POINTER(CSRAttr), C.c_void_p]
        L.MatrixSetup.argtypes = [C.POINTER(Matrix)]
        L.MatrixZero.argtypes = [C.POINTER(Matrix)]
        L.MatrixMatVec.argtypes = [C.POINTER(Matrix), C.c_void_p, C.c_void_p]
        L.KrylovCreateGMRES.restype = C.c_void_p
        L.KrylovCreateGMRES.argtypes = [C.c_int32, C.c_double, C.c_double, C.c_void_p]
        L.KrylovSolve.argtypes = [C.c_void_p, C.POINTER(Matrix), C.c_void_p, C.c_void_p]
        L.Mesh3DGenerateColorBatch.argtypes = [C.POINTER(Mesh3D)]
        L.Mesh3DUpdateDevice.argtypes = [C.POINTER(Mesh3D)]
        L.DirichletCreate.restype = C.POINTER(Dirichlet)
        L.DirichletCreate.argtypes = [C.POINTER(Mesh3D), C.c_int32, C.c_int32]
        L.DirichletApplyVec.argtypes = [C.POINTER(Dirichlet), C.c_void_p]
        L.DirichletApplyMat.argtypes = [C.POINTER(Dirichlet), C.POINTER(Matrix)]
        L.AssembleSystemTet.argtypes = [C.POINTER(Mesh3D), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.AssembleSystemTetFace.argtypes = [C.POINTER(Mesh3D), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.Init.argtypes = [C.c_int, C.c_void_p]

        with self._q():
            L.Init(0, None)
            # --- mesh (what Mesh3DCreateH5 + ReadBoundFromH5Private do, Mesh.c:12-104, without the file) ---
            self.mesh = L.Mesh3DCreate(N, E, 0, 0)
            mp = self.mesh.contents
            C.memmove(mp.host.contents.xg, mesh.xg.ctypes.data, mesh.xg.nbytes)
            C.memmove(mp.host.contents.ien, mesh.ien.ctypes.data, mesh.ien.nbytes)
            L.Mesh3DUpdateDevice(self.mesh)
            nb = mesh.num_bound
            self._keep = []
            offs = np.concatenate([mesh.bound_node_offset, mesh.bound_elem_offset]).astype(np.int32)
            self._keep.append(offs)
            mp.num_bound = nb
            mp.bound_node_offset = offs.ctypes.data
            mp.bound_elem_offset = offs.ctypes.data + 4 * (nb + 1)
            dev_bound = torch.from_numpy(np.concatenate([mesh.bound_node, mesh.bound_f2e, mesh.bound_forn])).to(dev)
            self._keep.append(dev_bound)
            mp.bound_node = dev_bound.data_ptr()
            mp.bound_f2e = dev_bound.data_ptr() + 4 * mesh.bound_node.size
            mp.bound_forn = dev_bound.data_ptr() + 4 * (mesh.bound_node.size + mesh.bound_f2e.size)
            # --- patterns + matrix (main.c:377-404) ---
            self.spy1x1 = L.CSRAttrCreate(self.mesh)
            self.spy1x3 = L.CSRAttrCreateBlock(self.spy1x1, 1, 3)
            self.spy3x1 = L.CSRAttrCreateBlock(self.spy1x1, 3, 1)
            self.spy3x3 = L.CSRAttrCreateBlock(self.spy1x1, 3, 3)
            torch.cuda.synchronize()
            self.patched_d1 = patch_d1
            if patch_d1:
                # defect D1 (csr_impl.cu:24-36 never writes the final row_ptr entry): the documented one-line
                # patch, applied from outside so the reference sources stay untouched.
                for a in (self.spy1x3, self.spy3x1, self.spy3x3):
                    ac = a.contents
                    last = torch.tensor([ac.nnz], dtype=torch.int32, device=dev)
                    self._copy_d2d(ac.row_ptr + 4 * ac.num_row, last.data_ptr(), 4)
            off = (C.c_int32 * 5)(0, 3, 4, 5, 6)
            self.J = L.MatrixCreateTypeFS(4, off, None)
            fs = C.cast(self.J.contents.data, C.POINTER(MatrixFS)).contents
            fs.spy1x1 = self.spy1x1
            fs.mat[0] = L.MatrixCreateTypeCSR(self.spy3x3, None)
            fs.mat[1] = L.MatrixCreateTypeCSR(self.spy3x1, None)
            fs.mat[4] = L.MatrixCreateTypeCSR(self.spy1x3, None)
            fs.mat[5] = L.MatrixCreateTypeCSR(self.spy1x1, None)
            self.fs = fs
            L.MatrixSetup(self.J)
            self.ksp = L.KrylovCreateGMRES(120, 1e-12, 1e-4, None)
            torch.cuda.synchronize()

    # ------------------------------------------------------------------
    def _q(self):
        import contextlib
        return CaptureStdout() if self.quiet else contextlib.nullcontext()

    @staticmethod
    def _cudart():
        for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12", "libcudart.so"):
            try:
                lib = C.CDLL(name)
                lib.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
                return lib
            except OSError:
                continue
        raise OSError("libcudart not found")

    def _copy_d2d(self, dst, src, nbytes):
        err = self._cudart().cudaMemcpy(dst, src, nbytes, 3)
        assert err == 0, f"cudaMemcpy D2D -> {err}"

    def _d2h(self, ptr, n, dtype):
        out = np.empty(n, dtype=dtype)
        err = self._cudart().cudaMemcpy(out.ctypes.data, ptr, out.nbytes, 2)
        assert err == 0, f"cudaMemcpy D2H -> {err}"
        return out

    def attr_arrays(self, attr):
        a = attr.contents
        return self._d2h(a.row_ptr, a.num_row + 1, np.int32), self._d2h(a.col_ind, a.nnz, np.int32)

    def patterns(self):
        return {"1x1": self.attr_arrays(self.spy1x1), "3x3": self.attr_arrays(self.spy3x3),
                "3x1": self.attr_arrays(self.spy3x1), "1x3": self.attr_arrays(self.spy1x3)}

    def color_batches(self):
        with self._q():
            self.L.Mesh3DGenerateColorBatch(self.mesh)
            self.torch.cuda.synchronize()
        mp = self.mesh.contents
        nc = mp.num_color
        color = self._d2h(mp.color, self.E, np.int32)
        batch_ind = self._d2h(mp.batch_ind, self.E, np.int32)
        batch_offset = np.ctypeslib.as_array(C.cast(mp.batch_offset, i32p), shape=(nc + 1,)).copy()
        return color, batch_offset, batch_ind, nc

    def make_bcs(self):
        """main.c:454-476: faces 0,2,3,4 with bctype rows {SSS, -S-, --S, ---}."""
        if getattr(self, "bcs", None):
            return
        spec = {0: (1, 1, 1), 2: (0, 1, 0), 3: (0, 0, 1), 4: (0, 0, 0)}
        self.bcs = []
        for face, types in spec.items():
            bc = self.L.DirichletCreate(self.mesh, face, 3)
            arr = C.cast(C.addressof(bc.contents) + C.sizeof(Dirichlet), C.POINTER(C.c_int))
            for ic in range(3):
                arr[ic] = types[ic]
            self.bcs.append(bc)

    def block_vals(self):
        out = []
        for idx in (0, 1, 4, 5):
            m = C.cast(self.fs.mat[idx].contents.data, C.POINTER(MatrixCSR)).contents
            out.append(self._d2h(m.val, m.attr.contents.nnz, np.float64))
        return out

    def assemble(self, wg_t, dwg_t, F_t=None, J=False, faces=True, dirichlet=True):
        """AssembleSystem of main.c:31-75.  wg_t/dwg_t/F_t are torch cuda f64 tensors of length 6N."""
        L, N = self.L, self.N
        Fp = F_t.data_ptr() if F_t is not None else None
        Jp = self.J if J else None
        self.make_bcs()
        with self._q():
            if F_t is not None:
                F_t.zero_()
            if J:
                L.MatrixZero(self.J)
            L.AssembleSystemTet(self.mesh, wg_t.data_ptr(), dwg_t.data_ptr(), Fp, C.cast(Jp, C.c_void_p) if J else None)
            if faces:
                L.AssembleSystemTetFace(self.mesh, wg_t.data_ptr(), dwg_t.data_ptr(), Fp,
                                        C.cast(Jp, C.c_void_p) if J else None)
            if F_t is not None:
                F_t[4 * N:].zero_()
            if dirichlet:
                for bc in self.bcs:
                    if F_t is not None:
                        L.DirichletApplyVec(bc, Fp)
                    if J:
                        L.DirichletApplyMat(bc, self.J)

    def matvec(self, x_t, y_t):
        with self._q():
            self.L.MatrixMatVec(self.J, x_t.data_ptr(), y_t.data_ptr())

    def solve(self, dx_t, F_t, maxit=120, atol=1e-12, rtol=1e-4):
        """KrylovSolve(ksp, J, dx, F) (definition order, defect D9).  Returns the captured residual printout."""
        class _K(C.Structure):
            _fields_ = [("max_iter", C.c_int32), ("atol", C.c_double), ("rtol", C.c_double)]
        k = C.cast(self.ksp, C.POINTER(_K)).contents
        k.max_iter, k.atol, k.rtol = maxit, atol, rtol
        cap = CaptureStdout()
        with cap:
            self.L.KrylovSolve(self.ksp, self.J, dx_t.data_ptr(), F_t.data_ptr())
            self.torch.cuda.synchronize()
        hist = []
        for line in cap.text.splitlines():
            line = line.strip()
            if ") abs =" in line:
                it = int(line.split(")")[0])
                val = float(line.split("abs =")[1].split("(")[0])
                hist.append((it, val))
        return hist
