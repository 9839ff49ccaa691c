"""numpy/scipy restatement of the opt-in two-level Schur-complement preconditioner (dedflow_b200/csrc/pc2.cu) and of
right-preconditioned GMRES with it.  TEST INFRASTRUCTURE ONLY (the checker of tests/test_gpu_pc2.py); the reference has no
counterpart -- its AMGX slot (src/pc.c:160-235) is compiled out -- so this pins OUR definition, written independently in numpy:

    u = D^-1 r_u ;  rt = r_p - A10 u ;  p = omega rt / diag(S) + P Cheb_m(Sc, P^T rt),   S = A11 - A10 D^-1 A01,  Sc = P^T S P
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def aggregates(xg, agg_cells):
    """cells of agg_cells average node spacings; ids = rank of the occupied cell keys (pc2.cu dfb_pc2_create)"""
    xg = np.asarray(xg, np.float64).reshape(-1, 3)
    N = xg.shape[0]
    lo, hi = xg.min(axis=0), xg.max(axis=0)
    ext = hi - lo
    live = ext > 0
    h = (np.prod(ext[live]) / N) ** (1.0 / live.sum()) if live.any() else 1.0
    cell = agg_cells * h
    nc = np.maximum(1, np.ceil(ext / cell - 1e-9).astype(np.int64))
    c = np.minimum(nc - 1, np.floor((xg - lo) / cell).astype(np.int64))
    key = c[:, 0] + nc[0] * (c[:, 1] + nc[1] * c[:, 2])
    uniq, agg = np.unique(key, return_inverse=True)
    return agg.astype(np.int64), uniq.size


class Pc2Oracle:
    def __init__(self, O, mesh, pattern, blocks, agg_cells=4, cheb_degree=10, ratio=30.0, omega=0.7):
        rp, ci = pattern
        N = rp.size - 1
        self.N, self.deg, self.ratio, self.omega = N, cheb_degree, ratio, omega
        A00v, A01v, A10v, A11v = blocks
        r33, c33 = O.expand_block(rp, ci, 3, 3, fix_last=True)
        r31, c31 = O.expand_block(rp, ci, 3, 1, fix_last=True)
        r13, c13 = O.expand_block(rp, ci, 1, 3, fix_last=True)
        self.A00 = sp.csr_matrix((A00v, c33, r33), shape=(3 * N, 3 * N))
        self.A01 = sp.csr_matrix((A01v, c31, r31), shape=(3 * N, N))
        self.A10 = sp.csr_matrix((A10v, c13, r13), shape=(N, 3 * N))
        self.A11 = sp.csr_matrix((A11v, ci, rp), shape=(N, N))
        B = np.stack([self.A00[3 * i:3 * i + 3, 3 * i:3 * i + 3].toarray() for i in range(N)])
        self.Binv = np.linalg.inv(B)
        Dinv = sp.bsr_matrix((self.Binv, np.arange(N), np.arange(N + 1)), shape=(3 * N, 3 * N)).tocsr()
        S = (self.A11 - self.A10 @ (Dinv @ self.A01)).tocsr()
        self.dS = S.diagonal()
        agg, Nc = aggregates(mesh.xg, agg_cells)
        self.agg, self.Nc = agg, Nc
        self.P = sp.csr_matrix((np.ones(N), (np.arange(N), agg)), shape=(N, Nc))
        self.Sc = (self.P.T @ S @ self.P).tocsr()
        self.dSc = self.Sc.diagonal()
        self.lam = float((abs(self.Sc).sum(axis=1).A1 / np.abs(self.dSc)).max())      # Gershgorin bound of dSc^-1 Sc

    def cheb(self, rc):
        lmx = self.lam
        lmn = lmx / self.ratio
        theta, delta = 0.5 * (lmx + lmn), 0.5 * (lmx - lmn)
        sigma = theta / delta
        rho = 1.0 / sigma
        z = np.zeros_like(rc)
        res = rc.copy()
        d = (res / self.dSc) / theta
        for k in range(self.deg):
            z += d
            if k == self.deg - 1:
                break
            res = res - self.Sc @ d
            rho_new = 1.0 / (2.0 * sigma - rho)
            d = rho_new * rho * d + 2.0 * rho_new / delta * (res / self.dSc)
            rho = rho_new
        return z

    def apply(self, x):
        """6N ABI-layout vector -> 6N (rows [4N, 6N) copied through)"""
        N = self.N
        y = np.array(x, dtype=np.float64)
        u = np.einsum("nij,nj->ni", self.Binv, x[:3 * N].reshape(N, 3)).ravel()
        rt = x[3 * N:4 * N] - self.A10 @ u
        y[:3 * N] = u
        y[3 * N:4 * N] = self.omega * rt / self.dS + self.P @ self.cheb(self.P.T @ rt)
        return y

    def gmres(self, b, maxit=120, atol=1e-12, rtol=1e-4):
        """right-preconditioned GMRES on the live 4N rows with the reference's every-20th-iteration stopping rule; modified
        Gram-Schmidt (the iteration COUNT and the residual norms are what is compared, not round-off)"""
        N = self.N
        A = sp.bmat([[self.A00, self.A01], [self.A10, self.A11]]).tocsr()
        M = lambda v: self.apply(np.concatenate([v, np.zeros(2 * N)]))[:4 * N]
        r0 = b[:4 * N].copy()
        beta0 = np.linalg.norm(r0)
        Q = [r0 / beta0]
        H = np.zeros((maxit + 1, maxit))
        g = np.zeros(maxit + 1)
        g[0] = beta0
        cs, sn, hist = [], [], [beta0]
        it = 0
        while it < maxit:
            w = A @ M(Q[it])
            for j in range(it + 1):
                H[j, it] = Q[j] @ w
                w = w - H[j, it] * Q[j]
            H[it + 1, it] = np.linalg.norm(w)
            Q.append(w / H[it + 1, it])
            for i in range(it):
                t = cs[i] * H[i, it] + sn[i] * H[i + 1, it]
                H[i + 1, it] = cs[i] * H[i + 1, it] - sn[i] * H[i, it]
                H[i, it] = t
            rr = np.hypot(H[it, it], H[it + 1, it])
            c, s = H[it, it] / rr, H[it + 1, it] / rr
            cs.append(c)
            sn.append(s)
            H[it, it], H[it + 1, it] = rr, 0.0
            g[it + 1] = -s * g[it]
            g[it] *= c
            hist.append(abs(g[it + 1]))
            it += 1
            if it % 20 == 0 and (hist[-1] < atol or hist[-1] < (hist[0] + 1e-16) * rtol):
                break
        y = np.linalg.solve(np.triu(H[:it, :it]), g[:it])
        x = M(sum(yj * qj for yj, qj in zip(y, Q[:it])))
        return np.concatenate([x, np.zeros(2 * N)]), it, np.array(hist)
