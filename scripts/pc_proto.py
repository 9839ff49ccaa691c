"""CPU prototype bench for the opt-in preconditioner (csrc/pc2.cu): iteration counts of right-preconditioned GMRES to rtol 1e-4
on the first Newton system of the first time step (the reference's initial condition), for a list of mesh sizes and
preconditioner variants, all in numpy/scipy on top of the C oracle's assembly.  Design tool only (nothing here is shipped);
the trends it prints are what pc2.cu's structure was chosen from.
usage: python scripts/pc_proto.py [m ...]"""
import sys
import time
from pathlib import Path

import numpy as np
import scipy.sparse as sp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from dedflow_b200 import boxmesh  # noqa: E402
from oracle import pc2_oracle, pyoracle  # noqa: E402


def gmres_count(A, M, b, rtol=1e-4, maxit=200):
    """iterations of right-preconditioned GMRES (no restart, MGS) until |r| < rtol |r0|"""
    beta0 = np.linalg.norm(b)
    Q = [b / beta0]
    H = np.zeros((maxit + 1, maxit))
    g = np.zeros(maxit + 1)
    g[0] = beta0
    cs, sn = [], []
    for it in range(maxit):
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
        H[it, it] = rr
        g[it + 1] = -s * g[it]
        g[it] *= c
        if abs(g[it + 1]) < rtol * beta0:
            return it + 1
    return maxit


def system(O, m):
    mesh = boxmesh.make_box(m)
    N = mesh.num_node
    ctx = O.driver_setup(mesh)
    wgold, dwgold, dwg = boxmesh.state_initial(mesh)
    fac = (O.K_GAMMA - 1.0) / O.K_GAMMA
    dwg[:3 * N] *= fac
    dwg[4 * N:] *= fac
    wga, dwga = O.alpha_states(N, wgold, dwgold, dwg)
    F = O.assemble_system(ctx, wga, dwga, want_F=True)
    blocks = O.assemble_system(ctx, wga, dwga, want_J=True)
    return mesh, ctx, F, blocks


def variants(O, mesh, ctx, blocks):
    N = mesh.num_node
    R = pc2_oracle.Pc2Oracle(O, mesh, ctx["pattern"], blocks, agg_cells=4, cheb_degree=10)
    Dinv = sp.bsr_matrix((R.Binv, np.arange(N), np.arange(N + 1)), shape=(3 * N, 3 * N)).tocsr()
    S = (R.A11 - R.A10 @ (Dinv @ R.A01)).tocsr()
    d00, d11 = O.pc_setup(ctx["pattern"], blocks)

    def jacobi(v):                                  # the reference's block-Jacobi (with its transposed inverse, defect D3)
        return O.pc_apply(d00, d11, np.concatenate([v, np.zeros(2 * N)]))[:4 * N]

    def lower(v, schur):                            # block lower-triangular frame around a Schur-complement solve
        u = Dinv @ v[:3 * N]
        rt = v[3 * N:] - R.A10 @ u
        return u, schur(rt)

    def additive(rt):                               # what pc2.cu ships
        return R.omega * rt / R.dS + R.P @ R.cheb(R.P.T @ rt)

    def multiplicative(rt, post=True):              # smooth, coarse-correct the remaining residual, smooth again
        p = R.omega * rt / R.dS
        p = p + R.P @ R.cheb(R.P.T @ (rt - S @ p))
        if post:
            p = p + R.omega * (rt - S @ p) / R.dS
        return p

    def v_add(v):
        u, p = lower(v, additive)
        return np.concatenate([u, p])

    def v_mult(v):
        u, p = lower(v, multiplicative)
        return np.concatenate([u, p])

    def v_add_upper(v):                             # + the upper-triangular correction u -= D^-1 A01 p (SIMPLE proper)
        u, p = lower(v, additive)
        return np.concatenate([u - Dinv @ (R.A01 @ p), p])

    def v_mult_upper(v):
        u, p = lower(v, multiplicative)
        return np.concatenate([u - Dinv @ (R.A01 @ p), p])

    def v_add_u2(v):                                # two Jacobi sweeps on the velocity block instead of one
        r = v[:3 * N]
        u = Dinv @ r
        u = u + Dinv @ (r - R.A00 @ u)
        rt = v[3 * N:] - R.A10 @ u
        return np.concatenate([u, additive(rt)])

    # smoothed prolongation: P_s = (I - w dS^-1 S) P
    Ps = (R.P - sp.diags(0.66 / R.dS) @ (S @ R.P)).tocsr()
    Scs = (Ps.T @ S @ Ps).tocsr()
    import scipy.sparse.linalg as spla
    lu = spla.splu(Scs.tocsc())

    def v_sa(v):                                    # smoothed aggregation, exact coarse solve, additive
        u, p = lower(v, lambda rt: R.omega * rt / R.dS + Ps @ lu.solve(Ps.T @ rt))
        return np.concatenate([u, p])

    luc = spla.splu(R.Sc.tocsc())

    def v_add_exact(v):                             # plain aggregation, exact coarse solve, additive (upper bound of the Chebyshev)
        u, p = lower(v, lambda rt: R.omega * rt / R.dS + R.P @ luc.solve(R.P.T @ rt))
        return np.concatenate([u, p])

    # (sparse LU of the fine-level operators: minutes beyond ~30k nodes -- limit studies on small meshes only)
    luS = spla.splu(S.tocsc()) if N <= 30000 else None
    luA = spla.splu(R.A00.tocsc()) if N <= 30000 else None

    def v_exactS(v):                                # limit study: exact solve with S = A11 - A10 D^-1 A01 (velocity part as shipped)
        u, p = lower(v, luS.solve)
        return np.concatenate([u, p])

    def v_exactA(v):                                # limit study: exact velocity solve, pressure part as shipped
        u = luA.solve(v[:3 * N])
        return np.concatenate([u, additive(v[3 * N:] - R.A10 @ u)])

    def v_exactS_upper(v):
        u, p = lower(v, luS.solve)
        return np.concatenate([u - Dinv @ (R.A01 @ p), p])

    limits = []
    if luS is not None:
        limits = [("limit: exact S solve", v_exactS), ("limit: exact S solve + upper", v_exactS_upper),
                  ("limit: exact A00 solve", v_exactA)]

    def cheb_variant(ratio, deg):                   # the shipped structure with another Chebyshev interval / degree
        def f(v):
            keep = (R.ratio, R.deg)
            R.ratio, R.deg = ratio, deg
            try:
                return v_add(v)
            finally:
                R.ratio, R.deg = keep
        return f

    extra = [(f"additive, Chebyshev ratio {r} degree {d}", cheb_variant(r, d)) for r, d in ((100, 20), (300, 40), (1000, 60))]
    A = sp.bmat([[R.A00, R.A01], [R.A10, R.A11]]).tocsr()
    return A, [("block-Jacobi (reference)", jacobi), ("pc2: additive two-level", v_add), ("additive, exact coarse solve", v_add_exact),
               ("multiplicative + post-smooth", v_mult), ("additive + upper correction", v_add_upper),
               ("multiplicative + upper", v_mult_upper), ("additive, 2 velocity sweeps", v_add_u2),
               ("smoothed aggregation, exact coarse", v_sa)] + extra + limits


if __name__ == "__main__":
    ms = [int(a) for a in sys.argv[1:]] or [12, 20, 28]
    O = pyoracle.get()
    for m in ms:
        t0 = time.time()
        mesh, ctx, F, blocks = system(O, m)
        A, vs = variants(O, mesh, ctx, blocks)
        b = F[:4 * mesh.num_node]
        print(f"m={m}: {mesh.num_tet} tets, {mesh.num_node} nodes (setup {time.time() - t0:.1f} s)", flush=True)
        for name, M in vs:
            if "velocity sweeps" in name and m > 20:
                continue                            # diverges on finer meshes (the sweeps amplify the convective part)
            t0 = time.time()
            print(f"   {name:38s} {gmres_count(A, M, b):4d} iterations  ({time.time() - t0:.1f} s)", flush=True)
