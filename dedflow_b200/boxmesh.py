"""Synthetic structured tetrahedral box mesh (Kuhn 6-tet subdivision of the unit cube).

The reference ships no mesh; its driver reads ``box.h5`` written by
``tools/mesh_convert.py`` (reference tools/mesh_convert.py:116-126) with datasets
``mesh/xg``, ``mesh/ien/tet``, ``mesh/bound/{node_offset,node,elem_offset,ien,f2e,forn}``.
This module produces the same arrays for an ``m x m x m`` box without any file:

* ``E = 6 m^3`` tets, ``N = (m+1)^3`` nodes (m=20 -> 48,000 tets, BASELINE config 1);
* node id ``(k*(m+1) + j)*(m+1) + i``  (x fastest), ``xg[id] = (i, j, k)/m``;
* tets are positively oriented (det J > 0) so Nanson normals point outward;
* six boundary groups ``0:x=0 1:x=1 2:y=0 3:z=0 4:z=1 5:y=1`` with,
  per group, the sorted unique node list, the boundary faces' owning tet (``f2e``)
  and the local index of the vertex opposite the face (``forn``,
  reference tools/mesh_convert.py:57-66).

Host-side utility only (numpy); nothing here is on the GPU hot path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from itertools import permutations

import numpy as np

I32 = np.int32
F64 = np.float64

# (axis, side) for boundary groups 0..5
BOUND_PLANES = ((0, 0), (0, 1), (1, 0), (2, 0), (2, 1), (1, 1))


@dataclass
class BoxMesh:
    m: int
    num_node: int
    num_tet: int
    xg: np.ndarray            # f64 [N,3]
    ien: np.ndarray           # i32 [E,4]
    bound_node_offset: np.ndarray   # i32 [nb+1]
    bound_node: np.ndarray          # i32 [sum nodes]
    bound_elem_offset: np.ndarray   # i32 [nb+1]
    bound_f2e: np.ndarray           # i32 [sum faces]
    bound_forn: np.ndarray          # i32 [sum faces]
    bound_ien: np.ndarray           # i32 [sum faces, 3]
    extra: dict = field(default_factory=dict)

    @property
    def num_bound(self) -> int:
        return len(self.bound_node_offset) - 1

    def bound_nodes(self, b: int) -> np.ndarray:
        return self.bound_node[self.bound_node_offset[b]:self.bound_node_offset[b + 1]]

    def bound_faces(self, b: int):
        s, e = self.bound_elem_offset[b], self.bound_elem_offset[b + 1]
        return self.bound_f2e[s:e], self.bound_forn[s:e]


def _perm_parity(p) -> int:
    p = list(p)
    s = 0
    for i in range(len(p)):
        for j in range(i + 1, len(p)):
            s += p[i] > p[j]
    return s & 1


def make_box(m: int, lengths=(1.0, 1.0, 1.0)) -> BoxMesh:
    """Kuhn box mesh with ``m`` cells per direction."""
    assert m >= 1
    n1 = m + 1
    num_node = n1 ** 3
    # coordinates, x fastest
    ax = np.arange(n1, dtype=F64) / m
    kk, jj, ii = np.meshgrid(ax * lengths[2], ax * lengths[1], ax * lengths[0], indexing="ij")
    xg = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(F64)

    # cells, x fastest as well
    ci = np.arange(m, dtype=np.int64)
    ck, cj, cii = np.meshgrid(ci, ci, ci, indexing="ij")
    base = ((ck * n1 + cj) * n1 + cii).ravel()          # node id of cell corner (i,j,k)
    stride = np.array([1, n1, n1 * n1], dtype=np.int64)

    tets = []
    for p in permutations(range(3)):
        v0 = base
        v1 = v0 + stride[p[0]]
        v2 = v1 + stride[p[1]]
        v3 = v2 + stride[p[2]]
        if _perm_parity(p):
            # odd permutation -> negative orientation; swap two vertices
            t = np.stack([v0, v2, v1, v3], axis=1)
        else:
            t = np.stack([v0, v1, v2, v3], axis=1)
        tets.append(t)
    # element id = cell*6 + p : keeps the 6 tets of a cell adjacent
    ien = np.stack(tets, axis=1).reshape(-1, 4).astype(I32)
    num_tet = ien.shape[0]

    # boundary groups
    ijk = np.stack([np.arange(num_node) % n1,
                    (np.arange(num_node) // n1) % n1,
                    np.arange(num_node) // (n1 * n1)], axis=1)
    node_off, elem_off = [0], [0]
    nodes_all, f2e_all, forn_all, bien_all = [], [], [], []
    for axis, side in BOUND_PLANES:
        on = ijk[:, axis] == (m if side else 0)
        nodes = np.nonzero(on)[0].astype(I32)
        nodes_all.append(nodes)
        node_off.append(node_off[-1] + len(nodes))
        on_t = on[ien]                              # [E,4]
        cnt = on_t.sum(axis=1)
        te = np.nonzero(cnt == 3)[0]
        forn = np.argmin(on_t[te], axis=1)          # the one vertex off the plane
        f2e_all.append(te.astype(I32))
        forn_all.append(forn.astype(I32))
        tri = np.stack([ien[te, (forn + 1) % 4], ien[te, (forn + 2) % 4], ien[te, (forn + 3) % 4]], axis=1)
        bien_all.append(tri.astype(I32))
        elem_off.append(elem_off[-1] + len(te))
    return BoxMesh(
        m=m, num_node=num_node, num_tet=num_tet, xg=np.ascontiguousarray(xg), ien=np.ascontiguousarray(ien),
        bound_node_offset=np.array(node_off, dtype=I32),
        bound_node=np.concatenate(nodes_all).astype(I32),
        bound_elem_offset=np.array(elem_off, dtype=I32),
        bound_f2e=np.concatenate(f2e_all).astype(I32),
        bound_forn=np.concatenate(forn_all).astype(I32),
        bound_ien=np.concatenate(bien_all).astype(I32),
    )


# ----------------------------------------------------------------------------------------
# nodal states
# ----------------------------------------------------------------------------------------
K_RHOC = 0.5
K_DT = 5e-2
K_ALPHAM = (3.0 - K_RHOC) / (1.0 + K_RHOC)
K_ALPHAF = 1.0 / (1.0 + K_RHOC)
K_GAMMA = 0.5 + K_ALPHAM - K_ALPHAF


def state_random(num_node: int, seed: int = 1234):
    """State B: uniform(-1,1) nodal states with the slot rules of reference main.c:112,118.

    Layout of every 6N vector (reference main.c:297-319): ``[u: N x 3 interleaved | p | phi | T]``.
    ``wgalpha`` slot 3 is zero; pressure lives in ``dwgalpha`` slot 3 (defect D6).
    """
    rng = np.random.default_rng(seed)
    wgalpha = rng.uniform(-1.0, 1.0, 6 * num_node)
    dwgalpha = rng.uniform(-1.0, 1.0, 6 * num_node)
    wgalpha[3 * num_node:4 * num_node] = 0.0
    return wgalpha, dwgalpha


def state_default(mesh: BoxMesh):
    """State A: the reference's initial condition pushed through one predictor + alpha-level
    build (reference main.c:286-321, 95-118, 544-545): u=(1,0,0), p=0, phi=x, T=-x, dwg=0."""
    n = mesh.num_node
    wgold = np.zeros(6 * n)
    wgold[0:3 * n:3] = 1.0
    wgold[4 * n:5 * n] = mesh.xg[:, 0]
    wgold[5 * n:6 * n] = -mesh.xg[:, 0]
    dwgold = np.zeros(6 * n)
    dwg = np.zeros(6 * n)          # dwg[3N:4N] = buffer[3N:4N] = 0 (pressure)
    # predictor (main.c:544-545) scales zeros -> zeros
    fact1 = (1.0 - K_ALPHAM, K_ALPHAM)
    fact2 = (K_DT * K_ALPHAF * (1.0 - K_GAMMA), K_DT * K_ALPHAF * K_GAMMA)
    dwgalpha = fact1[0] * dwgold + fact1[1] * dwg
    dwgalpha[3 * n:4 * n] = dwg[3 * n:4 * n]
    wgalpha = wgold + fact2[0] * dwgold + fact2[1] * dwg
    wgalpha[3 * n:4 * n] = 0.0
    return wgalpha, dwgalpha


def state_initial(mesh: BoxMesh):
    """(wgold, dwgold, dwg) at step 0: the reference's initial condition (main.c:286-321) with zero rates."""
    n = mesh.num_node
    wgold = np.zeros(6 * n)
    wgold[0:3 * n:3] = 1.0
    wgold[4 * n:5 * n] = mesh.xg[:, 0]
    wgold[5 * n:6 * n] = -mesh.xg[:, 0]
    return wgold, np.zeros(6 * n), np.zeros(6 * n)


def delaunay_cube(n_interior=400, n_face=40, seed=3):
    """UNSTRUCTURED tetrahedral mesh of the unit cube (scipy Delaunay of random points): variable valence (nodal rows of
    8..40 entries, 10..60 tets per node), arbitrary local vertex order -- the kind of input the reference reads from gmsh.
    Same container and the same six boundary groups as boxmesh.make_box."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = [rng.uniform(0.05, 0.95, (n_interior, 3))]
    for axis in range(3):
        for side in (0.0, 1.0):
            p = rng.uniform(0.0, 1.0, (n_face, 3))
            p[:, axis] = side
            pts.append(p)
    g = np.array([0.0, 0.5, 1.0])
    frame = np.array([(x, y, z) for x in g for y in g for z in g if (x in (0, 1)) + (y in (0, 1)) + (z in (0, 1)) >= 2])
    pts.append(frame)                                   # corners and edge mid-points keep the hull a cube
    xg = np.ascontiguousarray(np.concatenate(pts))
    ien = Delaunay(xg).simplices.astype(np.int64)
    d = xg[ien[:, 1:]] - xg[ien[:, :1]]
    det = np.linalg.det(d)
    ien = ien[np.abs(det) > 1e-9]                       # flat tets qhull leaves on the coplanar hull points
    det = det[np.abs(det) > 1e-9]
    neg = det < 0
    ien[neg, 0], ien[neg, 1] = ien[neg, 1].copy(), ien[neg, 0].copy()      # positive orientation (outward Nanson normals)
    assert np.unique(ien).size == xg.shape[0], "orphan node"
    ien = np.ascontiguousarray(ien.astype(np.int32))
    node_off, elem_off, nodes_all, f2e_all, forn_all, bien_all = [0], [0], [], [], [], []
    for axis, side in BOUND_PLANES:
        on = xg[:, axis] == float(side)
        nodes = np.nonzero(on)[0].astype(np.int32)
        on_t = on[ien]
        te = np.nonzero(on_t.sum(axis=1) == 3)[0]
        forn = np.argmin(on_t[te], axis=1)
        nodes_all.append(nodes); f2e_all.append(te.astype(np.int32)); forn_all.append(forn.astype(np.int32))
        bien_all.append(np.stack([ien[te, (forn + 1) % 4], ien[te, (forn + 2) % 4], ien[te, (forn + 3) % 4]], axis=1).astype(np.int32))
        node_off.append(node_off[-1] + len(nodes)); elem_off.append(elem_off[-1] + len(te))
    return BoxMesh(m=0, num_node=xg.shape[0], num_tet=ien.shape[0], xg=xg, ien=ien,
                           bound_node_offset=np.array(node_off, np.int32), bound_node=np.concatenate(nodes_all).astype(np.int32),
                           bound_elem_offset=np.array(elem_off, np.int32), bound_f2e=np.concatenate(f2e_all).astype(np.int32),
                           bound_forn=np.concatenate(forn_all).astype(np.int32), bound_ien=np.concatenate(bien_all).astype(np.int32))
