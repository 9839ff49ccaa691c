import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


EVEN_PERMS = np.array([[0, 1, 2, 3], [0, 2, 3, 1], [0, 3, 1, 2], [1, 0, 3, 2], [1, 2, 0, 3], [1, 3, 2, 0],
                       [2, 0, 1, 3], [2, 1, 3, 0], [2, 3, 0, 1], [3, 0, 2, 1], [3, 1, 0, 2], [3, 2, 1, 0]])


def shuffled_mesh(m, seed=7):
    """Kuhn box whose local vertex order is permuted per element by a random EVEN permutation (orientation kept):
    exercises every face orientation (forn 0..3) and every local-index dependent code path."""
    from dedflow_b200 import boxmesh
    mesh = boxmesh.make_box(m)
    rng = np.random.default_rng(seed)
    perm = EVEN_PERMS[rng.integers(0, 12, mesh.num_tet)]
    mesh.ien = np.ascontiguousarray(np.take_along_axis(mesh.ien, perm, axis=1).astype(np.int32))
    # recompute forn for the permuted connectivity
    forn = mesh.bound_forn.copy()
    for b in range(mesh.num_bound):
        s, e = mesh.bound_elem_offset[b], mesh.bound_elem_offset[b + 1]
        on = np.isin(mesh.ien[mesh.bound_f2e[s:e]], mesh.bound_nodes(b))
        forn[s:e] = np.argmin(on, axis=1)
    mesh.bound_forn = forn.astype(np.int32)
    return mesh


def delaunay_mesh(n_interior=400, n_face=40, seed=3):
    """UNSTRUCTURED tetrahedral mesh of the unit cube (scipy Delaunay of random points): variable valence (nodal rows of
    8..40 entries, 10..60 tets per node), arbitrary local vertex order -- the kind of input the reference reads from gmsh.
    Same container and the same six boundary groups as boxmesh.make_box."""
    from scipy.spatial import Delaunay
    from dedflow_b200 import boxmesh
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
    for axis, side in boxmesh.BOUND_PLANES:
        on = xg[:, axis] == float(side)
        nodes = np.nonzero(on)[0].astype(np.int32)
        on_t = on[ien]
        te = np.nonzero(on_t.sum(axis=1) == 3)[0]
        forn = np.argmin(on_t[te], axis=1)
        nodes_all.append(nodes); f2e_all.append(te.astype(np.int32)); forn_all.append(forn.astype(np.int32))
        bien_all.append(np.stack([ien[te, (forn + 1) % 4], ien[te, (forn + 2) % 4], ien[te, (forn + 3) % 4]], axis=1).astype(np.int32))
        node_off.append(node_off[-1] + len(nodes)); elem_off.append(elem_off[-1] + len(te))
    return boxmesh.BoxMesh(m=0, num_node=xg.shape[0], num_tet=ien.shape[0], xg=xg, ien=ien,
                           bound_node_offset=np.array(node_off, np.int32), bound_node=np.concatenate(nodes_all).astype(np.int32),
                           bound_elem_offset=np.array(elem_off, np.int32), bound_f2e=np.concatenate(f2e_all).astype(np.int32),
                           bound_forn=np.concatenate(forn_all).astype(np.int32), bound_ien=np.concatenate(bien_all).astype(np.int32))


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    return pyoracle.get()
