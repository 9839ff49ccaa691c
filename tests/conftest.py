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
    """see dedflow_b200.boxmesh.delaunay_cube (the generator lives in the package: bench.py's embedded parity check uses it too)"""
    from dedflow_b200 import boxmesh
    return boxmesh.delaunay_cube(n_interior, n_face, seed)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    return pyoracle.get()
