"""The CPU oracle against golden vectors produced by the reference's OWN CUDA build (libdedflow_ref.so, built from
/root/reference/src by oracle/ref/Makefile) on a B200 -- see tests/golden/README.md.  Runs without a GPU."""
from pathlib import Path

import numpy as np
import pytest

from conftest import shuffled_mesh
from dedflow_b200 import boxmesh
from oracle import pyoracle

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
GOLDEN = GOLDEN_DIR / "ref_m6_shuffled_stateB.npz"
GOLDENS = ["ref_m6_shuffled_stateB.npz", "ref_delaunay_stateB.npz"]


def load_golden(name=None):
    return dict(np.load(GOLDEN_DIR / name if name else GOLDEN))


def golden_mesh(g):
    assert str(g["state"]) == "B"
    if "mesh_xg" in g:      # unstructured golden: the mesh travels inside the file (qhull output is not ours to reproduce)
        return boxmesh.BoxMesh(m=0, num_node=g["mesh_xg"].shape[0], num_tet=g["mesh_ien"].shape[0], xg=g["mesh_xg"],
                               ien=g["mesh_ien"], bound_node_offset=g["mesh_bound_node_offset"], bound_node=g["mesh_bound_node"],
                               bound_elem_offset=g["mesh_bound_elem_offset"], bound_f2e=g["mesh_bound_f2e"],
                               bound_forn=g["mesh_bound_forn"], bound_ien=g["mesh_bound_ien"])
    assert bool(g["shuffle"])
    return shuffled_mesh(int(g["m"]))


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("golden", GOLDENS)
def test_oracle_integer_outputs_bit_exact_vs_reference(oracle, golden):
    g = load_golden(golden)
    mesh = golden_mesh(g)
    N = mesh.num_node
    rp, ci = oracle.nodal_pattern(N, mesh.ien)
    assert np.array_equal(rp, g["row_ptr"]) and np.array_equal(ci, g["col_ind"])
    for name, (br, bc) in {"3x3": (3, 3), "3x1": (3, 1), "1x3": (1, 3)}.items():
        nrp, nci = oracle.expand_block(rp, ci, br, bc, fix_last=bool(g["d1_patched"]))
        assert np.array_equal(nrp, g[f"row_ptr_{name}"]) and np.array_equal(nci, g[f"col_ind_{name}"])
    w = oracle.weights(pyoracle.curand_host_u32(mesh.num_tet))
    color, rounds, ties = oracle.color_jpl(N, mesh.ien, w)
    assert ties == 0
    assert np.array_equal(color, g["color"])
    off, ind = oracle.color_batches(color)
    assert np.array_equal(off, g["batch_offset"]) and np.array_equal(ind, g["batch_ind"])


@pytest.mark.parametrize("golden", GOLDENS)
def test_oracle_assembly_and_solve_vs_reference(oracle, golden):
    g = load_golden(golden)
    mesh = golden_mesh(g)
    N = mesh.num_node
    wg, dwg = boxmesh.state_random(N)
    pat = (g["row_ptr"], g["col_ind"])
    Z = g["col_ind"].size
    F = np.zeros(6 * N)
    blocks = [np.zeros(9 * Z), np.zeros(3 * Z), np.zeros(3 * Z), np.zeros(Z)]
    nc = int(g["color"].max()) + 1
    oracle.assemble_tet(N, mesh.ien, mesh.xg, g["batch_offset"], g["batch_ind"], wg, dwg, F=F)
    oracle.assemble_tet(N, mesh.ien, mesh.xg, g["batch_offset"], g["batch_ind"], wg, dwg, pattern=pat, blocks=blocks)
    f2e, forn = mesh.bound_faces(4)
    oracle.assemble_face(f2e, forn, N, mesh.ien, mesh.xg, g["color"], nc, wg, dwg, F=F)
    oracle.assemble_face(f2e, forn, N, mesh.ien, mesh.xg, g["color"], nc, wg, dwg, pattern=pat, blocks=blocks)
    F[4 * N:] = 0
    for b, t in {0: (1, 1, 1), 2: (0, 1, 0), 3: (0, 0, 1), 4: (0, 0, 0)}.items():
        oracle.dirichlet_vec(mesh.bound_nodes(b), np.array(t, np.int32), F)
        oracle.dirichlet_mat(mesh.bound_nodes(b), np.array(t, np.int32), N, pat, blocks[0], blocks[1])
    assert rel(F, g["F"]) <= 1e-12
    for a, nme in zip(blocks, ("A00", "A01", "A10", "A11")):
        assert rel(a, g[nme]) <= 1e-12, nme
    gb = [g["A00"], g["A01"], g["A10"], g["A11"]]
    y = np.zeros(6 * N)
    oracle.fs_amvpby(pat, gb, 1.0, g["x"], 0.0, y)
    assert rel(y[:4 * N], g["y"][:4 * N]) <= 1e-13 and np.all(g["y"][4 * N:] == 0)
    x, it, hist = oracle.gmres(pat, gb, g["F"])
    assert it == int(g["printed"][-1, 0])
    assert rel(x[:4 * N], g["dx"][:4 * N]) <= 1e-10
    for k, v in g["printed"]:
        assert abs(hist[int(k)] - v) <= 6e-5 * v
    x120, _, h120 = oracle.gmres(pat, gb, g["F"], atol=0.0, rtol=0.0)
    for k, r in zip(g["res_iters"], g["res_true"]):
        assert abs(h120[int(k)] - r) <= 1e-10 * h120[0]
    assert rel(x120[:4 * N], g["dx120"][:4 * N]) <= 1e-10
