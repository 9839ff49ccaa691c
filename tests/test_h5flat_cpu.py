"""CPU tests of the flat-file stand-in for DEDFlow's HDF5 interface (include/dedflow_h5flat.h, dedflow_b200/csrc/h5flat.c,
dedflow_b200/h5flat.py): the 20 functions of reference src/h5util.h:24-58 through ctypes, against the numpy reader/writer."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from dedflow_b200 import boxmesh, h5flat

ROOT = Path(__file__).resolve().parents[1]
SO = ROOT / "dedflow_b200" / "libdedflow_h5flat.so"


class H5FileInfo(C.Structure):
    _fields_ = [("filename", C.c_char * 256), ("file_id", C.c_long)]


@pytest.fixture(scope="module")
def L():
    if not SO.exists():
        import __graft_entry__
        __graft_entry__.build_h5flat()
    lib = C.CDLL(str(SO))
    lib.H5OpenFile.restype = C.POINTER(H5FileInfo)
    lib.H5OpenFile.argtypes = [C.c_char_p, C.c_char_p]
    lib.H5CloseFile.argtypes = [C.POINTER(H5FileInfo)]
    for f in ("H5FileIsWritable", "H5FileIsReadable"):
        getattr(lib, f).argtypes = [C.POINTER(H5FileInfo)]
    for f in ("H5GroupExist", "H5DatasetExist"):
        getattr(lib, f).argtypes = [C.POINTER(H5FileInfo), C.c_char_p]
    lib.H5GetDatasetSize.argtypes = [C.POINTER(H5FileInfo), C.c_char_p, C.POINTER(C.c_int32)]
    for f in ("i32", "u32", "f32", "f64", "Ind", "Val"):
        getattr(lib, "H5ReadDataset" + f).argtypes = [C.POINTER(H5FileInfo), C.c_char_p, C.c_void_p]
        getattr(lib, "H5WriteDataset" + f).argtypes = [C.POINTER(H5FileInfo), C.c_char_p, C.c_int32, C.c_void_p]
    return lib


def test_exports_the_reference_interface(L):
    for name in ("H5OpenFile H5CloseFile H5FileExist H5FileIsWritable H5FileIsReadable H5GroupExist H5DatasetExist H5GetDatasetSize "
                 "H5ReadDataseti32 H5ReadDatasetu32 H5ReadDatasetf32 H5ReadDatasetf64 H5ReadDatasetInd H5ReadDatasetVal "
                 "H5WriteDataseti32 H5WriteDatasetu32 H5WriteDatasetf32 H5WriteDatasetf64 H5WriteDatasetInd H5WriteDatasetVal").split():
        assert hasattr(L, name), name


def test_mesh_file_round_trip_like_Mesh3DCreateH5(L, tmp_path):
    """a mesh written by the numpy writer (int64 connectivity, like numpy / h5py produce) read the way Mesh.c / MeshData.c read
    it: sizes first (absent prism / hex datasets report 0), then typed reads with conversion to index_type = i32"""
    mesh = boxmesh.make_box(3)
    p = str(tmp_path / "box.h5").encode()
    assert L.H5FileExist(p) == 0
    h5flat.write_mesh(p.decode(), mesh)
    assert L.H5FileExist(p) == 1
    f = L.H5OpenFile(p, b"r")
    assert L.H5FileIsReadable(f) == 1 and L.H5FileIsWritable(f) == 0
    assert L.H5GroupExist(f, b"mesh") == 1 and L.H5GroupExist(f, b"mesh/bound") == 1 and L.H5GroupExist(f, b"fields") == 0
    assert L.H5DatasetExist(f, b"mesh/xg") == 1 and L.H5DatasetExist(f, b"mesh") == 0
    n = C.c_int32(-1)
    sizes = {}
    for name in ("mesh/xg", "mesh/ien/tet", "mesh/ien/prism", "mesh/ien/hex", "mesh/bound/node_offset", "/mesh/bound/f2e"):
        L.H5GetDatasetSize(f, name.encode(), C.byref(n))
        sizes[name] = n.value
    assert sizes == {"mesh/xg": 3 * mesh.num_node, "mesh/ien/tet": 4 * mesh.num_tet, "mesh/ien/prism": 0, "mesh/ien/hex": 0,
                     "mesh/bound/node_offset": mesh.num_bound + 1, "/mesh/bound/f2e": mesh.bound_f2e.size}
    xg = np.zeros(3 * mesh.num_node)
    L.H5ReadDatasetf64(f, b"mesh/xg", xg.ctypes.data)
    assert np.array_equal(xg, mesh.xg.reshape(-1))
    ien = np.zeros(4 * mesh.num_tet, np.int32)
    L.H5ReadDatasetInd(f, b"mesh/ien/tet", ien.ctypes.data)           # stored i64 -> i32
    assert np.array_equal(ien, mesh.ien.reshape(-1))
    forn = np.zeros(mesh.bound_forn.size, np.int32)
    L.H5ReadDatasetInd(f, b"mesh/bound/forn", forn.ctypes.data)
    assert np.array_equal(forn, mesh.bound_forn)
    x32 = np.zeros(3 * mesh.num_node, np.float32)
    L.H5ReadDatasetf32(f, b"mesh/xg", x32.ctypes.data)                # f64 -> f32
    assert np.array_equal(x32, mesh.xg.reshape(-1).astype(np.float32))
    L.H5CloseFile(f)
    d = h5flat.read(p.decode())
    assert d["mesh/ien/tet"].dtype == np.int64 and np.array_equal(d["mesh/bound/node"], mesh.bound_node)
    tri = d["mesh/bound/ien"].reshape(-1, 3)                              # boundary triangles lie on their element, opposite forn
    for t, e, o in zip(tri, d["mesh/bound/f2e"], d["mesh/bound/forn"]):
        assert set(t) == set(np.delete(mesh.ien[e], o))


def test_solution_file_write_append_replace(L, tmp_path):
    """what main.c:521-532,571-591 does: open "w", write u / p / phi / T, close; reopen "r"; "a" appends and replaces"""
    p = str(tmp_path / "sol.10.h5").encode()
    rng = np.random.default_rng(0)
    u, ph = rng.standard_normal(30), rng.standard_normal(10)
    f = L.H5OpenFile(p, b"w")
    assert L.H5FileIsWritable(f) == 1 and L.H5FileIsReadable(f) == 0
    L.H5WriteDatasetf64(f, b"u", 30, u.ctypes.data)
    L.H5WriteDatasetVal(f, b"phi", 10, ph.ctypes.data)
    ids = np.arange(7, dtype=np.int32)
    L.H5WriteDatasetInd(f, b"ptc/ids", 7, ids.ctypes.data)
    L.H5WriteDatasetu32(f, b"ptc/flags", 0, None)                          # empty dataset
    assert L.H5DatasetExist(f, b"u") == 1                                  # visible before the file is closed
    L.H5CloseFile(f)
    d = h5flat.read(p.decode())
    assert np.array_equal(d["u"], u) and np.array_equal(d["phi"], ph) and np.array_equal(d["ptc/ids"], ids) and d["ptc/flags"].size == 0
    f = L.H5OpenFile(p, b"a")
    assert L.H5FileIsWritable(f) == 1 and L.H5FileIsReadable(f) == 1
    u2 = u[:12] * 2.0
    L.H5WriteDatasetf64(f, b"u", 12, u2.ctypes.data)                       # a later record replaces the earlier one
    n = C.c_int32(0)
    L.H5GetDatasetSize(f, b"u", C.byref(n))
    assert n.value == 12
    back = np.zeros(12)
    L.H5ReadDatasetVal(f, b"u", back.ctypes.data)
    assert np.array_equal(back, u2)
    L.H5CloseFile(f)
    assert np.array_equal(h5flat.read(p.decode())["u"], u2)
    h5flat.write(p.decode(), {"T": np.ones(4, np.float32)}, mode="a")
    f = L.H5OpenFile(p, b"r")
    t = np.zeros(4)
    L.H5ReadDatasetf64(f, b"T", t.ctypes.data)                             # f32 -> f64
    assert np.array_equal(t, np.ones(4))
    L.H5CloseFile(f)
