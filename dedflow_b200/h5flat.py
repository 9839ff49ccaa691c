"""Flat stand-in for DEDFlow's HDF5 files (include/dedflow_h5flat.h): numpy reader / writer of the container format, the
mesh writer in the schema of the reference's tools/mesh_convert.py:116-126 (what Mesh3DCreateH5 reads, src/Mesh.c:12-104,
src/MeshData.c:57-109) and the solution-file layout of src/main.c:521-532,571-591.

    write(path, {"mesh/xg": xg, ...})      read(path) -> {name: array}
    write_mesh(path, mesh)                 mesh: dedflow_b200.boxmesh.BoxMesh (or anything with the same fields)
    from_hdf5(src, dst) / to_hdf5(src, dst) conversions, when h5py is installed (it is not in this image)
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

MAGIC = b"DFBH5\x00\x01\x00"
_DTYPES = [np.dtype("<i4"), np.dtype("<u4"), np.dtype("<f4"), np.dtype("<f8"), np.dtype("<i8"), np.dtype("<u8")]


def _code(dt: np.dtype) -> int:
    for i, d in enumerate(_DTYPES):
        if dt == d:
            return i
    raise TypeError(f"h5flat: unsupported dtype {dt}")


def write(path, datasets: dict, mode: str = "w") -> None:
    """mode "w": new container; "a": append records (a later record replaces an earlier one of the same name)."""
    with open(path, "wb" if mode == "w" else "ab") as f:
        if mode == "w":
            f.write(MAGIC)
        for name, arr in datasets.items():
            a = np.ascontiguousarray(arr).reshape(-1)          # every DEDFlow dataset is flattened to 1-D
            if a.dtype == np.bool_:
                a = a.astype(np.int32)
            a = a.astype(a.dtype.newbyteorder("<"), copy=False)
            nm = name.lstrip("/").encode()
            f.write(struct.pack("<I", len(nm)) + nm + struct.pack("<BQ", _code(a.dtype), a.size))
            f.write(a.tobytes())


def read(path) -> dict:
    out = {}
    data = Path(path).read_bytes()
    if data[:8] != MAGIC:
        raise ValueError(f"{path}: not a dedflow flat container")
    pos = 8
    while pos < len(data):
        (nl,) = struct.unpack_from("<I", data, pos)
        pos += 4
        name = data[pos:pos + nl].decode()
        pos += nl
        code, count = struct.unpack_from("<BQ", data, pos)
        pos += 9
        dt = _DTYPES[code]
        out[name] = np.frombuffer(data, dtype=dt, count=count, offset=pos).copy()
        pos += count * dt.itemsize
    return out


def mesh_datasets(mesh, group: str = "mesh") -> dict:
    """the datasets Mesh3DCreateH5 reads, schema of tools/mesh_convert.py:116-126"""
    nb = mesh.num_bound
    f2e = np.asarray(mesh.bound_f2e, np.int64)
    forn = np.asarray(mesh.bound_forn, np.int64)
    # boundary triangles: the three vertices of the element that are not the opposite vertex forn (mesh_convert.py:80-98)
    ien = np.asarray(mesh.ien, np.int64)
    if getattr(mesh, "bound_ien", None) is not None:
        tri = np.asarray(mesh.bound_ien, np.int64).reshape(-1, 3)
    else:
        tri = np.stack([np.delete(ien[e], o) for e, o in zip(f2e, forn)]) if f2e.size else np.zeros((0, 3), np.int64)
    d = {f"{group}/xg": np.asarray(mesh.xg, np.float64).reshape(-1),
         f"{group}/ien/tet": ien.reshape(-1),
         f"{group}/bound/node_offset": np.asarray(mesh.bound_node_offset, np.int64),
         f"{group}/bound/node": np.asarray(mesh.bound_node, np.int64),
         f"{group}/bound/elem_offset": np.asarray(mesh.bound_elem_offset, np.int64),
         f"{group}/bound/ien": tri.reshape(-1),
         f"{group}/bound/f2e": f2e,
         f"{group}/bound/forn": forn}
    assert len(d[f"{group}/bound/node_offset"]) == nb + 1
    return d


def write_mesh(path, mesh, group: str = "mesh") -> None:
    write(path, mesh_datasets(mesh, group))


def from_hdf5(src, dst) -> None:
    import h5py                                   # not in this image; the conversion is for machines that have it
    out = {}
    with h5py.File(src, "r") as f:
        f.visititems(lambda name, obj: out.__setitem__(name, obj[()]) if isinstance(obj, h5py.Dataset) else None)
    write(dst, out)


def to_hdf5(src, dst) -> None:
    import h5py
    with h5py.File(dst, "w") as f:
        for name, arr in read(src).items():
            f.create_dataset(name, data=arr)
