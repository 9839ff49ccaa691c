/* Minimal stand-in for <hdf5.h> so that the reference's h5util.h (reference src/h5util.h:6) can be
 * #included by Mesh.c / MeshData.c when building oracle/_ref.  libhdf5 is not in this image; no HDF5
 * function is ever called by the hot path.  Test infrastructure only. */
#ifndef ORACLE_REF_HDF5_STUB_H
#define ORACLE_REF_HDF5_STUB_H
typedef long hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
#define H5_VERS_MAJOR 1
#define H5_VERS_MINOR 10
#endif
