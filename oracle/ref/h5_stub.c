/* Trapping stand-ins for the reference's HDF5 helper (reference src/h5util.h:24-58).  oracle/_ref builds the
 * reference's hot-path sources unmodified; Mesh.c / MeshData.c reference these symbols from their file-reading
 * constructors, which the parity harness never calls (it fills Mesh3D from the synthetic box mesh instead).
 * Test infrastructure only. */
#include <stdio.h>
#include <stdlib.h>
#include "h5util.h"

static void trap(const char* name) {
	fprintf(stderr, "oracle/_ref: %s called but HDF5 is not available in this image\n", name);
	abort();
}
H5FileInfo* H5OpenFile(const char* filename, const char* mode) { (void)filename; (void)mode; trap("H5OpenFile"); return NULL; }
void H5CloseFile(H5FileInfo* f) { (void)f; trap("H5CloseFile"); }
b32 H5FileExist(const char* filename) { (void)filename; return FALSE; }
b32 H5FileIsWritable(H5FileInfo* f) { (void)f; return FALSE; }
b32 H5FileIsReadable(H5FileInfo* f) { (void)f; return FALSE; }
b32 H5GroupExist(H5FileInfo* f, const char* n) { (void)f; (void)n; return FALSE; }
b32 H5DatasetExist(H5FileInfo* f, const char* n) { (void)f; (void)n; return FALSE; }
void H5GetDatasetSize(H5FileInfo* f, const char* n, index_type* size) { (void)f; (void)n; (void)size; trap("H5GetDatasetSize"); }
#define RD(name, T) void name(H5FileInfo* f, const char* n, T* d) { (void)f; (void)n; (void)d; trap(#name); }
RD(H5ReadDataseti32, i32) RD(H5ReadDatasetu32, u32) RD(H5ReadDatasetf32, f32) RD(H5ReadDatasetf64, f64)
RD(H5ReadDatasetInd, index_type) RD(H5ReadDatasetVal, value_type)
#define WR(name, T) void name(H5FileInfo* f, const char* n, index_type len, const T* d) { (void)f; (void)n; (void)len; (void)d; trap(#name); }
WR(H5WriteDataseti32, i32) WR(H5WriteDatasetu32, u32) WR(H5WriteDatasetf32, f32) WR(H5WriteDatasetf64, f64)
WR(H5WriteDatasetInd, index_type) WR(H5WriteDatasetVal, value_type)
