/* dedflow_h5flat.h -- DEDFlow's file interface (reference src/h5util.h:17-58: the 20 functions Mesh.c, MeshData.c, Field.c,
 * Array.c, Particle.c and main.c read meshes / write solutions through) over a FLAT container file, for machines without
 * libhdf5 (this image has none).  Same names, same argument meaning, same struct: the reference's host code links against
 * libdedflow_h5flat.so (or h5flat.o) IN PLACE OF h5util.o and runs unchanged -- `box.h5`, `sol.10.h5` are then flat containers.
 * With libhdf5 present keep the reference's h5util.c; dedflow_b200/h5flat.py converts between the two.
 *
 * Container format (little endian):  "DFBH5\0\1\0" | records ...,  record = u32 name_len | name | u8 dtype | u64 count | data
 *   dtype: 0 i32, 1 u32, 2 f32, 3 f64, 4 i64, 5 u64.  Dataset names are HDF5 paths without the leading '/' ("mesh/ien/tet");
 *   a group exists when some dataset name starts with "<group>/".  A later record with the same name replaces an earlier one.
 *   Reads convert between numeric types like H5Dread does (a numpy int64 connectivity reads back as index_type = i32). */
#ifndef DEDFLOW_H5FLAT_H
#define DEDFLOW_H5FLAT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef long dfh_hid_t;                 /* the reference's hid_t slot (src/h5util.h:20; any 8-byte handle) */
typedef struct H5FileInfo H5FileInfo;
struct H5FileInfo {                     /* reference src/h5util.h:17-21, member for member */
  char filename[256];
  dfh_hid_t file_id;
};

H5FileInfo* H5OpenFile(const char* filename, const char* mode);           /* "r", "w" (truncate), "a"   src/h5util.c:7-27 */
void H5CloseFile(H5FileInfo* h5file);                                      /* src/h5util.c:30-33 */
int32_t H5FileExist(const char* filename);                                 /* src/h5util.c:36-43 */
int32_t H5FileIsWritable(H5FileInfo* h5file);                              /* src/h5util.c:45-50 */
int32_t H5FileIsReadable(H5FileInfo* h5file);                              /* src/h5util.c:52-57 (as there: "r" and "a", not "w") */
int32_t H5GroupExist(H5FileInfo* h5file, const char* group_name);          /* src/h5util.c:60-69 */
int32_t H5DatasetExist(H5FileInfo* h5file, const char* dataset_name);      /* src/h5util.c:71-80 */
void H5GetDatasetSize(H5FileInfo* h5file, const char* dataset_name, int32_t* size);   /* 0 when absent, src/h5util.c:82-102 */
void H5ReadDataseti32(H5FileInfo* h5file, const char* dataset_name, int32_t* data);
void H5ReadDatasetu32(H5FileInfo* h5file, const char* dataset_name, uint32_t* data);
void H5ReadDatasetf32(H5FileInfo* h5file, const char* dataset_name, float* data);
void H5ReadDatasetf64(H5FileInfo* h5file, const char* dataset_name, double* data);
void H5ReadDatasetInd(H5FileInfo* h5file, const char* dataset_name, int32_t* data);   /* index_type = i32 (USE_I32_INDEX) */
void H5ReadDatasetVal(H5FileInfo* h5file, const char* dataset_name, double* data);    /* value_type = f64 (USE_F64_VALUE) */
void H5WriteDataseti32(H5FileInfo* h5file, const char* dataset_name, int32_t len, const int32_t* data);
void H5WriteDatasetu32(H5FileInfo* h5file, const char* dataset_name, int32_t len, const uint32_t* data);
void H5WriteDatasetf32(H5FileInfo* h5file, const char* dataset_name, int32_t len, const float* data);
void H5WriteDatasetf64(H5FileInfo* h5file, const char* dataset_name, int32_t len, const double* data);
void H5WriteDatasetInd(H5FileInfo* h5file, const char* dataset_name, int32_t len, const int32_t* data);
void H5WriteDatasetVal(H5FileInfo* h5file, const char* dataset_name, int32_t len, const double* data);

#ifdef __cplusplus
}
#endif
#endif
