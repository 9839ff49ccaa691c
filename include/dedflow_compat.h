/*
 * dedflow_compat.h -- the DROP-IN layer of libdedflow_b200.so: DEDFlow's own entry points for the FEM linear-system
 * hot path, with DEDFlow's struct layouts, implemented on the B200 kernels of dedflow_b200.h.
 *
 * A DEDFlow build keeps its host code (src/main.c, Mesh.c, MeshData.c, Field.c, h5util.c, common.c, alloc.c) and its
 * own headers, drops assemble.cu color*.{c,cu} csr*.{c,cu} indexing.cu matrix*.{c,cu} dirichlet*.{c,cu} krylov.c
 * krylov_util.cu pc*.{c,cu} vec.cu from the link line and links -ldedflow_b200 instead.  This header
 * is for callers that do NOT have the DEDFlow headers: it declares the same structs (same member order, types and
 * therefore offsets -- the driver pokes into them, reference src/main.c:382-403,460-476) and the same functions.
 * Do not include it together with the DEDFlow headers.
 *
 * Build configuration mirrored: index_type = int32 (USE_I32_INDEX), value_type = double (USE_F64_VALUE),
 * color_t = int32 (reference config/config.mk:51, src/common.h:39-59, src/color.h:12).
 *
 * Error convention (reference src/common.h:90-98, src/matrix.c:730-738): entry points return void or a pointer; a CUDA
 * failure prints "GPUAssert: <message> <file> <line>" to stdout; an operation a matrix type does not implement prints
 * "Matrix operation <op> is not implemented for type: <t>" to stderr and returns.  In addition dfb_last_error()
 * (dedflow_b200.h) keeps the last message.
 *
 * Each declaration cites the reference declaration it replaces.
 */
#ifndef DEDFLOW_COMPAT_H
#define DEDFLOW_COMPAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

typedef int32_t dfc_index;  /* index_type */
typedef double dfc_value;   /* value_type */
typedef int32_t dfc_color;  /* color_t    */
typedef int32_t dfc_b32;    /* b32        */

/* ---- mesh containers: OWNED BY THE HOST CODE, only read here (src/MeshData.h:10-19, src/Mesh.h:14-45) ---- */
typedef struct Mesh3DData {
  dfc_b32 is_host;
  dfc_index num_node, num_tet, num_prism, num_hex;
  double* xg;      /* [3*num_node] interleaved */
  dfc_index* ien;  /* [4*num_tet + ...] */
} Mesh3DData;

typedef struct Mesh3D {
  dfc_index num_node, num_tet, num_prism, num_hex;
  Mesh3DData* host;
  Mesh3DData* device;
  dfc_index num_bound;
  dfc_index* bound_fid;
  dfc_index* bound_node_offset; /* host  [num_bound+1] */
  dfc_index* bound_node;        /* device */
  dfc_index* bound_elem_offset; /* host  [num_bound+1] */
  dfc_index* bound_ien;
  dfc_index* bound_f2e;         /* device: face -> element */
  dfc_index* bound_forn;        /* device: local index of the vertex opposite the face */
  dfc_index num_batch;
  dfc_index* batch_offset;      /* host  [num_batch+1] */
  dfc_index* batch_ind;         /* device [num_tet] */
  dfc_color num_color;
  dfc_color* color;             /* device [num_tet] */
} Mesh3D;

/* ---- sparsity pattern (src/csr.h:12-31) ---- */
typedef struct CSRAttr CSRAttr;
struct CSRAttr {
  dfc_index num_row, num_col, nnz;
  dfc_index* row_ptr; /* device */
  dfc_index* col_ind; /* device */
  const CSRAttr* parent;
};
CSRAttr* CSRAttrCreate(const Mesh3D* mesh);                                           /* csr.h:28, csr.c:143-190 */
CSRAttr* CSRAttrCreateBlock(const CSRAttr* attr, dfc_index block_row, dfc_index block_col); /* csr.h:31, csr.c:193-218 */
void CSRAttrDestroy(CSRAttr* attr);                                                   /* csr.h:29 */

/* ---- coloring and batches (src/color.h:13-17, src/indexing.h:9-14; called by Mesh.c:165-206) ---- */
void ColorMeshTet(const Mesh3D* mesh, dfc_index max_color_len, dfc_color* color);
dfc_color GetMaxColor(const dfc_color* color, dfc_index num_elem);
dfc_index CountValueI(const dfc_index* data, dfc_index n, dfc_index value);
void FindValueI(const dfc_index* data, dfc_index n, dfc_index value, dfc_index* result);
dfc_index CountValueColorLegacy(const dfc_color* data, dfc_index n, dfc_color value);
dfc_index CountValueColor(const dfc_color* data, dfc_index n, dfc_color value, void* buffer);
void FindValueColor(const dfc_color* data, dfc_index n, dfc_color value, dfc_index* result);

/* ---- matrix objects (src/matrix.h:13-137) ---- */
typedef enum MatType { MAT_TYPE_NONE = 0, MAT_TYPE_DENSE = 1, MAT_TYPE_CSR = 2, MAT_TYPE_FS = 4, MAT_TYPE_CUSTOM = 8 } MatType;
typedef struct Matrix Matrix;
typedef struct MatrixOp {  /* 15 slots, same order as src/matrix.h:27-61 */
  void (*setup)(Matrix*);
  void (*zero)(Matrix*);
  void (*zero_row)(Matrix*, dfc_index, const dfc_index* row, dfc_index shift, dfc_value diag);
  void (*amvpby)(Matrix*, dfc_value alpha, dfc_value* x, dfc_value beta, dfc_value* y);
  void (*amvpby_mask)(Matrix*, dfc_value, dfc_value*, dfc_value, dfc_value*, dfc_value*, dfc_value*);
  void (*matvec)(Matrix*, dfc_value* x, dfc_value* y);
  void (*matvec_mask)(Matrix*, dfc_value*, dfc_value*, dfc_value*, dfc_value*);
  void (*get_diag)(Matrix*, dfc_value* diag, dfc_index bs);
  void (*set_values_coo)(Matrix*, dfc_value, dfc_index, const dfc_index*, const dfc_index*, const dfc_value*, dfc_value);
  void (*set_values_ind)(Matrix*, dfc_value, dfc_index, const dfc_index*, const dfc_value*, dfc_value);
  void (*add_elem_value_batched)(Matrix*, dfc_index, dfc_index, const dfc_index*, const dfc_index*, const dfc_value*,
                                 const dfc_index*);
  void (*add_elem_value_blocked_batched)(Matrix*, dfc_index nshl, dfc_index batch_size, const dfc_index* batch_ptr,
                                         const dfc_index* ien, dfc_index block_row, dfc_index block_col,
                                         const dfc_value* val, int lda, int stride, const dfc_index* mask);
  void (*add_value_batched)(Matrix*, dfc_index, const dfc_index*, const dfc_index*, const dfc_value*);
  void (*add_value_blocked_batched)(Matrix*, dfc_index, const dfc_index*, const dfc_index*, dfc_index, dfc_index,
                                    const dfc_value*, int, int);
  void (*destroy)(Matrix*);
} MatrixOp;

struct Matrix {
  dfc_index size[2];
  MatType type;
  void* data;       /* MatrixCSR* or MatrixFS* */
  void* stream_ref; /* cudaStream_t */
  MatrixOp op[1];
};

typedef struct MatrixCSR {
  dfc_b32 external_attr;
  const CSRAttr* attr; /* borrowed */
  dfc_value* val;      /* device [attr->nnz], owned */
  void* descr;         /* cusparseSpMatDescr_t in the reference; always NULL here (no cuSPARSE) */
  dfc_index buffer_size;
  void* buffer;
} MatrixCSR;

typedef struct MatrixFS {
  dfc_index n_offset;
  dfc_index* offset;   /* host   [n_offset+1] */
  dfc_index* d_offset; /* device [n_offset+1] */
  void** stream;       /* cudaStream_t[n_offset] in the reference; entries are NULL here (single stream) */
  const CSRAttr* spy1x1; /* assigned by the caller (main.c:383) */
  dfc_value** d_matval;  /* device table of the sub-block value pointers */
  Matrix** mat;          /* host [n_offset*n_offset], assigned by the caller (main.c:385-391) */
} MatrixFS;

Matrix* MatrixCreateTypeCSR(const CSRAttr* attr, void* ctx);                          /* matrix.h:107 */
Matrix* MatrixCreateTypeFS(dfc_index n_offset, const dfc_index* offset, void* ctx);   /* matrix.h:108 */
void MatrixDestroy(Matrix* matrix);
void MatrixSetup(Matrix* matrix);
void MatrixZero(Matrix* matrix);
void MatrixZeroRow(Matrix* matrix, dfc_index n, const dfc_index* row, dfc_index shift, dfc_value diag);
void MatrixAMVPBY(Matrix* A, dfc_value alpha, dfc_value* x, dfc_value beta, dfc_value* y);
void MatrixAMVPBYWithMask(Matrix* A, dfc_value alpha, dfc_value* x, dfc_value beta, dfc_value* y, dfc_value* left_mask,
                          dfc_value* right_mask);
void MatrixMatVec(Matrix* matrix, dfc_value* x, dfc_value* y);
void MatrixMatVecWithMask(Matrix* matrix, dfc_value* x, dfc_value* y, dfc_value* left_mask, dfc_value* right_mask);
void MatrixGetDiag(Matrix* matrix, dfc_value* diag, dfc_index bs);
void MatrixSetValuesCOO(Matrix* matrix, dfc_value alpha, dfc_index n, const dfc_index* row, const dfc_index* col,
                        const dfc_value* val, dfc_value beta);
void MatrixSetValuesInd(Matrix* matrix, dfc_value alpha, dfc_index n, const dfc_index* ind, const dfc_value* val,
                        dfc_value beta);
void MatrixAddElemValueBatched(Matrix* matrix, dfc_index nshl, dfc_index num_batch, const dfc_index* batch_ptr,
                               const dfc_index* ien, const dfc_value* val, const dfc_index* mask);
void MatrixAddElemValueBlockedBatched(Matrix* matrix, dfc_index nshl, dfc_index num_batch, const dfc_index* batch_ptr,
                                      const dfc_index* ien, dfc_index block_row_size, dfc_index block_col_size,
                                      const dfc_value* val, int lda, int stride, const dfc_index* mask);
void MatrixAddValueBatched(Matrix* matrix, dfc_index batch_size, const dfc_index* batch_row_ind,
                           const dfc_index* batch_col_ind, const dfc_value* A);
void MatrixAddValueBlockedBatched(Matrix* matrix, dfc_index batch_size, const dfc_index* batch_row_ind,
                                  const dfc_index* batch_col_ind, dfc_index block_row_size, dfc_index block_col_size,
                                  const dfc_value* A, int lda, int stride);
MatrixCSR* MatrixCSRCreate(const CSRAttr* attr, void* ctx);                           /* matrix.h:141 */
void MatrixCSRDestroy(Matrix* matrix);
MatrixFS* MatrixFSCreate(dfc_index n_offset, const dfc_index* offset, void* ctx);     /* matrix.h:146 */
void MatrixFSDestroy(Matrix* matrix);

/* ---- assembly (src/assemble.h:13-14) ---- */
void AssembleSystemTet(Mesh3D* mesh, double* wgalpha, double* dwgalpha, double* F, Matrix* J);
void AssembleSystemTetFace(Mesh3D* mesh, double* wgalpha, double* dwgalpha, double* F, Matrix* J);

/* ---- Dirichlet (src/dirichlet.h:8-33) ---- */
typedef enum BCType { BC_NONE = 0, BC_STRONG = 1, BC_WEAK = 2, BC_OUTFLOW = 4 } BCType;
typedef struct Dirichlet {
  const Mesh3D* mesh;
  dfc_index face_ind;
  dfc_index shape;
  size_t buffer_size;
  void* buffer;     /* device copy of the face's node list */
  BCType bctype[];  /* [shape], written by the caller (main.c:460-476) */
} Dirichlet;
Dirichlet* DirichletCreate(const Mesh3D* mesh, dfc_index face_ind, dfc_index shape);
void DirichletDestroy(Dirichlet* dirichlet);
void DirichletApplyVec(Dirichlet* dirichlet, dfc_value* b);
void DirichletApplyMat(Dirichlet* dirichlet, Matrix* A);

/* ---- preconditioners (src/pc.h:14-88) ---- */
typedef enum PCType { PC_NONE = 0x0, PC_JACOBI = 0x1, PC_DECOMPOSITION = 0x2, PC_AMGX = 0x3, PC_CUSTOM = 0x4 } PCType;
typedef struct PC PC;
typedef struct PCOps {
  void (*setup)(PC*);
  void (*destroy)(PC*);
  void (*apply)(PC*, dfc_value* x, dfc_value* y);
} PCOps;
struct PC {
  PCType type;
  void* mat;
  PCOps op[1];
  void* data;
  void* cublas_handle; /* kept for layout; unused (no cuBLAS) */
};
typedef struct PCNone { dfc_index n; } PCNone;
typedef struct PCJacobi { dfc_index n; dfc_index bs; void* diag; } PCJacobi;
typedef struct PCDecomposition { dfc_index n_sec; dfc_index* offset; PC** pc; } PCDecomposition;
PC* PCCreateNone(Matrix* mat, dfc_index n);
PC* PCCreateJacobi(Matrix* mat, dfc_index bs, void* cublas_handle);
PC* PCCreateDecomposition(Matrix* mat, dfc_index n, const dfc_index* offset, void* cublas_handle);
PC* PCCreateAMGX(Matrix* mat, void* options); /* returns NULL, as the reference built without USE_AMGX */
void PCSetup(PC* pc);
void PCDestroy(PC* pc);
void PCApply(PC* pc, double* x, double* y);

/* ---- Krylov (src/krylov.h:11-30) ---- */
typedef void (*KSPSolveFunc)(Matrix*, dfc_value*, dfc_value*, void*);
typedef struct Krylov {
  dfc_index max_iter;
  double atol, rtol;
  void* handle;
  KSPSolveFunc ksp_solve;
  size_t ksp_ctx_size;
  void* ksp_ctx; /* here: the persistent dfb_gmres workspace */
  void* pc;
} Krylov;
Krylov* KrylovCreateCG(dfc_index max_iter, double atol, double rtol, void* handle);
Krylov* KrylovCreateGMRES(dfc_index max_iter, double atol, double rtol, void* handle);
void KrylovDestroy(Krylov* krylov);
/* Argument order of the reference DEFINITION and of its call site (krylov.c:386, main.c:217), not of its
 * declaration (krylov.h:30): solve A x = b. */
void KrylovSolve(Krylov* krylov, Matrix* A, double* x, double* b);

/* ---- vector helpers (src/vec.h:7-10) ---- */
void VecAXPY(dfc_value a, const dfc_value* x, dfc_value* y, dfc_index n);
void VecPointwiseMult(const dfc_value* x, const dfc_value* y, dfc_value* z, dfc_index n);
void VecPointwiseDiv(const dfc_value* x, const dfc_value* y, dfc_value* z, dfc_index n);
void VecPointwiseInv(dfc_value* x, dfc_index n);

/* ---- additions of this library (not in the reference) ---- */
/* Residual history |beta_k| (k = 0..iters) of the last KrylovSolve on this solver; returns iters, or -1. */
int dfb_compat_last_history(const Krylov* krylov, double* hist, int capacity);
/* Drop every cached assembly plan that belongs to `mesh` (call before freeing or re-reading a mesh); NULL = all. */
void dfb_compat_release(const Mesh3D* mesh);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* DEDFLOW_COMPAT_H */
