/*
 * dedflow_b200.h -- C ABI of the B200-native FEM linear-system hot path (libdedflow_b200.so).
 *
 * Two layers, both extern "C", plain pointers and sizes only (no torch / C++ types):
 *
 *  (1) the CORE ABI (dfb_*): stateless kernels on raw DEVICE pointers plus two opaque handles
 *      (dfb_plan: integer assembly plan of one mesh; dfb_gmres: persistent Krylov workspace).
 *      Every function returns 0 on success or a negative dfb_status; dfb_last_error() gives
 *      the message.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream,
 *      the stream the reference runs on).
 *
 *  (2) the DROP-IN ABI: the reference's own entry points with the reference's struct layouts
 *      (declared in dedflow_compat.h) implemented on top of (1), so that reference src/main.c,
 *      Mesh.c, MeshData.c link against this library unchanged.
 *
 * Conventions shared with the reference: index = int32, value = double;
 * xg[3N] interleaved; ien[4E]; every state / Krylov vector has 6N entries laid out
 * [u: N x 3 interleaved | p: N | phi: N | T: N]  (reference src/main.c:297-319).
 * The field-split matrix is four scalar-CSR value arrays over the nodal pattern (row_ptr,col_ind):
 *   A00 (3x3): val[start*9 + ii*3*len + k*3 + jj]     A01 (3x1): val[start*3 + ii*len + k]
 *   A10 (1x3): val[start*3 + k*3 + jj]                A11 (1x1): val[start + k]
 * with start=row_ptr[i], len=row_ptr[i+1]-start, k the position of the column inside the nodal row
 * (reference src/csr_impl.cu:24-59, src/matrix_impl.cu:370-453).
 */
#ifndef DEDFLOW_B200_H
#define DEDFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

typedef enum dfb_status {
  DFB_OK = 0,
  DFB_ERR_CUDA = -1,      /* a CUDA runtime call failed */
  DFB_ERR_ARG = -2,       /* bad argument */
  DFB_ERR_OVERFLOW = -3,  /* nodal row longer than 64 (reference csr.c:10,64 asserts) or i32 index overflow */
  DFB_ERR_COLOR = -4,     /* more than max_color rounds needed */
  DFB_ERR_NODEVICE = -5,  /* no CUDA device: there is no CPU fallback */
  DFB_ERR_PEER = -6       /* peer-memory collective timed out (a rank died or left the solve early); the communicator is
                             unusable afterwards, destroy it */
} dfb_status;

const char* dfb_last_error(void);
int dfb_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long dfb_launch_count(void);
/* Variant switches (A/B measurements, tests).  `key` is one of the environment names the library reads once at load
 * (DFB_J_VARIANT, DFB_J_PAIR_ROWS, DFB_J_PAIR_NT, DFB_J_PAIR_ORDER, DFB_J_PULL_PLAIN, DFB_F_VARIANT, DFB_F_PATCH_CTAS, DFB_SPMV_G,
 * DFB_SPMV_TMA, DFB_SPMV_PEER_SPLIT, DFB_HALO_DEFER, DFB_GRAPH, DFB_GMRES_CHECK, DFB_GIVENS_DEFER, DFB_PROFILE, DFB_VERBOSE, DFB_ASSEMBLE_MODE, DFB_PC, DFB_PC_AGG, DFB_PC_DEGREE); no entry point reads the
 * environment on its launch path. */
int dfb_set_option(const char* key, const char* value);

/* ------------------------------------------------------------------------------------------
 * Sparsity pattern.  Replaces CSRAttrCreate (reference src/csr.c:143-190, host, single thread)
 * and ExpandCSRByBlockSize (src/csr_impl.cu:126-156).
 * ------------------------------------------------------------------------------------------ */
/* row_ptr[N+1] (device, out) of the scalar nodal pattern: row i = sorted unique {i} U neighbours. *nnz (host, out). */
int dfb_pattern_rows(int num_node, int num_tet, const int* d_ien, int* d_row_ptr, int* nnz, void* stream);
/* col_ind[nnz] (device, out), ascending inside each row. */
int dfb_pattern_cols(int num_node, int num_tet, const int* d_ien, const int* d_row_ptr, int* d_col_ind, void* stream);
/* blocked scalar-CSR pattern (br x bc): new_row_ptr[N*br+1], new_col_ind[nnz*br*bc].  Unlike the reference
 * (defect D1) the final row_ptr entry IS written. */
int dfb_pattern_expand(int num_node, const int* d_row_ptr, const int* d_col_ind, int br, int bc, int* d_new_row_ptr,
                       int* d_new_col_ind, void* stream);

/* ------------------------------------------------------------------------------------------
 * Element coloring.  Replaces ColorMeshTet / GetMaxColor (src/color.c:14-67, src/color_impl.cu)
 * and the batch construction of Mesh3DGenerateColorBatch (src/Mesh.c:165-206, src/indexing.cu:92-103).
 * ------------------------------------------------------------------------------------------ */
/* weight[e] = xorwow_u32(seed)[e] % 1073741823 in cuRAND's default ordering (src/color_impl.cu:185-192,225-237) */
int dfb_color_weights(int num_tet, unsigned long long seed, int* d_weight, void* stream);
/* Jones-Plassmann-Luby: color[e] = round in which e is the strict weight maximum among its still uncoloured
 * vertex-sharing neighbours.  Equal weights are broken by element id (the reference is non-deterministic there,
 * defect D2).  *num_color (host, out) = max color + 1. */
int dfb_color_jpl(int num_node, int num_tet, const int* d_ien, const int* d_weight, int max_color, int* d_color,
                  int* num_color, void* stream);
/* batch_offset[num_color+1] (HOST, out), batch_ind[E] (device, out): ascending element ids per color. */
int dfb_color_batches(int num_tet, const int* d_color, int num_color, int* h_batch_offset, int* d_batch_ind,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Assembly.  Replaces AssembleSystemTet / AssembleSystemTetFace (src/assemble.cu:1467-1964),
 * MatrixAddElemValueBlockedBatched (src/matrix.c:574-592, src/matrix_impl.cu:370-453),
 * ElemRHSLocal2Global (src/assemble.cu:188-208) and the Dirichlet kernels.
 * ------------------------------------------------------------------------------------------ */
typedef struct dfb_plan dfb_plan;

typedef enum dfb_assemble_mode {
  DFB_MODE_AUTO = 0,     /* fastest measured variant (gather) */
  DFB_MODE_GATHER = 1,   /* atomic-free: per row group, the element records are formed in shared memory and one thread
                            per upper nodal nonzero (i,j) accumulates A_ij and A_ji; every CSR value and F entry is
                            written exactly once, fixed summation order (deterministic) */
  DFB_MODE_ATOMIC = 2,   /* one launch over all elements, fp64 red.global.add scatter */
  DFB_MODE_COLORED = 3   /* one launch per color batch, plain read-modify-write in the reference's order */
} dfb_assemble_mode;

/* Integer plan of one mesh: vertex->corner lists, corner->CSR-slot map, (optional) color batches.
 * d_ien, d_row_ptr, d_col_ind are borrowed and must outlive the plan.  h_batch_offset/d_batch_ind may be NULL
 * (then DFB_MODE_COLORED is unavailable).  d_row_ptr and d_col_ind may both be NULL: the plan then serves residual
 * (F) assembly only. */
int dfb_plan_create(dfb_plan** plan, int num_node, int num_tet, const int* d_ien, const int* d_row_ptr,
                    const int* d_col_ind, int num_batch, const int* h_batch_offset, const int* d_batch_ind,
                    void* stream);
void dfb_plan_destroy(dfb_plan* plan);
/* Data-parallel assembly: only rows / F entries of the first n_rows (owned) nodes are produced by DFB_MODE_GATHER. */
int dfb_plan_set_rows(dfb_plan* plan, int n_rows);
/* bytes of device memory held by the plan */
size_t dfb_plan_bytes(const dfb_plan* plan);

/* Interior (tet) assembly: F[6N] += element residuals and/or the four sub-block value arrays += element
 * Jacobians.  F or the A-pointers may be NULL.  In GATHER mode with `overwrite` != 0 the outputs are written
 * (=) instead of accumulated (+=), which makes the preceding memset (reference main.c:44-49) unnecessary. */
int dfb_assemble_tet(const dfb_plan* plan, const double* d_xg, const double* d_wgalpha, const double* d_dwgalpha,
                     double* d_F, double* d_A00, double* d_A01, double* d_A10, double* d_A11, int mode, int overwrite,
                     void* stream);
/* Weak-BC boundary faces of one boundary group (the reference runs group 4 only, defect D13). */
int dfb_assemble_face(const dfb_plan* plan, int num_face, const int* d_f2e, const int* d_forn, const double* d_xg,
                      const double* d_wgalpha, const double* d_dwgalpha, double* d_F, double* d_A00, double* d_A01,
                      double* d_A10, double* d_A11, void* stream);
/* Strong Dirichlet: b[node*shape + ic] = 0 for every ic with bctype[ic]==1 (src/dirichlet.c:31-42). */
int dfb_dirichlet_vec(int num_bnode, const int* d_bnode, int shape, const int* h_bctype, double* d_b, void* stream);
/* rows node*3+ic of A00 -> unit rows, of A01 -> zero rows (src/dirichlet.c:47-61, src/matrix.c:449-469). */
int dfb_dirichlet_mat(int num_bnode, const int* d_bnode, int shape, const int* h_bctype, int num_node,
                      const int* d_row_ptr, const int* d_col_ind, double* d_A00, double* d_A01, void* stream);

/* ------------------------------------------------------------------------------------------
 * Solve.  Replaces MatrixFSAMVPBY/MatVec (src/matrix.c:471-524, four cuSPARSE SpMVs), PCSetup/PCApply
 * (src/pc.c:44-147) and GMRESSolvePrivate (src/krylov.c:56-334).
 * ------------------------------------------------------------------------------------------ */
/* y[0:4N) = beta*y[0:4N) + alpha * A x ; y[4N:6N) untouched (defect D4).  One fused kernel over the 4 blocks. */
int dfb_spmv_fs(int num_node, const int* d_row_ptr, const int* d_col_ind, const double* d_A00, const double* d_A01,
                const double* d_A10, const double* d_A11, double alpha, const double* d_x, double beta, double* d_y,
                void* stream);
/* dinv00[9N]: (B^-1)^T of the nodal 3x3 diagonal blocks of A00 in the layout the reference's gemv consumes
 * (defect D3); dinv11[N] = 1/diag(A11). */
int dfb_pc_setup(int num_node, const int* d_row_ptr, const int* d_col_ind, const double* d_A00, const double* d_A11,
                 double* d_dinv00, double* d_dinv11, void* stream);
/* y = P^-1 x on 6N vectors: block-Jacobi on u, Jacobi on p, identity on phi and T. */
int dfb_pc_apply(int num_node, const double* d_dinv00, const double* d_dinv11, const double* d_x, double* d_y,
                 void* stream);

typedef struct dfb_gmres dfb_gmres;
/* Persistent workspace for right-preconditioned, un-restarted GMRES(max_iter) on 6N vectors. */
int dfb_gmres_create(dfb_gmres** ws, int num_node, int max_iter);
void dfb_gmres_destroy(dfb_gmres* ws);
size_t dfb_gmres_bytes(const dfb_gmres* ws);
/* Per-kernel CUDA-event times of the LAST solve of this workspace, "name:launches:total_ms;..." (empty unless the option
 * DFB_PROFILE is non-zero: 1 also prints the table to stderr, 2 every launch, -1 only records). */
int dfb_gmres_profile(const dfb_gmres* w, char* buf, int capacity);
/* Optional data-parallel hooks (multi-GPU, one process per GPU).  Local node numbering of a rank is
 * [interior-owned | boundary-owned | ghost]: rows [0,n_own) are owned and assembled completely on this rank, rows
 * [0,n_interior) reference no ghost column.  Inner products run over the owned rows and are summed over ranks with
 * `allreduce` (device buffer of doubles, in place, enqueued on `stream`); before every mat-vec the ghost entries of the
 * input vector are refreshed: halo_begin() starts the exchange (it may run concurrently on another stream), the
 * interior rows are multiplied meanwhile, halo_end() makes `stream` wait for the ghosts, then the boundary rows follow. */
typedef struct dfb_parallel_ops {
  int n_own;
  int n_interior;
  int (*allreduce)(double* d_buf, int count, void* stream, void* user);
  int (*halo_begin)(double* d_x, void* stream, void* user);
  int (*halo_end)(double* d_x, void* stream, void* user);
  void* user;
  /* Optional (may be NULL): dfb_comm_p2p_view().  When set, the collectives INSIDE a GMRES iteration are fused into the
   * compute kernels over NVLink peer memory (partial sums and halo values are stored directly into the peers' memory,
   * summed in rank order): no NCCL call and no extra launch per reduction.  The callbacks above are still used outside
   * the iteration loop (initial residual, final ghost refresh). */
  const void* p2p;
  /* Needed when p2p is NULL (NCCL path): like halo_begin, for a vector in the solver's interleaved layout x[4*node + c]
   * (c = 0..2 velocity, 3 pressure; 4 * num_local_nodes doubles).  halo_end is shared. */
  int (*halo_begin_aos)(double* d_x4, void* stream, void* user);
} dfb_parallel_ops;
int dfb_gmres_set_parallel(dfb_gmres* ws, const dfb_parallel_ops* ops);
/* ------------------------------------------------------------------------------------------
 * Stronger preconditioner (opt-in; the slot the reference reserves for AMGX on the pressure block, src/pc.c:160-235,
 * src/krylov.c:392-453, compiled out there): block lower-triangular with an additive two-level approximation of the pressure
 * Schur complement S = A11 - A10 D^-1 A01 (dedflow_b200/csrc/pc2.cu).  One GPU.
 *   create: once per mesh (aggregates of agg_cells^3 average node spacings from the coordinates; cheb_degree Chebyshev steps
 *           on the Galerkin coarse matrix).  d_row_ptr / d_col_ind are borrowed.
 *   setup : once per solve (the Jacobian changes every Newton iteration); dfb_gmres_solve* calls it when a workspace carries
 *           the preconditioner (dfb_gmres_set_pc2; NULL restores the reference's block-Jacobi).
 *   apply : y = P^-1 x on 6N ABI-layout vectors (tests / drop-in PCApply). */
typedef struct dfb_pc2 dfb_pc2;
int dfb_pc2_create(dfb_pc2** out, int num_node, const int* d_row_ptr, const int* d_col_ind, const double* d_xg, int agg_cells,
                   int cheb_degree, void* stream);
int dfb_pc2_info(const dfb_pc2* pc, int* num_aggregates, int* coarse_nnz);
int dfb_pc2_setup(dfb_pc2* pc, const double* d_A00, const double* d_A01, const double* d_A10, const double* d_A11, void* stream);
int dfb_pc2_apply(dfb_pc2* pc, const double* d_A10, const double* d_x, double* d_y, void* stream);
void dfb_pc2_destroy(dfb_pc2* pc);
int dfb_gmres_set_pc2(dfb_gmres* ws, dfb_pc2* pc);

/* Solve A x = b (x in/out, b in; both 6N device vectors).  Convergence is tested only when (iter+1)%20==0 against
 * |r| < atol || |r| < (|r0| + 1e-16)*rtol (src/krylov.c:281-290, defect D10).  res_hist (HOST, may be NULL) receives
 * max_iter+1 entries: |beta[k]| for k = 0..iters.  *iters (host, out). */
int dfb_gmres_solve(dfb_gmres* ws, int num_node, const int* d_row_ptr, const int* d_col_ind, const double* d_A00,
                    const double* d_A01, const double* d_A10, const double* d_A11, double* d_x, const double* d_b,
                    double atol, double rtol, int* iters, double* res_hist, void* stream);

/* Same solve with the preconditioner arrays supplied by the caller (layouts of dfb_pc_setup); both NULL = set up
 * internally, which is what dfb_gmres_solve does. */
int dfb_gmres_solve_pc(dfb_gmres* ws, int num_node, const int* d_row_ptr, const int* d_col_ind, const double* d_A00,
                       const double* d_A01, const double* d_A10, const double* d_A11, const double* d_dinv00,
                       const double* d_dinv11, double* d_x, const double* d_b, double atol, double rtol, int* iters,
                       double* res_hist, void* stream);

/* ------------------------------------------------------------------------------------------
 * Newton / generalised-alpha vector work of the driver around the path.  Replaces the ~25
 * cuBLAS axpy/copy/scal/memset calls and the 8 blocking cublasDnrm2 of one Newton iteration (src/main.c:107-118, 127-130,
 * 226-246, 262-265, 544-545, 559-563).  All vectors are 6N device vectors.
 * ------------------------------------------------------------------------------------------ */
/* dwgalpha = (1-am) dwgold + am dwg, pressure slot = dwg;  wgalpha = wgold + dt af (1-g) dwgold + dt af g dwg, pressure slot = 0 */
int dfb_genalpha_stage(int num_node, const double* d_wgold, const double* d_dwgold, const double* d_dwg, double* d_wgalpha,
                       double* d_dwgalpha, void* stream);
/* dwg -= dx */
int dfb_newton_update(int num_node, const double* d_dx, double* d_dwg, void* stream);
/* predictor: dwg[u, phi, T] *= (g-1)/g (pressure slot untouched) */
int dfb_genalpha_predict(int num_node, double* d_dwg, void* stream);
/* corrector: wgold[u, phi, T] += dt (1-g) dwgold + dt g dwg; dwgold = dwg */
int dfb_genalpha_correct(int num_node, double* d_wgold, double* d_dwgold, const double* d_dwg, void* stream);
/* d_out4 (device) = sums of squares of the blocks [u | p | phi | T] of F over the first n_own nodes (data-parallel callers sum
 * them over ranks); one kernel, fixed summation order */
int dfb_block_sumsq(int num_node, int n_own, const double* d_F, double* d_out4, void* stream);
/* h_out4 (HOST) = the four block norms of F (synchronises `stream`, like the reference's cublasDnrm2 with a host result) */
int dfb_block_norms(int num_node, const double* d_F, double* h_out4, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel communicator (NCCL over NVLink 5 / NVSwitch, resolved with dlopen at run time so that the
 * single-GPU library has no NCCL dependency).  One process per GPU; the 128-byte unique id is created on rank 0
 * and distributed by the host (torch.distributed in the bench).  New work: the reference is single-GPU
 * (SURVEY.md §0, §8e).
 * ------------------------------------------------------------------------------------------ */
typedef struct dfb_comm dfb_comm;
int dfb_comm_unique_id(void* id128);
int dfb_comm_create(dfb_comm** comm, int rank, int nranks, const void* id128);
void dfb_comm_destroy(dfb_comm* comm);
/* Halo plan in local node ids: for neighbour q, send the (u,p) of owned nodes send_nodes[send_offset[q]..) and
 * receive into ghost nodes recv_nodes[recv_offset[q]..).  Both sides list the nodes in ascending GLOBAL id. */
int dfb_comm_set_halo(dfb_comm* comm, int num_local_nodes, int n_neighbors, const int* h_neighbor_rank,
                      const int* h_send_offset, const int* h_send_nodes, const int* h_recv_offset,
                      const int* h_recv_nodes);
/* the three dfb_parallel_ops callbacks; `user` is the dfb_comm* */
int dfb_comm_allreduce(double* d_buf, int count, void* stream, void* user);
int dfb_comm_halo_begin(double* d_x, void* stream, void* user);
int dfb_comm_halo_end(double* d_x, void* stream, void* user);
int dfb_comm_halo_begin_aos(double* d_x4, void* stream, void* user);
/* blocking convenience: refresh the ghosts of a 6N-layout vector */
int dfb_comm_halo(dfb_comm* comm, double* d_x, void* stream);
/* Peer-memory (CUDA IPC over NVLink) mode, one node only.  (1) every rank allocates its shared region (mailbox + the
 * preconditioned Krylov vector z with num_local_nodes*6 doubles) and gets a 64-byte IPC handle; (2) the host gathers the
 * handles of all ranks (rank order) and every rank connects; (3) the halo plan gains, for every node it sends, the node's
 * local id on the receiving rank and that rank's local node count.  dfb_comm_p2p_view() is then passed in
 * dfb_parallel_ops.p2p.  Call dfb_comm_set_halo first. */
int dfb_comm_p2p_alloc(dfb_comm* comm, void* handle64);
int dfb_comm_p2p_connect(dfb_comm* comm, const void* handles /* nranks x 64 bytes */, const int* h_remote_nodes,
                         const int* h_neighbor_num_local);
const void* dfb_comm_p2p_view(dfb_comm* comm);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* DEDFLOW_B200_H */
