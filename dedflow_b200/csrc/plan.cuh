// plan.cuh -- the integer "assembly plan" of one mesh, built once (setup.cu) and reused by every assembly.
#pragma once
#include <vector>

#include "common.cuh"

struct dfb_plan {
  int N = 0, E = 0;
  int n_rows = 0;                // rows produced by the gather kernels (== N unless data-parallel)
  const int* ien = nullptr;      // [4E] borrowed
  const int* row_ptr = nullptr;  // [N+1] nodal pattern, borrowed
  const int* col_ind = nullptr;  // [nnz] borrowed
  int* v2c_ptr = nullptr;        // [N+1] offsets into v2c
  int* v2c = nullptr;            // [4E] corner ids (e*4+a) of every node, ascending
  u8* slot = nullptr;            // [16E] slot[(e*4+a)*4+b] = position of ien[e,b] in nodal row ien[e,a]
  int max_valence = 0;           // max corners per node
  int max_row_len = 0;           // max nodal row length
  // color batches (only for DFB_MODE_COLORED)
  int num_batch = 0;
  std::vector<int> batch_offset;
  const int* batch_ind = nullptr;  // borrowed
  // lazily allocated element-residual scratch for the deterministic F gather (24 doubles per element)
  mutable f64* elemF = nullptr;      // corner residuals in node-major order: slot cpos[e*4+a]
  mutable int* cpos = nullptr;       // [4E] position of corner e*4+a in v2c
  mutable size_t elemF_bytes = 0;
  // lazily built work lists of the PULL Jacobian assembly (setup.cu build_pull): one work item per off-diagonal nodal
  // nonzero plus four "virtual" items per diagonal entry (its contributions dealt round-robin), rows padded to multiples
  // of four items so that the four diagonal items of a row sit in one aligned lane quad.
  mutable int n_items = 0;            // total work items (multiple of 4)
  mutable int* row_item = nullptr;    // [N+1] first item of every row
  mutable uint2* item_meta = nullptr; // [n_items] {row (0xffffffff: padding), k | 0x100 for a diagonal item}
  mutable int* item_ptr = nullptr;    // [n_items+1] offsets into contrib
  mutable u32* contrib = nullptr;     // [16E] corner*4 + b, ascending inside an item
  mutable f64* prec = nullptr;        // [48E] element records of the pull assembly (assemble.cu k_jprep2)
  // staged pull: the rows are cut into groups of PULL_ROWS; one CTA per group stages the group's distinct element records in
  // shared memory.  cta_elems = ascending distinct element ids per group, contrib16 = (local record index << 4 | a << 2 | b)
  // parallel to contrib (0xffff: group too large for shared memory, the CTA reads the records from global memory instead).
  mutable int n_cta = 0;
  mutable int* cta_elem_ptr = nullptr;   // [n_cta+1]
  mutable int* cta_elems = nullptr;
  mutable unsigned short* contrib16 = nullptr;   // [16E]
  mutable int max_cta_elems = 0;
  mutable size_t pull_bytes = 0;
  mutable int items_rows = -1, items_active = 0;  // cache: items of the first n_rows rows
  // PAIR Jacobian assembly (setup.cu build_pairs, assemble.cu k_pairJ; the default).  The first n_rows rows, ordered along a
  // Morton curve of their coordinates, are cut into groups of pr_rows (a multiple of 8); one CTA per group.  Work items of a group: first 4 "virtual" items per row for the diagonal entry
  // (4*pr_rows: whole warps), then one item per UPPER off-diagonal nonzero (i,j), j > i, which also produces (j,i).
  mutable int pr_state = 0;            // 0: not built, 1: usable, -1: a group does not fit shared memory (fall back to pull)
  mutable int pr_built_rows = -1;     // n_rows the lists were built for (rebuilt when dfb_plan_set_rows changes it)
  mutable int pr_rows = 0, pr_n_cta = 0, pr_n_items = 0, pr_max_elems = 0;
  mutable int4* pr_grp = nullptr;       // [pr_n_cta] {first staged element, #elements, first item, #items}: one 16-byte load per CTA
  mutable int4* pr_enodes = nullptr;    // the four node ids of every staged element (parallel to pr_elems): skips elems -> ien
  mutable int* pr_grp_item = nullptr;   // [pr_n_cta+1] first item of every group (build time only, folded into pr_grp)
  mutable uint2* pr_meta = nullptr;     // [pr_n_items] {row i (0xffffffff: padding), k_ij | 0x100 diagonal | k_ji << 16}
  mutable int* pr_item_ptr = nullptr;   // [pr_n_items+1] offsets into pr_contrib
  mutable unsigned short* pr_contrib = nullptr;   // local record index << 4 | a << 2 | b, ascending (element, a, b) per item
  mutable int* pr_elem_ptr = nullptr;   // [pr_n_cta+1]                              (build time only, folded into pr_grp)
  mutable int* pr_elems = nullptr;      // ascending distinct element ids per group  (build time only, see pr_enodes)
  mutable size_t pr_bytes = 0;
  // PATCH residual assembly (setup.cu build_fpatch, assemble.cu k_patchF; the F path of DFB_MODE_GATHER).  The elements,
  // ordered along a Morton curve of their centroids, are cut into patches of FP_PE; one CTA per patch.  A patch knows its
  // distinct nodes ("patch-nodes", numbered pn = fp_hdr[p].x + k), the local node indices of its elements, and its corners
  // sorted by node; it writes ONE 48-byte partial residual per patch-node (about 2 per mesh node instead of 24 corner records),
  // which k_gatherF2 sums per node in ascending pn order.
  mutable int fp_state = 0;             // 0: not built, 1: usable, -1: a patch touches too many nodes (scratch variant is used)
  mutable int fp_n_patch = 0, fp_n_pn = 0, fp_max_nodes = 0;
  mutable int2* fp_hdr = nullptr;       // [n_patch] {first patch-node, number of nodes}
  mutable int* fp_nodes = nullptr;      // [n_pn] global node of every patch-node (ascending inside a patch)
  mutable ushort4* fp_lnode = nullptr;  // [n_patch * FP_PE] local node indices of every element, patch order (0xffff: padding)
  mutable unsigned short* fp_corner = nullptr;   // [n_patch * 4 FP_PE] local corners (element * 4 + a) sorted by (node, corner)
  mutable unsigned short* fp_cstart = nullptr;   // [n_pn + n_patch] first sorted corner of every patch-node (+ one end entry per patch)
  mutable f64* fp_part = nullptr;       // [6 n_pn] partial residuals
  mutable int* fp_np_ptr = nullptr;     // [N+1] patch-nodes of every mesh node ...
  mutable int* fp_np = nullptr;         // [n_pn] ... ascending
  mutable size_t fp_bytes = 0;
};

namespace dfb {
constexpr int PULL_ROWS = 6;         // rows per CTA of the staged pull assembly
constexpr int PULL_MAX_STAGED = 160; // most element records a CTA stages (160 x 384 B = 60 KB)
int build_v2c(int N, int E, const int* d_ien, int** d_ptr_out, int** d_v2c_out, cudaStream_t st);
int build_pull(const dfb_plan* plan, cudaStream_t st);
constexpr int PAIR_MAX_STAGED = 600; // most element records a CTA of the pair assembly stages (600 x 368 B = 216 KB)
int build_pairs(const dfb_plan* plan, int rows_per_cta, const f64* d_xg, cudaStream_t st);
void free_pairs(const dfb_plan* plan);
constexpr int FP_PE = 256;          // elements per patch of the residual assembly
constexpr int FP_MAX_NODES = 448;   // most distinct nodes a patch may touch (448 x 15 doubles + 256 x 25 doubles of staging = 105 KB)
int build_fpatch(const dfb_plan* plan, const f64* d_xg, cudaStream_t st);
void free_fpatch(const dfb_plan* plan);
}
