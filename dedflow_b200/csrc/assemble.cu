// assemble.cu -- element assembly of the field-split system (F and the four CSR sub-blocks) on B200.
//
// Replaces (reference paths relative to /root/reference/src):
//   assemble.cu:1467-1762  AssembleSystemTet: per color batch ~25 launches (geometry, batched LU, 8 gathers,
//                          3 batched GEMMs, weak form, memset + 576-double elem_J round trip, 5 scatters)
//   assemble.cu:1764-1964  AssembleSystemTetFace (num_color x 7 masked scatter launches)
//   matrix_impl.cu:370-453 SetBlockValueToSubmatKernel (linear search + scattered 8-byte RMW)
//   dirichlet_impl.cu:15-37, matrix_impl.cu:6-23  Dirichlet rows
//
// B200 design:
//   * J, GATHER (default), three atomic-free variants that write every CSR value exactly once in a fixed order (no
//     atomics, no colors, no memset, no elem_J round trip, deterministic):
//       pairs (default, k_pairJ): one CTA per Morton-ordered group of 8 rows forms the records of the elements around the
//         group in shared memory; one thread per UPPER nodal nonzero (i,j) accumulates both A_ij and A_ji in registers;
//       pull (k_jprep2 + k_pullJ_staged): per-element records in global memory, one thread per nodal nonzero;
//       fused (k_rowJ): one warp per nodal row, lanes = the row's corners, slot-peer reduction through shared memory.
//   * J, ATOMIC: one thread per corner, red.global.add.f64 scatter through the precomputed slot map.
//   * J, COLORED: same kernel without atomics, one launch per color batch (the reference's structure).
//   * F: one thread per element evaluates the residual once; GATHER writes the four 48-byte corner residuals to a
//     node-major scratch and a node-gather kernel streams and sums them in fixed order; ATOMIC / COLORED scatter directly.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "elem_math.cuh"
#include "plan.cuh"

namespace dfb {
using namespace em;

constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------------------------------------------------
// loads
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_nodes(const int* __restrict__ ien, int e, int nodes[4]) {
  int4 nd = __ldg(reinterpret_cast<const int4*>(ien) + e);
  nodes[0] = nd.x; nodes[1] = nd.y; nodes[2] = nd.z; nodes[3] = nd.w;
}

__device__ __forceinline__ void load_xyz(const f64* __restrict__ v, const int nodes[4], f64 out[4][3]) {
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const f64* p = v + (size_t)nodes[a] * 3;
    out[a][0] = __ldg(p); out[a][1] = __ldg(p + 1); out[a][2] = __ldg(p + 2);
  }
}

// CSR positions of the 4x4 block of (row node with nodal row [start, start+len), slot k)
template <int OP>  // 0: plain +=, 1: atomic +=, 2: =
__device__ __forceinline__ void put(f64* p, f64 v) {
  if (OP == 0) *p += v;
  else if (OP == 1) atomicAdd(p, v);
  else *p = v;
}

template <int OP>
__device__ __forceinline__ void scatter_block(f64* __restrict__ A00, f64* __restrict__ A01, f64* __restrict__ A10,
                                              f64* __restrict__ A11, size_t start, int len, int k, const f64 blk[16]) {
  f64* p00 = A00 + start * 9 + (size_t)k * 3;
#pragma unroll
  for (int ii = 0; ii < 3; ii++) {
#pragma unroll
    for (int jj = 0; jj < 3; jj++) put<OP>(p00 + (size_t)ii * len * 3 + jj, blk[ii * 4 + jj]);
    put<OP>(A01 + start * 3 + (size_t)ii * len + k, blk[ii * 4 + 3]);
    put<OP>(A10 + start * 3 + (size_t)k * 3 + ii, blk[12 + ii]);
  }
  put<OP>(A11 + start + k, blk[15]);
}

// ------------------------------------------------------------------------------------------------------------
// F: one thread per element
// ------------------------------------------------------------------------------------------------------------
// position of every corner (e*4+a) inside the node-sorted corner list v2c: the element kernel writes its four 48-byte corner
// residuals in NODE-MAJOR order, so that the node gather streams contiguous memory (every fetched sector fully used)
__global__ void k_corner_pos(int n_corner, const int* __restrict__ v2c, int* __restrict__ cpos) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_corner) cpos[v2c[p]] = p;
}

// 236 registers, 2 CTAs per SM; capping at 168 registers (3 CTAs) spills 540 bytes and measured slower (218 vs 193 us at 1M tets)
template <int MODE>  // 0: write scratch[cpos[corner]*6..] (node-major), 1: atomic scatter, 2: plain scatter of a color batch
__global__ void __launch_bounds__(128) k_elemF(int n, const int* __restrict__ elem_ids, int N, const int* __restrict__ ien,
                                               const f64* __restrict__ xg, const f64* __restrict__ wg,
                                               const f64* __restrict__ dwg, f64* __restrict__ out, const int* __restrict__ cpos) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int e = elem_ids ? elem_ids[t] : t;
  int nodes[4];
  load_nodes(ien, e, nodes);
  f64 x[4][3], val[6][4], dval[6][4];
  load_xyz(xg, nodes, x);
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int nd = nodes[a];
    const f64* pu = wg + (size_t)nd * 3;
    const f64* pd = dwg + (size_t)nd * 3;
    val[0][a] = __ldg(pu); val[1][a] = __ldg(pu + 1); val[2][a] = __ldg(pu + 2);
    dval[0][a] = __ldg(pd); dval[1][a] = __ldg(pd + 1); dval[2][a] = __ldg(pd + 2);
    const f64 p = __ldg(dwg + (size_t)3 * N + nd);  // pressure lives in the increment vector (defect D6)
    val[3][a] = p; dval[3][a] = p;
    val[4][a] = __ldg(wg + (size_t)4 * N + nd); dval[4][a] = __ldg(dwg + (size_t)4 * N + nd);
    val[5][a] = __ldg(wg + (size_t)5 * N + nd); dval[5][a] = __ldg(dwg + (size_t)5 * N + nd);
  }
  Geom g;
  geometry(x, g);
  f64 eF[4][6];
  residual(g, val, dval, eF);
  if (MODE == 0) {
    const int4 cp = __ldg(reinterpret_cast<const int4*>(cpos) + e);
    const int pos[4] = {cp.x, cp.y, cp.z, cp.w};
#pragma unroll
    for (int a = 0; a < 4; a++) {
      f64* dst = out + (size_t)pos[a] * 6;
#pragma unroll
      for (int i = 0; i < 6; i += 2) *reinterpret_cast<double2*>(dst + i) = make_double2(eF[a][i], eF[a][i + 1]);
    }
  } else {
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const size_t nd = (size_t)nodes[a];
      put<MODE == 1 ? 1 : 0>(out + nd * 3 + 0, eF[a][0]);
      put<MODE == 1 ? 1 : 0>(out + nd * 3 + 1, eF[a][1]);
      put<MODE == 1 ? 1 : 0>(out + nd * 3 + 2, eF[a][2]);
      put<MODE == 1 ? 1 : 0>(out + (size_t)3 * N + nd, eF[a][3]);
      put<MODE == 1 ? 1 : 0>(out + (size_t)4 * N + nd, eF[a][4]);
      put<MODE == 1 ? 1 : 0>(out + (size_t)5 * N + nd, eF[a][5]);
    }
  }
}

// node gather: F[node] (+)= sum over the node's corners, ascending corner id
__global__ void k_gatherF(int N, int n_rows, const int* __restrict__ v2c_ptr, const f64* __restrict__ scratch,
                          f64* __restrict__ F, int overwrite) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  f64 s[6] = {0, 0, 0, 0, 0, 0};
  for (int p = v2c_ptr[i]; p < v2c_ptr[i + 1]; p++) {
    const f64* src = scratch + (size_t)p * 6;
    double2 a = *reinterpret_cast<const double2*>(src), b = *reinterpret_cast<const double2*>(src + 2),
            c = *reinterpret_cast<const double2*>(src + 4);
    s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y;
  }
  f64* fu = F + (size_t)i * 3;
  if (overwrite) {
    fu[0] = s[0]; fu[1] = s[1]; fu[2] = s[2];
    F[(size_t)3 * N + i] = s[3]; F[(size_t)4 * N + i] = s[4]; F[(size_t)5 * N + i] = s[5];
  } else {
    fu[0] += s[0]; fu[1] += s[1]; fu[2] += s[2];
    F[(size_t)3 * N + i] += s[3]; F[(size_t)4 * N + i] += s[4]; F[(size_t)5 * N + i] += s[5];
  }
}

// ------------------------------------------------------------------------------------------------------------
// F, PATCH variant (default of DFB_MODE_GATHER; plan.cuh fp_*, setup.cu build_fpatch).  One CTA per patch of FP_PE = 256
// Morton-ordered elements, three phases:
//   A. the patch's ~90-130 distinct NODES are staged once in shared memory (14 doubles each, thread t <-> node t: 14
//      independent loads in flight per thread) instead of 56 gathers per element;
//   B. every thread evaluates the residual of two elements out of the staged node records (no global load in the FP64 phase:
//      the dependent-load chain ien -> node ids -> 56 values that kept the FP64 pipe at 40 % is gone) and parks the 24 results in
//      a staging area (stride 25 doubles: conflict-free);
//   C. one thread per (patch-node, component) sums that node's corners in ascending (element, a) order -- the patch's corners
//      are pre-sorted by node -- and writes ONE partial per patch-node: 48 B x ~2 per mesh node instead of 24 corner records x
//      48 B (the 2 x 192 MB scratch round trip of k_elemF + k_gatherF becomes 2 x 17 MB).
// k_gatherF2 then adds the ~2 partials of every node in ascending patch order.  Fixed order everywhere: deterministic.
// ------------------------------------------------------------------------------------------------------------
constexpr int FP_SN = em::NREC + 1;   // shared-memory stride of a node record (odd: bank spreading)
constexpr int FP_SE = 25;             // stride of an element's 24 staged results

template <int MINB>   // resident CTAs per SM the register budget is set for: 3 -> 168 registers (some spills), 2 -> 255
__global__ void __launch_bounds__(128, MINB) k_patchF(int N, const int2* __restrict__ hdr, const int* __restrict__ pnodes,
                                                   const ushort4* __restrict__ lnode, const unsigned short* __restrict__ corner,
                                                   const unsigned short* __restrict__ cstart, const f64* __restrict__ xg,
                                                   const f64* __restrict__ wg, const f64* __restrict__ dwg, f64* __restrict__ part,
                                                   int max_nodes) {
  extern __shared__ __align__(16) f64 fp_smem[];
  f64* sn = fp_smem;                               // [max_nodes][FP_SN]
  f64* se = fp_smem + (size_t)max_nodes * FP_SN;   // [FP_PE][FP_SE]
  unsigned short* scl = reinterpret_cast<unsigned short*>(se + (size_t)FP_PE * FP_SE);   // [4 FP_PE] corners sorted by node
  unsigned short* scs = scl + 4 * FP_PE;                                                  // [max_nodes + 1] their ranges
  const int p = blockIdx.x;
  const int2 h = __ldg(hdr + p);
  const int pb = h.x, nn = h.y;
  // element connectivity of phase B and the corner lists of phase C are fetched together with the node ids of phase A
  // (independent chains, all coalesced)
  const ushort4 ln0 = lnode[(size_t)p * FP_PE + threadIdx.x];
  const ushort4 ln1 = lnode[(size_t)p * FP_PE + 128 + threadIdx.x];
  {
    const uint4 c8 = __ldg(reinterpret_cast<const uint4*>(corner + (size_t)p * 4 * FP_PE) + threadIdx.x);   // 8 corners per thread
    reinterpret_cast<uint4*>(scl)[threadIdx.x] = c8;
    const unsigned short* cs = cstart + (size_t)pb + p;
    for (int k = threadIdx.x; k <= nn; k += 128) scs[k] = cs[k];
  }
  // ---- A: node records (thread t <-> nodes t and t + 128: up to 28 independent loads in flight) ----
  for (int k0 = 0; k0 < nn; k0 += 256) {
    f64 v[2][em::NREC];
    int kk[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      kk[r] = k0 + r * 128 + threadIdx.x;
      if (kk[r] < nn) {
        const size_t nd = (size_t)__ldg(pnodes + pb + kk[r]);
        const f64* px = xg + nd * 3;
        const f64* pu = wg + nd * 3;
        const f64* pd = dwg + nd * 3;
        v[r][0] = __ldg(px); v[r][1] = __ldg(px + 1); v[r][2] = __ldg(px + 2);
        v[r][3] = __ldg(pu); v[r][4] = __ldg(pu + 1); v[r][5] = __ldg(pu + 2);
        v[r][6] = __ldg(pd); v[r][7] = __ldg(pd + 1); v[r][8] = __ldg(pd + 2);
        v[r][9] = __ldg(dwg + (size_t)3 * N + nd);
        v[r][10] = __ldg(wg + (size_t)4 * N + nd); v[r][11] = __ldg(dwg + (size_t)4 * N + nd);
        v[r][12] = __ldg(wg + (size_t)5 * N + nd); v[r][13] = __ldg(dwg + (size_t)5 * N + nd);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (kk[r] < nn) {
        f64* rec = sn + (size_t)kk[r] * FP_SN;
#pragma unroll
        for (int c = 0; c < em::NREC; c++) rec[c] = v[r][c];
      }
  }
  __syncthreads();
  // ---- B: two elements per thread ----
#pragma unroll 1
  for (int j = 0; j < 2; j++) {
    const ushort4 ln = j ? ln1 : ln0;
    if (ln.x == 0xffffu) continue;   // padding of the last patch
    const f64* n0 = sn + (size_t)ln.x * FP_SN;
    const f64* n1 = sn + (size_t)ln.y * FP_SN;
    const f64* n2 = sn + (size_t)ln.z * FP_SN;
    const f64* n3 = sn + (size_t)ln.w * FP_SN;
    f64 x[4][3];
#pragma unroll
    for (int d = 0; d < 3; d++) { x[0][d] = n0[d]; x[1][d] = n1[d]; x[2][d] = n2[d]; x[3][d] = n3[d]; }
    Geom g;
    geometry(x, g);
    residual_rec(g, n0, n1, n2, n3, se + (size_t)(j * 128 + threadIdx.x) * FP_SE);
  }
  __syncthreads();
  // ---- C: one thread per patch-node, six independent sums over the node's corners in ascending (element, a) order ----
  for (int k = threadIdx.x; k < nn; k += 128) {
    const int b = scs[k], e = scs[k + 1];
    f64 s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = b; i < e; i++) {
      const int cr = scl[i];
      const f64* src = se + (cr >> 2) * FP_SE + (cr & 3) * 6;
#pragma unroll
      for (int c = 0; c < 6; c++) s[c] += src[c];
    }
    f64* dst = part + (size_t)(pb + k) * 6;
#pragma unroll
    for (int c = 0; c < 6; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(s[c], s[c + 1]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// F, PATCH variant, persistent and double-buffered (DFB_F_VARIANT=pipe).  Same three phases as k_patchF, but a CTA walks
// patches p, p + grid, ... and the node records of the NEXT patch are copied into a second shared-memory buffer by cp.async
// (LDGSTS: global -> shared without passing through registers) while the FP64 phase of the current patch runs; the node ids of
// the patch after that travel in two registers.  k_patchF's stall samples were 26 % node staging / 52 % FP64 / 22 % sums: the
// staging latency is what this removes.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// issue the copies of one node record (14 doubles from five places) into shared memory
__device__ __forceinline__ void stage_node_async(f64* rec, size_t nd, int N, const f64* __restrict__ xg, const f64* __restrict__ wg,
                                                 const f64* __restrict__ dwg) {
  const f64* px = xg + nd * 3;
  const f64* pu = wg + nd * 3;
  const f64* pd = dwg + nd * 3;
#pragma unroll
  for (int c = 0; c < 3; c++) { cp_async8(rec + c, px + c); cp_async8(rec + 3 + c, pu + c); cp_async8(rec + 6 + c, pd + c); }
  cp_async8(rec + 9, dwg + (size_t)3 * N + nd);
  cp_async8(rec + 10, wg + (size_t)4 * N + nd);
  cp_async8(rec + 11, dwg + (size_t)4 * N + nd);
  cp_async8(rec + 12, wg + (size_t)5 * N + nd);
  cp_async8(rec + 13, dwg + (size_t)5 * N + nd);
}

__global__ void __launch_bounds__(128, 2) k_patchF_pipe(int N, int n_patch, const int2* __restrict__ hdr, const int* __restrict__ pnodes,
                                                        const ushort4* __restrict__ lnode, const unsigned short* __restrict__ corner,
                                                        const unsigned short* __restrict__ cstart, const f64* __restrict__ xg,
                                                        const f64* __restrict__ wg, const f64* __restrict__ dwg,
                                                        f64* __restrict__ part, int max_nodes) {
  extern __shared__ __align__(16) f64 fp_smem[];
  // [2][max_nodes][FP_SN] node records | [FP_PE][FP_SE] staging | [2][4 FP_PE] corner lists | [2][max_nodes + 2] corner ranges
  f64* sn0 = fp_smem;
  f64* se = fp_smem + (size_t)2 * max_nodes * FP_SN;
  unsigned short* scl0 = reinterpret_cast<unsigned short*>(se + (size_t)FP_PE * FP_SE);
  unsigned short* scs0 = scl0 + 2 * 4 * FP_PE;
  const int tid = threadIdx.x, stride = gridDim.x;
  int p = blockIdx.x;
  if (p >= n_patch) return;
  // ---- prologue: patch p into buffer 0 (waited for), node ids of patch p + stride into registers ----
  int2 h = __ldg(hdr + p);
  {
    for (int k = tid; k < h.y; k += 128) stage_node_async(sn0 + (size_t)k * FP_SN, (size_t)__ldg(pnodes + h.x + k), N, xg, wg, dwg);
    cp_async16(scl0 + 8 * tid, corner + (size_t)p * 4 * FP_PE + 8 * tid);
    const unsigned short* cs = cstart + (size_t)h.x + p;
    for (int k = tid; k <= h.y; k += 128) scs0[k] = cs[k];
  }
  ushort4 ln0 = lnode[(size_t)p * FP_PE + tid], ln1 = lnode[(size_t)p * FP_PE + 128 + tid];
  int pn = p + stride;
  int2 hn = make_int2(0, 0);
  int nid[2] = {0, 0};
  if (pn < n_patch) {
    hn = __ldg(hdr + pn);
    if (tid < hn.y) nid[0] = __ldg(pnodes + hn.x + tid);
    if (tid + 128 < hn.y) nid[1] = __ldg(pnodes + hn.x + tid + 128);
  }
  cp_async_wait_all();
  __syncthreads();
  for (int it = 0; p < n_patch; it++) {
    const int b = it & 1;
    f64* sn = sn0 + (size_t)b * max_nodes * FP_SN;
    f64* snn = sn0 + (size_t)(b ^ 1) * max_nodes * FP_SN;
    unsigned short* scl = scl0 + b * 4 * FP_PE;
    unsigned short* scs = scs0 + b * (max_nodes + 2);
    // ---- the next patch: its node records start travelling now (ids already in registers); lists and connectivity too ----
    ushort4 ln0n = make_ushort4(0xffffu, 0, 0, 0), ln1n = ln0n;
    unsigned short csn[2] = {0, 0};
    if (pn < n_patch) {
      if (tid < hn.y) stage_node_async(snn + (size_t)tid * FP_SN, (size_t)nid[0], N, xg, wg, dwg);
      if (tid + 128 < hn.y) stage_node_async(snn + (size_t)(tid + 128) * FP_SN, (size_t)nid[1], N, xg, wg, dwg);
      for (int k = tid + 256; k < hn.y; k += 128)   // (patches with more than 256 nodes: the rest is fetched here)
        stage_node_async(snn + (size_t)k * FP_SN, (size_t)__ldg(pnodes + hn.x + k), N, xg, wg, dwg);
      cp_async16(scl0 + (b ^ 1) * 4 * FP_PE + 8 * tid, corner + (size_t)pn * 4 * FP_PE + 8 * tid);
      ln0n = lnode[(size_t)pn * FP_PE + tid];
      ln1n = lnode[(size_t)pn * FP_PE + 128 + tid];
      const unsigned short* cs = cstart + (size_t)hn.x + pn;
      if (tid <= hn.y) csn[0] = cs[tid];
      if (tid + 128 <= hn.y) csn[1] = cs[tid + 128];
    }
    // node ids of the patch after the next one: two registers, needed one iteration from now
    const int pnn = pn + stride;
    int2 hnn = make_int2(0, 0);
    int nidn[2] = {0, 0};
    if (pnn < n_patch) {
      hnn = __ldg(hdr + pnn);
      if (tid < hnn.y) nidn[0] = __ldg(pnodes + hnn.x + tid);
      if (tid + 128 < hnn.y) nidn[1] = __ldg(pnodes + hnn.x + tid + 128);
    }
    // ---- B: two elements per thread out of the staged records ----
#pragma unroll 1
    for (int j = 0; j < 2; j++) {
      const ushort4 ln = j ? ln1 : ln0;
      if (ln.x == 0xffffu) continue;
      const f64* n0 = sn + (size_t)ln.x * FP_SN;
      const f64* n1 = sn + (size_t)ln.y * FP_SN;
      const f64* n2 = sn + (size_t)ln.z * FP_SN;
      const f64* n3 = sn + (size_t)ln.w * FP_SN;
      f64 x[4][3];
#pragma unroll
      for (int d = 0; d < 3; d++) { x[0][d] = n0[d]; x[1][d] = n1[d]; x[2][d] = n2[d]; x[3][d] = n3[d]; }
      Geom g;
      geometry(x, g);
      residual_rec(g, n0, n1, n2, n3, se + (size_t)(j * 128 + tid) * FP_SE);
    }
    __syncthreads();
    // ---- C: per patch-node partial sums, fixed corner order ----
    for (int k = tid; k < h.y; k += 128) {
      const int c0 = scs[k], c1 = scs[k + 1];
      f64 sacc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (int i = c0; i < c1; i++) {
        const int cr = scl[i];
        const f64* src = se + (cr >> 2) * FP_SE + (cr & 3) * 6;
#pragma unroll
        for (int c = 0; c < 6; c++) sacc[c] += src[c];
      }
      f64* dst = part + (size_t)(h.x + k) * 6;
#pragma unroll
      for (int c = 0; c < 6; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(sacc[c], sacc[c + 1]);
    }
    // ---- rotate: the next patch's ranges into shared memory, its records have landed ----
    if (pn < n_patch) {
      unsigned short* scsn = scs0 + (b ^ 1) * (max_nodes + 2);
      if (tid <= hn.y) scsn[tid] = csn[0];
      if (tid + 128 <= hn.y) scsn[tid + 128] = csn[1];
      const unsigned short* cs = cstart + (size_t)hn.x + pn;
      for (int k = tid + 256; k <= hn.y; k += 128) scsn[k] = cs[k];
    }
    cp_async_wait_all();
    __syncthreads();
    p = pn; h = hn; ln0 = ln0n; ln1 = ln1n;
    pn = pnn; hn = hnn; nid[0] = nidn[0]; nid[1] = nidn[1];
  }
}

// F[node] (+)= sum of the node's patch partials, ascending patch order
__global__ void k_gatherF2(int N, int n_rows, const int* __restrict__ np_ptr, const int* __restrict__ np, const f64* __restrict__ part,
                           f64* __restrict__ F, int overwrite) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  f64 s[6] = {0, 0, 0, 0, 0, 0};
  for (int j = np_ptr[i]; j < np_ptr[i + 1]; j++) {
    const double2* src = reinterpret_cast<const double2*>(part + (size_t)np[j] * 6);
    const double2 a = src[0], b = src[1], c = src[2];
    s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y;
  }
  f64* fu = F + (size_t)i * 3;
  if (overwrite) {
    fu[0] = s[0]; fu[1] = s[1]; fu[2] = s[2];
    F[(size_t)3 * N + i] = s[3]; F[(size_t)4 * N + i] = s[4]; F[(size_t)5 * N + i] = s[5];
  } else {
    fu[0] += s[0]; fu[1] += s[1]; fu[2] += s[2];
    F[(size_t)3 * N + i] += s[3]; F[(size_t)4 * N + i] += s[4]; F[(size_t)5 * N + i] += s[5];
  }
}

// ------------------------------------------------------------------------------------------------------------
// J: per-corner work shared by all variants
// ------------------------------------------------------------------------------------------------------------
struct CornerCtx {
  Geom g;
  JPrep p;
  ARow r;
  int row;  // global node of the corner
};

__device__ __forceinline__ void corner_setup(int corner, const int* __restrict__ ien, const f64* __restrict__ xg,
                                             const f64* __restrict__ wg, CornerCtx& c) {
  const int e = corner >> 2, a = corner & 3;
  int nodes[4];
  load_nodes(ien, e, nodes);
  f64 x[4][3], u[4][3];
  load_xyz(xg, nodes, x);
  load_xyz(wg, nodes, u);
  geometry(x, c.g);
  jac_prep(c.g, u, c.p);
  extract_row(c.g, c.p, a, c.r);
  c.row = a == 0 ? nodes[0] : (a == 1 ? nodes[1] : (a == 2 ? nodes[2] : nodes[3]));
}

// ATOMIC / COLORED: one thread per corner, scatter through the slot map
template <int OP>
__global__ void __launch_bounds__(128) k_cornerJ(int n_elem, const int* __restrict__ elem_ids, const int* __restrict__ ien,
                                                 const f64* __restrict__ xg, const f64* __restrict__ wg,
                                                 const int* __restrict__ row_ptr, const u32* __restrict__ slot32,
                                                 f64* __restrict__ A00, f64* __restrict__ A01, f64* __restrict__ A10,
                                                 f64* __restrict__ A11) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * n_elem) return;
  int e = elem_ids ? elem_ids[t >> 2] : (t >> 2);
  int corner = e * 4 + (t & 3);
  CornerCtx c;
  corner_setup(corner, ien, xg, wg, c);
  const int start = row_ptr[c.row], len = row_ptr[c.row + 1] - start;
  const u32 slots = slot32[corner];
#pragma unroll
  for (int b = 0; b < 4; b++) {
    f64 blk[16];
    jac_block_row(c.g, c.p, c.r, b, blk);
    scatter_block<OP>(A00, A01, A10, A11, (size_t)start, len, (int)((slots >> (8 * b)) & 0xffu), blk);
  }
}

// GATHER: one warp per nodal row
template <int NSG>  // slot groups of 16 (row length <= 16*NSG)
__global__ void __launch_bounds__(128) k_rowJ(int N, const int* __restrict__ ien, const f64* __restrict__ xg,
                                              const f64* __restrict__ wg, const int* __restrict__ row_ptr,
                                              const int* __restrict__ v2c_ptr, const int* __restrict__ v2c,
                                              const u32* __restrict__ slot32, f64* __restrict__ A00,
                                              f64* __restrict__ A01, f64* __restrict__ A10, f64* __restrict__ A11,
                                              int overwrite) {
  __shared__ f64 stage_s[4][16 * 32];
  __shared__ u32 smask_s[4][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= N) return;
  f64* stage = stage_s[warp];
  u32* smask = smask_s[warp];
  smask[lane] = 0u;
  smask[lane + 32] = 0u;
  const int cs = v2c_ptr[row], ce = v2c_ptr[row + 1];
  const int start = row_ptr[row], len = row_ptr[row + 1] - start;
  f64 acc[NSG][8];
#pragma unroll
  for (int sg = 0; sg < NSG; sg++)
#pragma unroll
    for (int v = 0; v < 8; v++) acc[sg][v] = 0.0;
  const int myslot = lane >> 1, half = lane & 1;
  __syncwarp();
  for (int base = cs; base < ce; base += 32) {
    const bool active = base + lane < ce;
    CornerCtx c;
    u32 slots = 0xffffffffu;
    if (active) {
      const int corner = v2c[base + lane];
      corner_setup(corner, ien, xg, wg, c);
      slots = slot32[corner];
    }
#pragma unroll
    for (int b = 0; b < 4; b++) {
      f64 blk[16];
      if (active) jac_block_row(c.g, c.p, c.r, b, blk);
      const u32 tgt = active ? ((slots >> (8 * b)) & 0xffu) : 255u;
      const u32 peers = __match_any_sync(FULL, tgt);
      const bool leader = active && ((__ffs(peers) - 1) == lane);
      if (leader) smask[tgt] = peers;
      if (active) {
#pragma unroll
        for (int v = 0; v < 16; v++) stage[v * 32 + lane] = blk[v];
      }
      __syncwarp();
#pragma unroll
      for (int sg = 0; sg < NSG; sg++) {
        const int s = sg * 16 + myslot;
        if (s < len) {
          u32 m = smask[s];
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
#pragma unroll
            for (int v = 0; v < 8; v++) acc[sg][v] += stage[(half * 8 + v) * 32 + src];
          }
        }
      }
      __syncwarp();
      if (leader) smask[tgt] = 0u;
      __syncwarp();
    }
  }
  // write-out: lane (slot, half) owns rows ii = 2*half, 2*half+1 of the 4x4 block
#pragma unroll
  for (int sg = 0; sg < NSG; sg++) {
    const int s = sg * 16 + myslot;
    if (s >= len) continue;
    const size_t st = (size_t)start;
    f64* p00 = A00 + st * 9 + (size_t)s * 3;
    f64* p01 = A01 + st * 3 + s;
    f64* p10 = A10 + st * 3 + (size_t)s * 3;
    f64* p11 = A11 + st + s;
    if (half == 0) {
      if (overwrite) {
#pragma unroll
        for (int ii = 0; ii < 2; ii++) {
#pragma unroll
          for (int jj = 0; jj < 3; jj++) p00[(size_t)ii * len * 3 + jj] = acc[sg][ii * 4 + jj];
          p01[(size_t)ii * len] = acc[sg][ii * 4 + 3];
        }
      } else {
#pragma unroll
        for (int ii = 0; ii < 2; ii++) {
#pragma unroll
          for (int jj = 0; jj < 3; jj++) p00[(size_t)ii * len * 3 + jj] += acc[sg][ii * 4 + jj];
          p01[(size_t)ii * len] += acc[sg][ii * 4 + 3];
        }
      }
    } else {
      if (overwrite) {
#pragma unroll
        for (int jj = 0; jj < 3; jj++) { p00[(size_t)2 * len * 3 + jj] = acc[sg][jj]; p10[jj] = acc[sg][4 + jj]; }
        p01[(size_t)2 * len] = acc[sg][3];
        p11[0] = acc[sg][7];
      } else {
#pragma unroll
        for (int jj = 0; jj < 3; jj++) { p00[(size_t)2 * len * 3 + jj] += acc[sg][jj]; p10[jj] += acc[sg][4 + jj]; }
        p01[(size_t)2 * len] += acc[sg][3];
        p11[0] += acc[sg][7];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// J, PULL (default).  Phase 1 (k_jprep2) evaluates the per-element part of the hoisted Jacobian once per element and
// parks it in a 384-byte record (three 128-byte lines).  Phase 2 (k_pullJ) runs one thread per WORK ITEM of the plan
// (plan.cuh): an off-diagonal nodal nonzero, or one of the four virtual items of a diagonal entry.  The thread walks the
// (element, a, b) contributions of its item in fixed order, forms each 4x4 block from the two corner sub-records and the
// element tail (13 16-byte loads, ~75 fp64 instructions) and accumulates the 16 values in registers; the four diagonal
// items of a row then fold their partial sums with a two-step butterfly inside their lane quad.  Every CSR value is
// written exactly once: no atomics, no colors, no memset, no shared-memory staging, deterministic.
// Record (doubles): corner x at [10x, 10x+10): g0 g1 | g2 P | c0 c1 | c2 c3 | R pad   (c_q = u(q).grad N_x)
//                   tail at [40,48): w sTM | sTC pad | tM0 tM1 | tM2 tM3
// ------------------------------------------------------------------------------------------------------------
constexpr int PREC = 48;
constexpr int PULL_SREC = 25;   // shared-memory stride of a staged record in 16-byte chunks (odd: bank spreading)

// The 32 records of a warp are contiguous in global memory (32 x 384 B): each lane parks its record in shared memory
// (stride 49 doubles: conflict-free), then the warp streams the 12 KB out with fully coalesced 8-byte stores.
constexpr int PREC_S = 49;
__global__ void __launch_bounds__(128) k_jprep2(int E, const int* __restrict__ ien, const f64* __restrict__ xg,
                                                const f64* __restrict__ wg, f64* __restrict__ rec) {
  extern __shared__ f64 stage_rec[];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  f64* mine = stage_rec + (size_t)(warp * 32 + lane) * PREC_S;
  if (e < E) {
    int nodes[4];
    load_nodes(ien, e, nodes);
    f64 x[4][3], u[4][3];
    load_xyz(xg, nodes, x);
    load_xyz(wg, nodes, u);
    Geom g;
    geometry(x, g);
    JPrep p;
    jac_prep(g, u, p);
#pragma unroll
    for (int a = 0; a < 4; a++) {
      mine[a * 10 + 0] = g.sh[a][0]; mine[a * 10 + 1] = g.sh[a][1]; mine[a * 10 + 2] = g.sh[a][2]; mine[a * 10 + 3] = p.P[a];
      mine[a * 10 + 4] = p.c[0][a]; mine[a * 10 + 5] = p.c[1][a]; mine[a * 10 + 6] = p.c[2][a]; mine[a * 10 + 7] = p.c[3][a];
      mine[a * 10 + 8] = p.R[a]; mine[a * 10 + 9] = 0.0;
    }
    mine[40] = p.w; mine[41] = p.sTM; mine[42] = p.sTC; mine[43] = 0.0;
    mine[44] = p.tM[0]; mine[45] = p.tM[1]; mine[46] = p.tM[2]; mine[47] = p.tM[3];
  }
  __syncwarp();
  const int e0 = (blockIdx.x * blockDim.x) + warp * 32;           // first element of this warp
  const int nrec = min(32, E - e0);
  if (nrec <= 0) return;
  f64* dst = rec + (size_t)e0 * PREC;
  const f64* src = stage_rec + (size_t)warp * 32 * PREC_S;
  const int total = nrec * PREC;
  for (int t = lane; t < total; t += 32) {
    const int r = t / PREC, o = t - r * PREC;
    dst[t] = src[r * PREC_S + o];
  }
}

// one (element, a, b) contribution from the two corner sub-records and the element tail (record layout above), added to acc
__device__ __forceinline__ void pull_block(const double2 a0, const double2 a1, const double2 a2, const double2 a3, const double2 b0,
                                           const double2 b1, const double2 b2, const double2 b3, const double2 b4, const double2 t0,
                                           const double2 t1, const double2 t2, const double2 t3, int a, int b, f64 acc[16]) {
  const f64 ga[3] = {a0.x, a0.y, a1.x}, gb[3] = {b0.x, b0.y, b1.x};
  const f64 Pa = a1.y, Pb = b1.y, Rb = b4.x;
  const f64 w = t0.x, sTM = t0.y, sTC = t1.x;
  const f64 tMb = sel4(b, t2.x, t2.y, t3.x, t3.y);
  const f64 cab = sel4(a, b2.x, b2.y, b3.x, b3.y);          // c[q=a][b]
  const f64 cba = sel4(b, a2.x, a2.y, a3.x, a3.y);          // c[q=b][a]
  const f64 stc = (t2.x * a2.x) * b2.x + (t2.y * a2.y) * b2.y + (t3.x * a3.x) * b3.x + (t3.y * a3.y) * b3.y;
  const f64 eK = ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2];
  const f64 mab = (a == b) ? (SA * SA + 3.0 * SB * SB) : (2.0 * SA * SB + 2.0 * SB * SB);
  const f64 T = w * (FACT1 * RHO * mab + FACT1 * RHO * RHO * (SB * Pa + SD * (tMb * cba)) + FACT2 * RHO * (SB * Rb + SD * cab) +
                     FACT2 * RHO * RHO * stc + 4.0 * FACT2 * MU * eK);
  const f64 k1 = 4.0 * w * FACT2 * MU, k2 = w * FACT2 * RHO * sTC;
#pragma unroll
  for (int ii = 0; ii < 3; ii++)
#pragma unroll
    for (int jj = 0; jj < 3; jj++) acc[ii * 4 + jj] += k1 * ga[jj] * gb[ii] + k2 * ga[ii] * gb[jj] + (ii == jj ? T : 0.0);
  const f64 k3 = w * SN, k4 = RHO * w * Pa;
  const f64 k5 = w * RHO * (FACT1 * (SB * sTM + SD * tMb) + FACT2 * Pb);
  const f64 k6 = FACT2 * w * SN;
#pragma unroll
  for (int ii = 0; ii < 3; ii++) {
    acc[ii * 4 + 3] += -k3 * ga[ii] + k4 * gb[ii];
    acc[12 + ii] += k5 * ga[ii] + k6 * gb[ii];
  }
  acc[15] += w * sTM * eK;
}

// Fold the four diagonal items of a row (an aligned lane quad) -- after two exchange steps lane q of the quad holds block
// row q -- and write the item's values.  Executed by every lane of the warp (no divergence before the shuffles); only
// diagonal items use the folded result.
template <int OVERWRITE>
__device__ __forceinline__ void pull_finish(const f64 acc[16], bool valid, uint2 meta, const int* __restrict__ row_ptr,
                                            f64* __restrict__ A00, f64* __restrict__ A01, f64* __restrict__ A10,
                                            f64* __restrict__ A11) {
  const int lane = threadIdx.x & 31;
  const bool hi = lane & 2, odd = lane & 1;
  f64 t[2][4], u[4];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const f64 send = hi ? acc[r * 4 + j] : acc[(r + 2) * 4 + j];
      const f64 keep = hi ? acc[(r + 2) * 4 + j] : acc[r * 4 + j];
      t[r][j] = keep + __shfl_xor_sync(FULL, send, 2);
    }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const f64 send = odd ? t[0][j] : t[1][j];
    const f64 keep = odd ? t[1][j] : t[0][j];
    u[j] = keep + __shfl_xor_sync(FULL, send, 1);
  }
  if (!valid || meta.x == 0xffffffffu) return;
  const int row = (int)meta.x, k = (int)(meta.y & 0xffu);
  const int start = __ldg(row_ptr + row), len = __ldg(row_ptr + row + 1) - start;
  const size_t st = (size_t)start;
  f64* p00 = A00 + st * 9 + (size_t)k * 3;
  f64* p01 = A01 + st * 3 + k;
  f64* p10 = A10 + st * 3 + (size_t)k * 3;
  f64* p11 = A11 + st + k;
  if (meta.y & 0x100u) {
    const int q = lane & 3;
    if (q < 3) {
      f64* d = p00 + (size_t)q * len * 3;
      f64* d1 = p01 + (size_t)q * len;
      if (OVERWRITE) { d[0] = u[0]; d[1] = u[1]; d[2] = u[2]; *d1 = u[3]; }
      else { d[0] += u[0]; d[1] += u[1]; d[2] += u[2]; *d1 += u[3]; }
    } else {
      if (OVERWRITE) { p10[0] = u[0]; p10[1] = u[1]; p10[2] = u[2]; *p11 = u[3]; }
      else { p10[0] += u[0]; p10[1] += u[1]; p10[2] += u[2]; *p11 += u[3]; }
    }
  } else {
#pragma unroll
    for (int ii = 0; ii < 3; ii++) {
      f64* d = p00 + (size_t)ii * len * 3;
      f64* d1 = p01 + (size_t)ii * len;
      if (OVERWRITE) { d[0] = acc[ii * 4]; d[1] = acc[ii * 4 + 1]; d[2] = acc[ii * 4 + 2]; *d1 = acc[ii * 4 + 3]; }
      else { d[0] += acc[ii * 4]; d[1] += acc[ii * 4 + 1]; d[2] += acc[ii * 4 + 2]; *d1 += acc[ii * 4 + 3]; }
    }
    if (OVERWRITE) { p10[0] = acc[12]; p10[1] = acc[13]; p10[2] = acc[14]; *p11 = acc[15]; }
    else { p10[0] += acc[12]; p10[1] += acc[13]; p10[2] += acc[14]; *p11 += acc[15]; }
  }
}

template <int OVERWRITE>
__global__ void __launch_bounds__(128) k_pullJ(int n_items, const uint2* __restrict__ item_meta, const int* __restrict__ item_ptr,
                                               const u32* __restrict__ contrib, const f64* __restrict__ rec,
                                               const int* __restrict__ row_ptr, f64* __restrict__ A00, f64* __restrict__ A01,
                                               f64* __restrict__ A10, f64* __restrict__ A11) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = item < n_items;
  uint2 meta = make_uint2(0xffffffffu, 0u);
  int cs = 0, ce = 0;
  if (valid) {
    meta = item_meta[item];
    cs = __ldg(item_ptr + item);
    ce = __ldg(item_ptr + item + 1);
  }
  f64 acc[16];
#pragma unroll
  for (int v = 0; v < 16; v++) acc[v] = 0.0;
  for (int idx = cs; idx < ce; idx++) {
    const u32 cid = __ldg(contrib + idx);
    const int a = (int)((cid >> 2) & 3u), b = (int)(cid & 3u);
    const double2* R = reinterpret_cast<const double2*>(rec + (size_t)(cid >> 4) * PREC);
    const double2 a0 = R[a * 5], a1 = R[a * 5 + 1], a2 = R[a * 5 + 2], a3 = R[a * 5 + 3];
    const double2 b0 = R[b * 5], b1 = R[b * 5 + 1], b2 = R[b * 5 + 2], b3 = R[b * 5 + 3], b4 = R[b * 5 + 4];
    const double2 t0 = R[20], t1 = R[21], t2 = R[22], t3 = R[23];
    pull_block(a0, a1, a2, a3, b0, b1, b2, b3, b4, t0, t1, t2, t3, a, b, acc);
  }
  pull_finish<OVERWRITE>(acc, valid, meta, row_ptr, A00, A01, A10, A11);
}

// STAGED pull (default): one CTA per group of PULL_ROWS rows.  The group's distinct element records (plan: cta_elems) are
// first copied into shared memory with coalesced 16-byte loads -- a record is 24 16-byte chunks at a stride of 25, so that
// lanes reading the same chunk of different records spread over the banks -- and the work items of the group then pull their contributions from there (the plain k_pullJ above is L1-tag bound: every lane of a load touches
// a different record line).  Groups whose records do not fit (contrib16 == 0xffff) read global memory like k_pullJ.
template <int OVERWRITE>
__global__ void __launch_bounds__(128) k_pullJ_staged(int N, int n_rows, const int* __restrict__ row_item,
                                                      const uint2* __restrict__ item_meta, const int* __restrict__ item_ptr,
                                                      const u32* __restrict__ contrib, const unsigned short* __restrict__ contrib16,
                                                      const int* __restrict__ cta_elem_ptr, const int* __restrict__ cta_elems,
                                                      const f64* __restrict__ rec, const int* __restrict__ row_ptr,
                                                      f64* __restrict__ A00, f64* __restrict__ A01, f64* __restrict__ A10,
                                                      f64* __restrict__ A11) {
  extern __shared__ __align__(16) unsigned char pull_smem[];
  double2* srec = reinterpret_cast<double2*>(pull_smem);
  const int r0 = blockIdx.x * PULL_ROWS, r1 = min(N, r0 + PULL_ROWS);
  const int it0 = __ldg(row_item + r0), it1 = __ldg(row_item + r1);
  const int e0 = __ldg(cta_elem_ptr + blockIdx.x), ne = __ldg(cta_elem_ptr + blockIdx.x + 1) - e0;
  const bool staged = ne > 0 && ne <= PULL_MAX_STAGED;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (staged) {
    for (int r = warp; r < ne; r += 4) {
      const int e = __ldg(cta_elems + e0 + r);
      if (lane < 24) srec[r * PULL_SREC + lane] = __ldg(reinterpret_cast<const double2*>(rec + (size_t)e * PREC) + lane);
    }
  }
  __syncthreads();
  for (int base = it0; base < it1; base += 128) {   // block-uniform trip count (the fold below shuffles)
    const int item = base + threadIdx.x;
    const bool valid = item < it1;
    uint2 meta = make_uint2(0xffffffffu, 0u);
    int cs = 0, ce = 0;
    if (valid) {
      meta = item_meta[item];
      cs = __ldg(item_ptr + item);
      ce = __ldg(item_ptr + item + 1);
    }
    f64 acc[16];
#pragma unroll
    for (int v = 0; v < 16; v++) acc[v] = 0.0;
    if (staged) {
      unsigned nxt = cs < ce ? contrib16[cs] : 0u;   // one-ahead prefetch: the list load is off the LDS -> FP64 critical path
      for (int idx = cs; idx < ce; idx++) {
        const unsigned c16 = nxt;
        if (idx + 1 < ce) nxt = contrib16[idx + 1];
        const int li = (int)(c16 >> 4), a = (int)((c16 >> 2) & 3u), b = (int)(c16 & 3u);
        const double2* R = srec + li * PULL_SREC;
        const double2* Ra = R + a * 5;
        const double2* Rb = R + b * 5;
        pull_block(Ra[0], Ra[1], Ra[2], Ra[3], Rb[0], Rb[1], Rb[2], Rb[3], Rb[4], R[20], R[21], R[22], R[23], a, b, acc);
      }
    } else {
      for (int idx = cs; idx < ce; idx++) {
        const u32 cid = __ldg(contrib + idx);
        const int a = (int)((cid >> 2) & 3u), b = (int)(cid & 3u);
        const double2* R = reinterpret_cast<const double2*>(rec + (size_t)(cid >> 4) * PREC);
        pull_block(R[a * 5], R[a * 5 + 1], R[a * 5 + 2], R[a * 5 + 3], R[b * 5], R[b * 5 + 1], R[b * 5 + 2], R[b * 5 + 3], R[b * 5 + 4],
                   R[20], R[21], R[22], R[23], a, b, acc);
      }
    }
    pull_finish<OVERWRITE>(acc, valid && (meta.x == 0xffffffffu || (int)meta.x < n_rows), meta, row_ptr, A00, A01, A10, A11);
  }
}

// ------------------------------------------------------------------------------------------------------------
// J, PAIRS (default).  One CTA per group of R rows (plan.cuh pr_*), two phases, no global intermediate:
//   A. every thread evaluates the per-element part of the hoisted Jacobian (geometry + jac_prep) of one of the group's
//      distinct elements and parks the 368-byte record in shared memory (23 16-byte chunks: odd stride, conflict-free);
//   B. one thread per work item.  The first 4R items are the diagonal entries (4 virtual items each, whole warps, folded by
//      the quad butterfly); the others are the UPPER off-diagonal nonzeros (i,j), j > i: the thread walks the elements
//      around the edge once and accumulates BOTH blocks A_ij and A_ji (elem_math.cuh jrec_pair: 13 LDS.128 and ~90 FP64
//      instructions per element instead of 2 x (13 + ~80)), then writes A_ij into row i (coalesced across the lanes of a
//      row) and A_ji into row j (scattered 24-byte pieces that L2 merges with the rest of row j).
// Every CSR value is written exactly once, in a fixed order: no atomics, no colors, no memset, deterministic.
// ------------------------------------------------------------------------------------------------------------
constexpr int PAIR_SREC = 23;   // shared-memory stride of a staged record in 16-byte chunks (JREC = 46 doubles)

__device__ __forceinline__ void fold_quad(const f64 acc[16], f64 u[4]) {
  const int lane = threadIdx.x & 31;
  const bool hi = lane & 2, odd = lane & 1;
  f64 t[2][4];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const f64 send = hi ? acc[r * 4 + j] : acc[(r + 2) * 4 + j];
      const f64 keep = hi ? acc[(r + 2) * 4 + j] : acc[r * 4 + j];
      t[r][j] = keep + __shfl_xor_sync(FULL, send, 2);
    }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const f64 send = odd ? t[0][j] : t[1][j];
    const f64 keep = odd ? t[1][j] : t[0][j];
    u[j] = keep + __shfl_xor_sync(FULL, send, 1);
  }
}

__device__ __forceinline__ void load_chunks5(const double2* __restrict__ p, f64 out[10]) {
#pragma unroll
  for (int c = 0; c < 5; c++) { const double2 v = p[c]; out[2 * c] = v.x; out[2 * c + 1] = v.y; }
}

__device__ __forceinline__ void pair_prep_element(const int4 nd, const f64* __restrict__ xg, const f64* __restrict__ wg,
                                                  double2* __restrict__ dst) {
  const int nodes[4] = {nd.x, nd.y, nd.z, nd.w};
  f64 x[4][3], u[4][3];
  load_xyz(xg, nodes, x);
  load_xyz(wg, nodes, u);
  Geom gm;
  geometry(x, gm);
  JPrep p;
  jac_prep(gm, u, p);
  f64 rec[JREC];
  jrec_store(gm, p, rec);
#pragma unroll
  for (int c = 0; c < PAIR_SREC; c++) dst[c] = make_double2(rec[2 * c], rec[2 * c + 1]);
}

template <int OVERWRITE, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_pairJ(int N, int n_rows, int R, const int4* __restrict__ grp,
                                                  const uint2* __restrict__ item_meta, const int* __restrict__ item_ptr,
                                                  const unsigned short* __restrict__ contrib, const int4* __restrict__ enodes,
                                                  const f64* __restrict__ xg, const f64* __restrict__ wg,
                                                  const int* __restrict__ row_ptr, const int* __restrict__ col_ind,
                                                  f64* __restrict__ A00, f64* __restrict__ A01, f64* __restrict__ A10,
                                                  f64* __restrict__ A11) {
  extern __shared__ __align__(16) unsigned char pair_smem[];
  double2* srec = reinterpret_cast<double2*>(pair_smem);
  const int4 gd = __ldg(grp + blockIdx.x);
  const int e0 = gd.x, ne = gd.y, it0 = gd.z, n_it = gd.w;
  const int nd = 4 * R;
  // The dependent-load chains of both phases start together: group -> {element nodes, item meta / list bounds} ->
  // {coordinates + velocities, first list entry}.  (With ~12 resident warps per SM these chains, not the arithmetic, set the pace.)
  uint2 meta0 = make_uint2(0xffffffffu, 0u);
  int cs0 = 0, ce0 = 0;
  if ((int)threadIdx.x < n_it) {
    meta0 = item_meta[it0 + threadIdx.x];
    cs0 = __ldg(item_ptr + it0 + threadIdx.x);
    ce0 = __ldg(item_ptr + it0 + threadIdx.x + 1);
  }
  // ---- phase A: element records into shared memory (node ids fetched one trip ahead) ----
  int4 nd_cur = make_int4(0, 0, 0, 0);
  if ((int)threadIdx.x < ne) nd_cur = __ldg(enodes + e0 + threadIdx.x);
  unsigned first0 = 0u;
  for (int r = threadIdx.x; r < ne; r += blockDim.x) {
    int4 nd_next = nd_cur;
    if (r + (int)blockDim.x < ne) nd_next = __ldg(enodes + e0 + r + blockDim.x);
    pair_prep_element(nd_cur, xg, wg, srec + r * PAIR_SREC);
    nd_cur = nd_next;
  }
  if (cs0 < ce0) first0 = contrib[cs0];
  __syncthreads();
  // ---- phase B: work items ----
  for (int base = 0; base < n_it; base += blockDim.x) {
    const int t = base + threadIdx.x;
    const int item = it0 + t;
    uint2 meta = meta0;
    int cs = cs0, ce = ce0;
    unsigned nxt = first0;
    if (base > 0) {
      meta = make_uint2(0xffffffffu, 0u);
      cs = ce = 0;
      if (t < n_it) {
        meta = item_meta[item];
        cs = __ldg(item_ptr + item);
        ce = __ldg(item_ptr + item + 1);
      }
      nxt = cs < ce ? contrib[cs] : 0u;
    }
    if (base + (int)(threadIdx.x & ~31u) < nd) {
      // a warp of diagonal items (nd is a multiple of 32 and n_it >= nd: every lane has an item)
      f64 acc[16];
#pragma unroll
      for (int v = 0; v < 16; v++) acc[v] = 0.0;
      for (int idx = cs; idx < ce; idx++) {
        const unsigned c16 = nxt;
        if (idx + 1 < ce) nxt = contrib[idx + 1];
        const int li = (int)(c16 >> 4), a = (int)((c16 >> 2) & 3u);
        const double2* Rr = srec + li * PAIR_SREC;
        f64 A[10], T[6];
        load_chunks5(Rr + a * 5, A);
        const double2 t0 = Rr[20], t1 = Rr[21], t2 = Rr[22];
        T[0] = t0.x; T[1] = t0.y; T[2] = t1.x; T[3] = t1.y; T[4] = t2.x; T[5] = t2.y;
        jrec_diag(A, T, a, acc);
      }
      f64 u[4];
      fold_quad(acc, u);
      if (meta.x != 0xffffffffu && (int)meta.x < n_rows) {
        const int row = (int)meta.x, k = (int)(meta.y & 0xffu);
        const int start = __ldg(row_ptr + row), len = __ldg(row_ptr + row + 1) - start;
        const size_t st = (size_t)start;
        const int q = threadIdx.x & 3;
        if (q < 3) {
          f64* d = A00 + st * 9 + (size_t)k * 3 + (size_t)q * len * 3;
          f64* d1 = A01 + st * 3 + k + (size_t)q * len;
          if (OVERWRITE) { d[0] = u[0]; d[1] = u[1]; d[2] = u[2]; *d1 = u[3]; }
          else { d[0] += u[0]; d[1] += u[1]; d[2] += u[2]; *d1 += u[3]; }
        } else {
          f64* p10 = A10 + st * 3 + (size_t)k * 3;
          f64* p11 = A11 + st + k;
          if (OVERWRITE) { p10[0] = u[0]; p10[1] = u[1]; p10[2] = u[2]; *p11 = u[3]; }
          else { p10[0] += u[0]; p10[1] += u[1]; p10[2] += u[2]; *p11 += u[3]; }
        }
      }
    } else if (t < n_it) {
      // row pointers of both rows are fetched before the accumulation loop: their latency hides behind it
      const int row = (int)meta.x;
      const int kij = (int)(meta.y & 0xffu), kji = (int)((meta.y >> 16) & 0xffu);
      const int start = __ldg(row_ptr + row), len = __ldg(row_ptr + row + 1) - start;
      const int j = __ldg(col_ind + start + kij);
      const int sj = __ldg(row_ptr + j), lj = __ldg(row_ptr + j + 1) - sj;
      f64 acc[24];
#pragma unroll
      for (int v = 0; v < 24; v++) acc[v] = 0.0;
      for (int idx = cs; idx < ce; idx++) {
        const unsigned c16 = nxt;
        if (idx + 1 < ce) nxt = contrib[idx + 1];
        const int li = (int)(c16 >> 4), a = (int)((c16 >> 2) & 3u), b = (int)(c16 & 3u);
        const double2* Rr = srec + li * PAIR_SREC;
        f64 A[10], B[10], T[6];
        load_chunks5(Rr + a * 5, A);
        load_chunks5(Rr + b * 5, B);
        const double2 t0 = Rr[20], t1 = Rr[21], t2 = Rr[22];
        T[0] = t0.x; T[1] = t0.y; T[2] = t1.x; T[3] = t1.y; T[4] = t2.x; T[5] = t2.y;
        jrec_pair(A, B, T, a, b, acc);
      }
      if (row < n_rows) {
        f64 ab[16], ba[16];
        jrec_pair_blocks(acc, ab, ba);
        scatter_block<OVERWRITE ? 2 : 0>(A00, A01, A10, A11, (size_t)start, len, kij, ab);
        if (j < n_rows) scatter_block<OVERWRITE ? 2 : 0>(A00, A01, A10, A11, (size_t)sj, lj, kji, ba);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// boundary faces (tiny: O(N^(2/3)) faces): one thread per face, atomic scatter
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_face(int nf, const int* __restrict__ f2e, const int* __restrict__ forn, int N,
                                              const int* __restrict__ ien, const f64* __restrict__ xg,
                                              const f64* __restrict__ wg, const f64* __restrict__ dwg,
                                              const int* __restrict__ row_ptr, const u32* __restrict__ slot32,
                                              f64* __restrict__ F, f64* __restrict__ A00, f64* __restrict__ A01,
                                              f64* __restrict__ A10, f64* __restrict__ A11) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const int e = f2e[f], iorn = forn[f];
  int nodes[4];
  load_nodes(ien, e, nodes);
  f64 x[4][3], u[4][3];
  load_xyz(xg, nodes, x);
  load_xyz(wg, nodes, u);
  Geom g;
  geometry(x, g);
  FacePrep fp;
  face_prep(g, iorn, fp);
  if (F) {
    f64 val[4][4], eF[4][6];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      val[0][a] = u[a][0]; val[1][a] = u[a][1]; val[2][a] = u[a][2];
      val[3][a] = __ldg(dwg + (size_t)3 * N + nodes[a]);
    }
    face_residual(g, fp, iorn, val, eF);
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const size_t nd = (size_t)nodes[a];
      atomicAdd(F + nd * 3 + 0, eF[a][0]);
      atomicAdd(F + nd * 3 + 1, eF[a][1]);
      atomicAdd(F + nd * 3 + 2, eF[a][2]);
      atomicAdd(F + (size_t)3 * N + nd, eF[a][3]);
    }
  }
  if (A00) {
#pragma unroll 1
    for (int a = 0; a < 4; a++) {
      const int row = nodes[a];
      const int start = row_ptr[row], len = row_ptr[row + 1] - start;
      const u32 slots = slot32[e * 4 + a];
#pragma unroll 1
      for (int b = 0; b < 4; b++) {
        f64 blk[16];
        face_block(g, fp, iorn, u, a, b, blk);
        scatter_block<1>(A00, A01, A10, A11, (size_t)start, len, (int)((slots >> (8 * b)) & 0xffu), blk);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Dirichlet
// ------------------------------------------------------------------------------------------------------------
__global__ void k_dirichlet_vec(int nb, const int* __restrict__ bnode, int shape, int mask, f64* __restrict__ b) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const size_t base = (size_t)bnode[i] * shape;
  for (int ic = 0; ic < shape; ic++)
    if (mask & (1 << ic)) b[base + ic] = 0.0;
}

// one thread per (boundary node, slot k): rows node*3+ic of A00 become unit rows, of A01 zero rows
__global__ void k_dirichlet_mat(int nb, const int* __restrict__ bnode, int mask, int N, const int* __restrict__ row_ptr,
                                const int* __restrict__ col_ind, f64* __restrict__ A00, f64* __restrict__ A01) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int i = t >> 6, k = t & 63;
  if (i >= nb) return;
  const int node = bnode[i];
  if (node < 0 || node >= N) return;
  const int start = row_ptr[node], len = row_ptr[node + 1] - start;
  if (k >= len) return;
  const bool diag = col_ind[start + k] == node;
#pragma unroll
  for (int ic = 0; ic < 3; ic++) {
    if (!(mask & (1 << ic))) continue;
    f64* p = A00 + (size_t)start * 9 + (size_t)ic * 3 * len + (size_t)k * 3;
    p[0] = (diag && ic == 0) ? 1.0 : 0.0;
    p[1] = (diag && ic == 1) ? 1.0 : 0.0;
    p[2] = (diag && ic == 2) ? 1.0 : 0.0;
    A01[(size_t)start * 3 + (size_t)ic * len + k] = 0.0;
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_assemble_tet(const dfb_plan* P, const double* d_xg, const double* d_wg, const double* d_dwg, double* d_F,
                     double* d_A00, double* d_A01, double* d_A10, double* d_A11, int mode, int overwrite, void* stream) {
  cudaStream_t st = as_stream(stream);
  NvtxRange nvtx(d_A00 ? (d_F ? "dfb_assemble_tet F+J" : "dfb_assemble_tet J") : "dfb_assemble_tet F");
  if (!P || !d_xg || !d_wg || !d_dwg) { set_error("dfb_assemble_tet: bad argument"); return DFB_ERR_ARG; }
  const bool doJ = d_A00 != nullptr;
  if (doJ && (!d_A01 || !d_A10 || !d_A11)) { set_error("dfb_assemble_tet: all four sub-block arrays are required"); return DFB_ERR_ARG; }
  if (doJ && !P->slot) { set_error("dfb_assemble_tet: the plan was created without a sparsity pattern (residual only)"); return DFB_ERR_ARG; }
  if (mode == DFB_MODE_AUTO) mode = DFB_MODE_GATHER;
  if (mode == DFB_MODE_COLORED && (P->num_batch <= 0 || !P->batch_ind)) { set_error("dfb_assemble_tet: plan has no color batches"); return DFB_ERR_ARG; }
  if (overwrite && mode != DFB_MODE_GATHER) { set_error("dfb_assemble_tet: overwrite needs DFB_MODE_GATHER"); return DFB_ERR_ARG; }
  const int N = P->N, E = P->E;
  const u32* slot32 = reinterpret_cast<const u32*>(P->slot);
  if (d_F) {
    bool patch_done = false;
    if (mode == DFB_MODE_GATHER && options().f_variant >= 1) {
      DFB_CHECK(build_fpatch(P, d_xg, st));
      if (P->fp_state == 1) {
        const int mn = (P->fp_max_nodes + 1) & ~1;   // even: keeps the 16-byte alignment of what follows the node records
        const size_t smem = sizeof(f64) * ((size_t)mn * FP_SN + (size_t)FP_PE * FP_SE) + sizeof(unsigned short) * ((size_t)4 * FP_PE + mn + 8);
        DFB_CHECK(ensure_dynamic_smem((const void*)k_patchF<2>, smem));
        DFB_CHECK(ensure_dynamic_smem((const void*)k_patchF<3>, smem));
        if (options().f_variant == 2) {
          const size_t smem2 = sizeof(f64) * ((size_t)2 * mn * FP_SN + (size_t)FP_PE * FP_SE) +
                               sizeof(unsigned short) * ((size_t)2 * 4 * FP_PE + 2 * (mn + 2) + 8);
          if (smem2 <= 113 * 1024) {   // two CTAs per SM
            DFB_CHECK(ensure_dynamic_smem((const void*)k_patchF_pipe, smem2));
            k_patchF_pipe<<<std::min(P->fp_n_patch, 2 * num_sms()), 128, smem2, st>>>(N, P->fp_n_patch, P->fp_hdr, P->fp_nodes, P->fp_lnode,
                                                                                     P->fp_corner, P->fp_cstart, d_xg, d_wg, d_dwg,
                                                                                     P->fp_part, mn);
          } else {
            k_patchF<2><<<P->fp_n_patch, 128, smem, st>>>(N, P->fp_hdr, P->fp_nodes, P->fp_lnode, P->fp_corner, P->fp_cstart, d_xg, d_wg,
                                                          d_dwg, P->fp_part, mn);
          }
        } else if (options().f_patch_ctas == 2)
          k_patchF<2><<<P->fp_n_patch, 128, smem, st>>>(N, P->fp_hdr, P->fp_nodes, P->fp_lnode, P->fp_corner, P->fp_cstart, d_xg, d_wg,
                                                        d_dwg, P->fp_part, mn);
        else
          k_patchF<3><<<P->fp_n_patch, 128, smem, st>>>(N, P->fp_hdr, P->fp_nodes, P->fp_lnode, P->fp_corner, P->fp_cstart, d_xg, d_wg,
                                                        d_dwg, P->fp_part, mn);
        DFB_LAUNCH_CHECK();
        k_gatherF2<<<ceil_div(P->n_rows, 128), 128, 0, st>>>(N, P->n_rows, P->fp_np_ptr, P->fp_np, P->fp_part, d_F, overwrite);
        DFB_LAUNCH_CHECK();
        patch_done = true;
      }
    }
    if (patch_done) {
    } else if (mode == DFB_MODE_GATHER) {
      if (!P->elemF) {
        P->elemF_bytes = sizeof(f64) * 24 * (size_t)E + sizeof(int) * 4 * (size_t)E;
        DFB_CUDA(cudaMalloc(&P->elemF, sizeof(f64) * 24 * (size_t)E));
        DFB_CUDA(cudaMalloc(&P->cpos, sizeof(int) * 4 * (size_t)E));
        k_corner_pos<<<ceil_div(4 * (i64)E, 256), 256, 0, st>>>(4 * E, P->v2c, P->cpos);
        DFB_LAUNCH_CHECK();
      }
      k_elemF<0><<<ceil_div(E, 128), 128, 0, st>>>(E, nullptr, N, P->ien, d_xg, d_wg, d_dwg, P->elemF, P->cpos);
      DFB_LAUNCH_CHECK();
      k_gatherF<<<ceil_div(P->n_rows, 128), 128, 0, st>>>(N, P->n_rows, P->v2c_ptr, P->elemF, d_F, overwrite);
      DFB_LAUNCH_CHECK();
    } else if (mode == DFB_MODE_ATOMIC) {
      k_elemF<1><<<ceil_div(E, 128), 128, 0, st>>>(E, nullptr, N, P->ien, d_xg, d_wg, d_dwg, d_F, nullptr);
      DFB_LAUNCH_CHECK();
    } else {
      for (int b = 0; b < P->num_batch; b++) {
        int n = P->batch_offset[b + 1] - P->batch_offset[b];
        if (n == 0) break;  // reference assemble.cu:1565-1567
        k_elemF<2><<<ceil_div(n, 128), 128, 0, st>>>(n, P->batch_ind + P->batch_offset[b], N, P->ien, d_xg, d_wg, d_dwg, d_F, nullptr);
        DFB_LAUNCH_CHECK();
      }
    }
  }
  if (doJ) {
    if (mode == DFB_MODE_GATHER) {
      const int grid = ceil_div(P->n_rows, 4);
      // variants of the atomic-free assembly: node pairs (default), pull (DFB_J_VARIANT=pull) and the fused row gather
      // (DFB_J_VARIANT=fused); see DESIGN.md section 3
      const int variant = options().j_variant;   // 0 pull, 1 fused, 2 pairs (dfb_set_option / DFB_J_VARIANT)
      bool done = false;
      if (variant == 2) {
        const int pair_rows = options().j_pair_rows;   // only looked at when the plan's pair lists are first built
        DFB_CHECK(build_pairs(P, pair_rows, d_xg, st));
        if (P->pr_state == 1) {
          const size_t smem = (size_t)std::max(1, P->pr_max_elems) * PAIR_SREC * sizeof(double2);
          const int R = P->pr_rows, ncta = P->pr_n_cta;
#define DFB_PAIR_LAUNCH(NT, MINB)                                                                                                  \
  do {                                                                                                                             \
    DFB_CHECK(ensure_dynamic_smem((const void*)k_pairJ<0, NT, MINB>, smem));                                                       \
    DFB_CHECK(ensure_dynamic_smem((const void*)k_pairJ<1, NT, MINB>, smem));                                                       \
    if (overwrite)                                                                                                                 \
      k_pairJ<1, NT, MINB><<<ncta, NT, smem, st>>>(N, P->n_rows, R, P->pr_grp, P->pr_meta, P->pr_item_ptr, P->pr_contrib,          \
                                                   P->pr_enodes, d_xg, d_wg, P->row_ptr, P->col_ind, d_A00, d_A01, d_A10, d_A11);  \
    else                                                                                                                           \
      k_pairJ<0, NT, MINB><<<ncta, NT, smem, st>>>(N, P->n_rows, R, P->pr_grp, P->pr_meta, P->pr_item_ptr, P->pr_contrib,          \
                                                   P->pr_enodes, d_xg, d_wg, P->row_ptr, P->col_ind, d_A00, d_A01, d_A10, d_A11);  \
  } while (0)
          // 96 threads (3 warps: one of diagonal items, two of pair items) and 4 CTAs per SM measured best on B200 at 1M tets:
          // 96x5 (128 registers, spills) 377 us, 128x3 432 us, 128x4 403 us, 160x2 547 us, 16 rows x 192 threads 404 us vs 372 us;
          // after the arithmetic trim: 64x5 and 64x6 370 us vs 362 us
          const int nt = options().j_pair_nt;
          if (R == 16) { if (nt == 224) DFB_PAIR_LAUNCH(224, 2); else DFB_PAIR_LAUNCH(192, 2); }
          else if (nt == 128) DFB_PAIR_LAUNCH(128, 3);
          else DFB_PAIR_LAUNCH(96, 4);
#undef DFB_PAIR_LAUNCH
          DFB_LAUNCH_CHECK();
          done = true;
        }
      }
      if (done) {
      } else if (variant != 1) {
        DFB_CHECK(build_pull(P, st));
        if (P->items_rows != P->n_rows) {
          if (P->n_rows == P->N) {
            P->items_active = P->n_items;
          } else {
            DFB_CUDA(cudaMemcpyAsync(&P->items_active, P->row_item + P->n_rows, sizeof(int), cudaMemcpyDeviceToHost, st));
            DFB_CUDA(cudaStreamSynchronize(st));
          }
          P->items_rows = P->n_rows;
        }
        DFB_CHECK(ensure_dynamic_smem((const void*)k_jprep2, sizeof(f64) * 128 * PREC_S));
        k_jprep2<<<ceil_div(E, 128), 128, sizeof(f64) * 128 * PREC_S, st>>>(E, P->ien, d_xg, d_wg, P->prec);
        DFB_LAUNCH_CHECK();
        if (options().j_pull_plain) {   // the unstaged kernel, kept for measurement
          const int ni = P->items_active;
          if (overwrite)
            k_pullJ<1><<<ceil_div(ni, 128), 128, 0, st>>>(ni, P->item_meta, P->item_ptr, P->contrib, P->prec, P->row_ptr, d_A00, d_A01, d_A10, d_A11);
          else
            k_pullJ<0><<<ceil_div(ni, 128), 128, 0, st>>>(ni, P->item_meta, P->item_ptr, P->contrib, P->prec, P->row_ptr, d_A00, d_A01, d_A10, d_A11);
        } else {
          const int nst = std::min(P->max_cta_elems, PULL_MAX_STAGED);
          const size_t smem = (size_t)std::max(1, nst) * PULL_SREC * sizeof(double2);
          DFB_CHECK(ensure_dynamic_smem((const void*)k_pullJ_staged<0>, smem));
          DFB_CHECK(ensure_dynamic_smem((const void*)k_pullJ_staged<1>, smem));
          const int ncta = ceil_div(P->n_rows, PULL_ROWS);
          if (overwrite)
            k_pullJ_staged<1><<<ncta, 128, smem, st>>>(N, P->n_rows, P->row_item, P->item_meta, P->item_ptr, P->contrib, P->contrib16,
                                                       P->cta_elem_ptr, P->cta_elems, P->prec, P->row_ptr, d_A00, d_A01, d_A10, d_A11);
          else
            k_pullJ_staged<0><<<ncta, 128, smem, st>>>(N, P->n_rows, P->row_item, P->item_meta, P->item_ptr, P->contrib, P->contrib16,
                                                       P->cta_elem_ptr, P->cta_elems, P->prec, P->row_ptr, d_A00, d_A01, d_A10, d_A11);
        }
        DFB_LAUNCH_CHECK();
      } else {
        if (P->max_row_len <= 16)
          k_rowJ<1><<<grid, 128, 0, st>>>(P->n_rows, P->ien, d_xg, d_wg, P->row_ptr, P->v2c_ptr, P->v2c, slot32, d_A00, d_A01, d_A10, d_A11, overwrite);
        else if (P->max_row_len <= 32)
          k_rowJ<2><<<grid, 128, 0, st>>>(P->n_rows, P->ien, d_xg, d_wg, P->row_ptr, P->v2c_ptr, P->v2c, slot32, d_A00, d_A01, d_A10, d_A11, overwrite);
        else
          k_rowJ<4><<<grid, 128, 0, st>>>(P->n_rows, P->ien, d_xg, d_wg, P->row_ptr, P->v2c_ptr, P->v2c, slot32, d_A00, d_A01, d_A10, d_A11, overwrite);
        DFB_LAUNCH_CHECK();
      }
    } else if (mode == DFB_MODE_ATOMIC) {
      k_cornerJ<1><<<ceil_div(4 * (i64)E, 128), 128, 0, st>>>(E, nullptr, P->ien, d_xg, d_wg, P->row_ptr, slot32, d_A00, d_A01, d_A10, d_A11);
      DFB_LAUNCH_CHECK();
    } else {
      for (int b = 0; b < P->num_batch; b++) {
        int n = P->batch_offset[b + 1] - P->batch_offset[b];
        if (n == 0) break;
        k_cornerJ<0><<<ceil_div(4 * (i64)n, 128), 128, 0, st>>>(n, P->batch_ind + P->batch_offset[b], P->ien, d_xg, d_wg, P->row_ptr, slot32, d_A00, d_A01, d_A10, d_A11);
        DFB_LAUNCH_CHECK();
      }
    }
  }
  return DFB_OK;
}

int dfb_assemble_face(const dfb_plan* P, int nf, const int* d_f2e, const int* d_forn, const double* d_xg,
                      const double* d_wg, const double* d_dwg, double* d_F, double* d_A00, double* d_A01, double* d_A10,
                      double* d_A11, void* stream) {
  cudaStream_t st = as_stream(stream);
  NvtxRange nvtx("dfb_assemble_face");
  if (!P || nf < 0 || !d_xg || !d_wg || !d_dwg) { set_error("dfb_assemble_face: bad argument"); return DFB_ERR_ARG; }
  if (nf == 0 || (!d_F && !d_A00)) return DFB_OK;
  if (d_A00 && (!d_A01 || !d_A10 || !d_A11)) { set_error("dfb_assemble_face: all four sub-block arrays are required"); return DFB_ERR_ARG; }
  if (d_A00 && !P->slot) { set_error("dfb_assemble_face: the plan was created without a sparsity pattern (residual only)"); return DFB_ERR_ARG; }
  k_face<<<ceil_div(nf, 128), 128, 0, st>>>(nf, d_f2e, d_forn, P->N, P->ien, d_xg, d_wg, d_dwg, P->row_ptr,
                                          reinterpret_cast<const u32*>(P->slot), d_F, d_A00, d_A01, d_A10, d_A11);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_dirichlet_vec(int nb, const int* d_bnode, int shape, const int* h_bctype, double* d_b, void* stream) {
  cudaStream_t st = as_stream(stream);
  if (nb < 0 || shape <= 0 || shape > 30 || !h_bctype || !d_b) { set_error("dfb_dirichlet_vec: bad argument"); return DFB_ERR_ARG; }
  int mask = 0;
  for (int ic = 0; ic < shape; ic++)
    if (h_bctype[ic] == 1) mask |= 1 << ic;  // BC_STRONG, reference dirichlet.h:8-13
  if (!mask || nb == 0) return DFB_OK;
  k_dirichlet_vec<<<ceil_div(nb, 128), 128, 0, st>>>(nb, d_bnode, shape, mask, d_b);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_dirichlet_mat(int nb, const int* d_bnode, int shape, const int* h_bctype, int N, const int* d_row_ptr,
                      const int* d_col_ind, double* d_A00, double* d_A01, void* stream) {
  cudaStream_t st = as_stream(stream);
  if (nb < 0 || shape != 3 || !h_bctype || !d_row_ptr || !d_col_ind || !d_A00 || !d_A01) { set_error("dfb_dirichlet_mat: bad argument (shape must be 3)"); return DFB_ERR_ARG; }
  int mask = 0;
  for (int ic = 0; ic < 3; ic++)
    if (h_bctype[ic] == 1) mask |= 1 << ic;
  if (!mask || nb == 0) return DFB_OK;
  k_dirichlet_mat<<<ceil_div((i64)nb * 64, 128), 128, 0, st>>>(nb, d_bnode, mask, N, d_row_ptr, d_col_ind, d_A00, d_A01);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

}  // extern "C"
