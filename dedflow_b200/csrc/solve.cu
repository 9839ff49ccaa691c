// solve.cu -- fused field-split SpMV, block-Jacobi preconditioner and GMRES on B200.
//
// Replaces (reference paths relative to /root/reference/src):
//   matrix.c:471-524,101-165  MatrixFSAMVPBY: Dscal + 4 x cusparseSpMV(CSR_ALG2), each fenced by two
//                             cudaStreamSynchronize and two dense-vector descriptor create/destroy
//   pc.c:44-147, matrix_impl.cu:642-683  PCJacobi setup (host-built pointer arrays, batched LU) / apply
//   krylov.c:56-334, krylov_util.cu:5-19 GMRESSolvePrivate: per iteration 2 Dgemv + Dnrm2 + blocking scalar read
//                             + k one-element Drot launches + Drotg + memset + <<<1,1>>> kernel; 5 cudaMalloc/Free
//                             and a 1 GB memset per solve
//
// B200 design (DESIGN.md sections 3, 5, 6):
//   * SpMV: ONE kernel over the four sub-blocks.  Eight lanes own a nodal row: the 9+3+3+1 value streams of the row are
//     contiguous (blocked scalar-CSR layout), the nodal column index is read once (4 B per 16 values instead of the
//     reference's 4 B per value), x is gathered through L2 as one 32-byte sector per column (interleaved vectors).
//     Algorithmic traffic 132 B per nodal nonzero instead of 192 B.
//   * Krylov vectors hold only the live 4N rows (defect D4), interleaved per node; the dead tail b[4N:6N) is carried exactly as
//     one scalar per basis vector (it stays a multiple of b's tail), see tail_* below.  The basis is stored UNNORMALISED with
//     its scales on the device, so no kernel waits for a norm that has only just been reduced.
//   * CGS, three launches per iteration: mat-vec; multi-dot (raw dots, fixed-order two-stage reduction; its last block also
//     runs the previous iteration's Givens step beside the partial sums); update (w~ -= Q c, fused with the sum of squares of the
//     new column and with z~ = P^-1 w~).  Givens rotations, the Hessenberg column, beta and the scales live on the device; the
//     residual history is written straight into mapped host memory: the host synchronises only at the convergence tests
//     (every 20th iteration like the reference, DFB_GMRES_CHECK for a shorter interval) and copies nothing.
//   * The iterations between two tests replay from a CUDA graph per chunk.
//   * Peer-memory mode (one process per GPU): the halo of z~ leaves the update as plain stores into the neighbours, the dots and
//     the norm travel as self-validating LL words; flags and publications that would hold a kernel's completion for an NVLink
//     round trip are raised by the first block of the NEXT kernel.
//   * Workspace is persistent (dfb_gmres), nothing is allocated per solve.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "tma.cuh"

namespace dfb {

constexpr unsigned FULLM = 0xffffffffu;

// DFB_PROFILE=1: CUDA events around every launch of a solve, aggregated per kernel and printed to stderr when the solve
// returns (diagnostics only; the events serialise nothing but add ~2 us of host work per launch).
struct SolveProfiler {
  struct Rec { const char* name; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  bool on;
  int level;
  SolveProfiler() { level = options().profile; on = level != 0; }
  void begin(const char* name, cudaStream_t st) {
    if (!on) return;
    Rec r; r.name = name;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    recs.push_back(r);
  }
  void end(cudaStream_t st) { if (on) cudaEventRecord(recs.back().b, st); }
  // aggregates the records (waits for the last event), prints them unless level < 0, and returns "name:launches:total_ms;..."
  std::string report() {
    std::string out;
    if (!on) return out;
    struct Agg { const char* name; int n; double ms; };
    std::vector<Agg> agg;
    double total = 0.0;
    for (auto& r : recs) {
      cudaEventSynchronize(r.b);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.a, r.b);
      cudaEventDestroy(r.a); cudaEventDestroy(r.b);
      total += ms;
      if (level >= 2) fprintf(stderr, "[dfb profile] launch %-18s %8.2f us\n", r.name, 1e3 * ms);
      bool found = false;
      for (auto& g : agg) if (!strcmp(g.name, r.name)) { g.n++; g.ms += ms; found = true; break; }
      if (!found) agg.push_back({r.name, 1, (double)ms});
    }
    char buf[160];
    for (auto& g : agg) {
      if (level > 0) fprintf(stderr, "[dfb profile] %-18s n=%4d total=%9.3f ms avg=%8.2f us\n", g.name, g.n, g.ms, 1e3 * g.ms / g.n);
      snprintf(buf, sizeof(buf), "%s:%d:%.6f;", g.name, g.n, g.ms);
      out += buf;
    }
    if (level > 0) fprintf(stderr, "[dfb profile] sum of kernels %.3f ms\n", total);
    recs.clear();
    return out;
  }
};

// ---- scalar state of the Arnoldi/Givens recurrence, all on the device --------------------------------------
struct GmresScalars {
  f64 nrm2_live;   // sum of squares of the live part (after the cross-rank reduction)
  f64 tail2;       // |b[4N:6N)|^2 (after the cross-rank reduction)
  f64 inv_norm;    // 1/||w||
  f64 rnrm_init;
};

struct UpdateScalars {   // device arrays of the recurrence (all persistent in the workspace)
  GmresScalars* S;
  f64 *qs, *gv, *beta, *tailc, *res_hist;
};
__device__ void gmres_step_dev(int it, GmresScalars* S, f64* hcol, f64* gv, f64* beta, f64* tailc, f64* res_hist, f64* qs,
                               const f64* h_in, const f64* gv_in, const f64* tc_in);
__device__ __forceinline__ void gmres_step_stage(int it, const f64* hcol, const f64* gv, const f64* tailc, f64* h_s, f64* gv_s, f64* tc_s);

// ------------------------------------------------------------------------------------------------------------
// SpMV.  Row i of the output: u -> y[3*i + ii], p -> y[y_poff + i]; input x: u -> x[3*c + l], p -> x[x_poff + c].
// ------------------------------------------------------------------------------------------------------------
// G lanes cooperate on one nodal row (32/G rows per warp in flight).  The row's value streams are contiguous, so a group
// reads G consecutive doubles per load.  Per chunk of G nodal nonzeros a lane issues its 16 value loads and ONE column
// index load back to back; the column indices are then exchanged inside the group with shuffles (no lane reads an index
// twice, no second dependent index->x chain), so only the x gathers wait on a previous load.
// PEER = true (data-parallel, peer-memory mode): x is the shared z vector whose ghost entries are stored by the
// neighbouring GPUs; a block that owns rows >= n_interior first waits for this mat-vec's halo flags, and x is read
// through L2 (ld.cg) for the ghost columns (>= n_rows) instead of the non-coherent path.
template <bool PEER>
__device__ __forceinline__ f64 ld_x(const f64* p, bool ghost) { return (PEER && ghost) ? __ldcg(p) : __ldg(p); }

// One nodal row by the G lanes of a group.  p00/p01/p10/p11/pcol point at the row's value / column streams (global memory, or
// a shared-memory stage filled by the TMA unit: SMEM), len = nodal nonzeros of the row.  Returns the group-reduced
// (y0, y1, y2, yp) in every lane of the group.
// AOSX: x is interleaved per node, x[4*node + c] (c = 0..2 velocity, 3 pressure) -- the solver's own layout, one 32-byte sector
// per gathered column -- instead of the ABI layout x[3*node + c] / x[x_poff + node].
template <int G, bool PEER, bool SMEM, bool AOSX>
__device__ __forceinline__ void spmv_row(const f64* __restrict__ p00, const f64* __restrict__ p01, const f64* __restrict__ p10,
                                         const f64* __restrict__ p11, const int* __restrict__ pcol, int len, int lane, unsigned gmask,
                                         const f64* __restrict__ x, size_t x_poff, int n_rows, f64& y0, f64& y1, f64& y2, f64& yp) {
  const int len3 = 3 * len;
  y0 = 0.0; y1 = 0.0; y2 = 0.0; yp = 0.0;
  for (int c0 = 0; c0 < len; c0 += G) {   // group-uniform trip count (one trip for rows of up to G nonzeros)
    const int k = c0 + lane;
    const bool okk = k < len;
    const int kk = okk ? k : 0;
    f64 a0[3], a1[3], a2[3], ap[3];
    bool oku[3];
#pragma unroll
    for (int u = 0; u < 3; u++) {
      const int t = 3 * c0 + u * G + lane;
      oku[u] = t < len3;
      const int tt = oku[u] ? t : 0;
      if (SMEM) {
        a0[u] = p00[tt]; a1[u] = p00[len3 + tt]; a2[u] = p00[2 * len3 + tt]; ap[u] = p10[tt];
      } else {
        a0[u] = __ldcs(p00 + tt);
        a1[u] = __ldcs(p00 + len3 + tt);
        a2[u] = __ldcs(p00 + 2 * len3 + tt);
        ap[u] = __ldcs(p10 + tt);
      }
    }
    f64 b0, b1, b2, bp;
    int col;
    if (SMEM) {
      b0 = p01[kk]; b1 = p01[len + kk]; b2 = p01[2 * len + kk]; bp = p11[kk]; col = pcol[kk];
    } else {
      b0 = __ldcs(p01 + kk); b1 = __ldcs(p01 + len + kk); b2 = __ldcs(p01 + 2 * len + kk);
      bp = __ldcs(p11 + kk);
      col = __ldg(pcol + kk);
    }
    const f64 xp = okk ? ld_x<PEER>(AOSX ? x + (size_t)col * 4 + 3 : x + x_poff + col, col >= n_rows) : 0.0;
    f64 xv[3];
#pragma unroll
    for (int u = 0; u < 3; u++) {
      const int q = u * G + lane;          // position inside the chunk's 3*G velocity entries
      const int kl = q / 3, l = q - 3 * kl;
      const int cu = __shfl_sync(gmask, col, kl, G);
      xv[u] = oku[u] ? ld_x<PEER>(x + (size_t)cu * (AOSX ? 4 : 3) + l, cu >= n_rows) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 3; u++) {
      y0 = fma(a0[u], xv[u], y0);
      y1 = fma(a1[u], xv[u], y1);
      y2 = fma(a2[u], xv[u], y2);
      yp = fma(ap[u], xv[u], yp);
    }
    y0 = fma(b0, xp, y0);
    y1 = fma(b1, xp, y1);
    y2 = fma(b2, xp, y2);
    yp = fma(bp, xp, yp);
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    y0 += __shfl_xor_sync(gmask, y0, o);
    y1 += __shfl_xor_sync(gmask, y1, o);
    y2 += __shfl_xor_sync(gmask, y2, o);
    yp += __shfl_xor_sync(gmask, yp, o);
  }
}

template <bool AOSY>
__device__ __forceinline__ void spmv_store(f64* __restrict__ y, size_t y_poff, int row, f64 alpha, f64 beta, f64 y0, f64 y1, f64 y2,
                                           f64 yp) {
  f64* yu = y + (size_t)row * (AOSY ? 4 : 3);
  f64* ypp = AOSY ? yu + 3 : y + y_poff + row;
  if (beta == 0.0) {
    yu[0] = alpha * y0; yu[1] = alpha * y1; yu[2] = alpha * y2; *ypp = alpha * yp;
  } else {
    yu[0] = beta * yu[0] + alpha * y0; yu[1] = beta * yu[1] + alpha * y1; yu[2] = beta * yu[2] + alpha * y2;
    *ypp = beta * *ypp + alpha * yp;
  }
}

// The same row out of a shared-memory stage, two chunks of G nonzeros per trip: the values cost nothing to fetch, the only
// long latency left is the x gather through L2, so all 8 gathers of a lane (rows of up to 2G nonzeros: ONE trip) are issued
// before the first FMA.  Same summation order as spmv_row (chunk by chunk), hence bit-identical results.
template <int G, bool PEER, bool AOSX>
__device__ __forceinline__ void spmv_row_smem(const f64* __restrict__ p00, const f64* __restrict__ p01, const f64* __restrict__ p10,
                                              const f64* __restrict__ p11, const int* __restrict__ pcol, int len, int lane,
                                              unsigned gmask, const f64* __restrict__ x, size_t x_poff, int n_rows, f64& y0, f64& y1,
                                              f64& y2, f64& yp) {
  const int len3 = 3 * len;
  y0 = 0.0; y1 = 0.0; y2 = 0.0; yp = 0.0;
  for (int c0 = 0; c0 < len; c0 += 2 * G) {   // group-uniform trip count
    f64 xp[2], xv[2][3];
    int tt[2][3], kk[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int cb = c0 + h * G;
      const int k = cb + lane;
      const bool okk = k < len;
      kk[h] = okk ? k : -1;
      const int col = okk ? pcol[k] : 0;
      xp[h] = okk ? ld_x<PEER>(AOSX ? x + (size_t)col * 4 + 3 : x + x_poff + col, col >= n_rows) : 0.0;
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int q = u * G + lane;
        const int kl = q / 3, l = q - 3 * kl;
        const int cu = __shfl_sync(gmask, col, kl, G);
        const int t = 3 * cb + q;
        const bool ok = t < len3;
        tt[h][u] = ok ? t : -1;
        xv[h][u] = ok ? ld_x<PEER>(x + (size_t)cu * (AOSX ? 4 : 3) + l, cu >= n_rows) : 0.0;
      }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int t = tt[h][u] < 0 ? 0 : tt[h][u];   // xv is 0 there
        y0 = fma(p00[t], xv[h][u], y0);
        y1 = fma(p00[len3 + t], xv[h][u], y1);
        y2 = fma(p00[2 * len3 + t], xv[h][u], y2);
        yp = fma(p10[t], xv[h][u], yp);
      }
      const int k = kk[h] < 0 ? 0 : kk[h];
      y0 = fma(p01[k], xp[h], y0);
      y1 = fma(p01[len + k], xp[h], y1);
      y2 = fma(p01[2 * len + k], xp[h], y2);
      yp = fma(p11[k], xp[h], yp);
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    y0 += __shfl_xor_sync(gmask, y0, o);
    y1 += __shfl_xor_sync(gmask, y1, o);
    y2 += __shfl_xor_sync(gmask, y2, o);
    yp += __shfl_xor_sync(gmask, yp, o);
  }
}

#ifndef SPMV_MINB
// Resident CTAs per SM the register allocation is held to: 3 x 256 threads = 24 warps at <= 80 registers.  Two registers more
// (82 -> allocated as 88) silently drop the kernel to 2 CTAs per SM and cost 13 % (78.8 vs 69.5 us in-solve at 1M tets); 4 CTAs
// (64 registers, 44 bytes of spills) are slower as well.  (-DSPMV_MINB=n: A/B builds, _build.build_variant.)
#define SPMV_MINB 3
#endif
template <int G, bool PEER, bool AOSX, bool AOSY>
__global__ void __launch_bounds__(256, SPMV_MINB) k_spmv_fs(int row0, int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col_ind,
                                                 const f64* __restrict__ A00, const f64* __restrict__ A01,
                                                 const f64* __restrict__ A10, const f64* __restrict__ A11, f64 alpha,
                                                 const f64* __restrict__ x, size_t x_poff, f64 beta, f64* __restrict__ y,
                                                 size_t y_poff, const P2PView* __restrict__ pv, unsigned hoff, int n_interior,
                                                 int post_flag, const unsigned* __restrict__ run_if) {
  if (run_if && *run_if == 0u) return;   // x is known to be all zero (y = y - A 0): the caller's flag, grid-uniform
  bool halo_block = false;   // block-uniform: this block owns boundary rows (they reference ghost columns)
  if (PEER) {
    // Deferred halo flag: the update kernel before this launch stored z~'s boundary entries into the neighbours but left the
    // flag to us, so that its own completion (and with it the start of this mat-vec's interior rows, on both sides) does not
    // wait for a system-scope fence.  The first block raises it before anything else; the blocks that need the NEIGHBOURS'
    // flags are the last ones of the grid.
    if ((post_flag & 1) && blockIdx.x == 0 && (int)threadIdx.x < pv->n_nbr) {
      __threadfence_system();
      p2p_signal(pv->mbox_peer[pv->nbr[threadIdx.x]] + p2p_h_flag(pv->nranks, pv->rank), pv->seq_base[1] + hoff);
    }
    // Same for the partial norm the update parked (bit 1): a store to a peer at the very end of a kernel holds the kernel's
    // completion for the NVLink round trip; here it costs one block a few hundred cycles.  Its readers (the next multi-dot's
    // last block) are a whole mat-vec away.
    if ((post_flag & 2) && blockIdx.x == 0 && (int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + pv->nranks) {
      const unsigned long long sq = pv->seq_base[0] + (hoff - 1u);   // the all-reduce sequence number of that update
      ll_store(pv->mbox_peer[threadIdx.x - 32] + p2p_b_ll(pv->nranks, (int)(sq & 1ull), pv->rank), *pv->nrm_part, (unsigned)sq);
    }
    const int last_row = row0 + (int)((((size_t)blockIdx.x + 1) * blockDim.x - 1) / G);
    halo_block = last_row >= n_interior;
    if (halo_block) {
      if ((int)threadIdx.x < pv->n_nbr)
        p2p_wait(pv, pv->mbox_local + p2p_h_flag(pv->nranks, pv->nbr[threadIdx.x]), pv->seq_base[1] + hoff);
      __syncthreads();
    }
  }
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = row0 + (int)(gt / G);
  const int lane = (int)(threadIdx.x & (G - 1));
  const bool live = row < n_rows;
  // the lanes of THIS group (the groups of a warp may run different trip counts)
  const unsigned gmask = G == 32 ? FULLM : (((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  int start = 0, len = 0;
  if (live) {
    start = __ldg(row_ptr + row);
    len = __ldg(row_ptr + row + 1) - start;
  }
  f64 y0, y1, y2, yp;
  if (PEER && halo_block)   // ghost columns are read through L2; interior blocks run exactly the single-GPU row code
    spmv_row<G, true, false, AOSX>(A00 + (size_t)start * 9, A01 + (size_t)start * 3, A10 + (size_t)start * 3, A11 + start,
                                   col_ind + start, len, lane, gmask, x, x_poff, n_rows, y0, y1, y2, yp);
  else
    spmv_row<G, false, false, AOSX>(A00 + (size_t)start * 9, A01 + (size_t)start * 3, A10 + (size_t)start * 3, A11 + start,
                                    col_ind + start, len, lane, gmask, x, x_poff, n_rows, y0, y1, y2, yp);
  if (live && lane == 0) spmv_store<AOSY>(y, y_poff, row, alpha, beta, y0, y1, y2, yp);
}

// ------------------------------------------------------------------------------------------------------------
// SpMV, TMA ring (default when the arrays are 16-byte aligned).  The value streams of CONSECUTIVE rows are contiguous in all
// five arrays (A00: 9 doubles per nodal nonzero, A01/A10: 3, A11: 1, col_ind: 1 int), so a tile of TS_TR rows is five
// contiguous byte ranges + its TS_TR + 1 row pointers: one elected producer thread hands them to the TMA unit as 1-D bulk
// copies (cp.async.bulk -> SASS UBLKCP) into a ring of TS_STAGES shared-memory stages, each guarded by a full / empty
// mbarrier pair; eight consumer warps run the row arithmetic out of shared memory and only gather x through L2.  The bytes
// in flight (up to 3 x 66 KB per SM) no longer depend on the register budget -- the register-staged kernel above keeps
// 24 warps x 17 loads per SM in flight and exposes the value-load and x-gather latencies in series.
// Persistent: one CTA per SM walks the tiles round-robin.  Bulk copies need 16-byte aligned addresses and sizes: a range is
// widened to the enclosing 16-byte window (its first element then sits 0..12 bytes into the stage) and never leaves the
// array: the few bytes past the last full window of the LAST tile are copied by hand.  A tile with more nonzeros than a
// stage holds (very long rows) is computed straight from global memory by the same row routine.
// ------------------------------------------------------------------------------------------------------------
constexpr int TS_TR = 32;       // rows per tile
constexpr int TS_CAP = 512;     // nodal nonzeros a stage holds
constexpr int TS_STAGES = 3;
constexpr int TS_CWARPS = 8;    // consumer warps per stage (4 rows each per tile)
// consumer groups: group g takes the tiles j = g, g + GROUPS, ... of its CTA, so GROUPS tiles are in the row arithmetic at once
// (template parameter of the kernel: 3 -> 25 warps at 72 registers, 2 -> 17 warps at up to 96)
struct __align__(16) SpmvStage {
  f64 a00[9 * TS_CAP + 2];
  f64 a01[3 * TS_CAP + 2];
  f64 a10[3 * TS_CAP + 2];
  f64 a11[TS_CAP + 2];
  int col[TS_CAP + 8];
  int rowp[TS_TR + 4];
  int direct, pad_[3];
};
static_assert(sizeof(SpmvStage) % 16 == 0, "stage size");
static_assert(TS_TR == 4 * TS_CWARPS, "eight lanes per row, four rows per consumer warp");

// elements [lo_b, hi_b) (byte offsets) of a global array whose last byte is end_b - 1 -> dst, element lo at dst + (lo_b & 15).
// Returns the bytes handed to the TMA unit (to be expected on the barrier).
__device__ __forceinline__ unsigned stage_copy(void* dst, const void* src, size_t lo_b, size_t hi_b, size_t end_b, uint64_t* bar) {
  const size_t b0 = lo_b & ~(size_t)15;
  size_t b1 = (hi_b + 15) & ~(size_t)15;
  const size_t lim = end_b & ~(size_t)15;
  if (b1 > lim) {   // last tile of the array: the tail beyond the last full 16-byte window by hand (4-byte words)
    for (size_t b = lim; b < hi_b; b += 4)
      *reinterpret_cast<int*>(static_cast<char*>(dst) + (b - b0)) = *reinterpret_cast<const int*>(static_cast<const char*>(src) + b);
    b1 = lim;
  }
  if (b1 <= b0) return 0u;
  tma::bulk_g2s(dst, static_cast<const char*>(src) + b0, (unsigned)(b1 - b0), bar);
  return (unsigned)(b1 - b0);
}
__device__ __forceinline__ unsigned stage_bytes(size_t lo_b, size_t hi_b, size_t end_b) {
  const size_t b0 = lo_b & ~(size_t)15, lim = end_b & ~(size_t)15;
  size_t b1 = (hi_b + 15) & ~(size_t)15;
  if (b1 > lim) b1 = lim;
  return b1 > b0 ? (unsigned)(b1 - b0) : 0u;
}

// oversized tile: the register-staged row routine straight from global memory (kept out of line: it must not set the register
// budget of the streaming path)
template <bool AOSX>
__device__ __noinline__ void spmv_row_direct(const f64* p00, const f64* p01, const f64* p10, const f64* p11, const int* pcol, int len,
                                             int lane, unsigned gmask, const f64* x, size_t x_poff, int n_rows, f64& y0, f64& y1,
                                             f64& y2, f64& yp) {
  spmv_row<8, false, false, AOSX>(p00, p01, p10, p11, pcol, len, lane, gmask, x, x_poff, n_rows, y0, y1, y2, yp);
}

template <int TS_GROUPS, bool AOSX, bool AOSY>
__global__ void __launch_bounds__(32 * (TS_CWARPS * TS_GROUPS + 1), 1)
k_spmv_tma(int row0, int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col_ind, const f64* __restrict__ A00,
           const f64* __restrict__ A01, const f64* __restrict__ A10, const f64* __restrict__ A11, f64 alpha,
           const f64* __restrict__ x, size_t x_poff, f64 beta, f64* __restrict__ y, size_t y_poff) {
  extern __shared__ __align__(16) unsigned char ts_smem[];
  SpmvStage* stages = reinterpret_cast<SpmvStage*>(ts_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(ts_smem + sizeof(SpmvStage) * TS_STAGES);
  uint64_t* empty = full + TS_STAGES;
  const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TS_STAGES; s++) { tma::mbar_init(full + s, 1); tma::mbar_init(empty + s, TS_CWARPS); }
    tma::fence_barrier_init();
  }
  __syncthreads();
  const int ntile = (n_rows - row0 + TS_TR - 1) / TS_TR;
  if (warp == TS_CWARPS * TS_GROUPS) {
    // ------------------------------ producer: one elected lane ------------------------------
    if (lane32 != 0) return;
    const size_t nnz_end = (size_t)__ldg(row_ptr + n_rows);
    int t = blockIdx.x;
    int s_nz = 0, e_nz = 0;
    if (t < ntile) {
      s_nz = __ldg(row_ptr + row0 + t * TS_TR);
      e_nz = __ldg(row_ptr + min(n_rows, row0 + (t + 1) * TS_TR));
    }
    for (int j = 0; t < ntile; j++, t += gridDim.x) {
      const int stg = j % TS_STAGES;
      const unsigned ph = (unsigned)((j / TS_STAGES) & 1);
      const int r0 = row0 + t * TS_TR, r1 = min(n_rows, r0 + TS_TR);
      const size_t s = (size_t)s_nz, e = (size_t)e_nz;
      const int tn = t + gridDim.x;        // row pointers of the next tile: one iteration ahead of their use
      if (tn < ntile) {
        s_nz = __ldg(row_ptr + row0 + tn * TS_TR);
        e_nz = __ldg(row_ptr + min(n_rows, row0 + (tn + 1) * TS_TR));
      }
      SpmvStage* st = stages + stg;
      tma::mbar_wait(empty + stg, ph ^ 1u);   // passes at once on the first lap
      if (e - s > (size_t)TS_CAP) {
        st->direct = 1;
        tma::mbar_arrive(full + stg);
        continue;
      }
      st->direct = 0;
      const size_t rlo = (size_t)4 * r0, rhi = (size_t)4 * (r1 + 1), rend = (size_t)4 * ((size_t)n_rows + 1);
      const unsigned total = stage_bytes(72 * s, 72 * e, 72 * nnz_end) + 2u * stage_bytes(24 * s, 24 * e, 24 * nnz_end) +
                             stage_bytes(8 * s, 8 * e, 8 * nnz_end) + stage_bytes(4 * s, 4 * e, 4 * nnz_end) +
                             stage_bytes(rlo, rhi, rend);
      // hand-copied tails (inside stage_copy) are ordinary stores: they must precede the arrive below, so the copies are
      // issued first and the barrier is armed afterwards (a transaction count may complete before it is expected)
      stage_copy(st->a00, A00, 72 * s, 72 * e, 72 * nnz_end, full + stg);
      stage_copy(st->a01, A01, 24 * s, 24 * e, 24 * nnz_end, full + stg);
      stage_copy(st->a10, A10, 24 * s, 24 * e, 24 * nnz_end, full + stg);
      stage_copy(st->a11, A11, 8 * s, 8 * e, 8 * nnz_end, full + stg);
      stage_copy(st->col, col_ind, 4 * s, 4 * e, 4 * nnz_end, full + stg);
      stage_copy(st->rowp, row_ptr, rlo, rhi, rend, full + stg);
      tma::mbar_arrive_expect_tx(full + stg, total);
    }
    return;
  }
  // ------------------------------ consumers: 8 lanes per row, 4 rows per warp and tile ------------------------------
  // consumer group g (8 warps) owns stage g: it takes the tiles j = g, g + TS_STAGES, ... of this CTA
  constexpr int G = 8;
  const int lane = lane32 & (G - 1), grp = lane32 >> 3;
  const unsigned gmask = 0xffu << (lane32 & ~(G - 1));
  const int cgrp = warp / TS_CWARPS, cw = warp % TS_CWARPS;
  int t = blockIdx.x + cgrp * gridDim.x;
  for (int j = cgrp; t < ntile; j += TS_GROUPS, t += TS_GROUPS * gridDim.x) {
    const int stg = j % TS_STAGES;
    const unsigned ph = (unsigned)((j / TS_STAGES) & 1);
    const int r0 = row0 + t * TS_TR, r1 = min(n_rows, r0 + TS_TR);
    const int row = r0 + cw * 4 + grp;
    const bool live = row < r1;
    const SpmvStage* st = stages + stg;
    tma::mbar_wait(full + stg, ph);
    f64 y0, y1, y2, yp;
    if (st->direct) {
      int start = 0, len = 0;
      if (live) { start = __ldg(row_ptr + row); len = __ldg(row_ptr + row + 1) - start; }
      spmv_row_direct<AOSX>(A00 + (size_t)start * 9, A01 + (size_t)start * 3, A10 + (size_t)start * 3, A11 + start, col_ind + start, len,
                      lane, gmask, x, x_poff, n_rows, y0, y1, y2, yp);
    } else {
      const int* rp = st->rowp + (r0 & 3);
      const int s = rp[0];
      int k0 = 0, len = 0;
      if (live) { k0 = rp[row - r0] - s; len = rp[row - r0 + 1] - s - k0; }
      const int sh = s & 1;
      spmv_row_smem<G, false, AOSX>(st->a00 + sh + 9 * k0, st->a01 + sh + 3 * k0, st->a10 + sh + 3 * k0, st->a11 + sh + k0,
                              st->col + (s & 3) + k0, len, lane, gmask, x, x_poff, n_rows, y0, y1, y2, yp);
    }
    __syncwarp();
    if (lane32 == 0) tma::mbar_arrive(empty + stg);   // this warp is done reading the stage
    if (live && lane == 0) spmv_store<AOSY>(y, y_poff, row, alpha, beta, y0, y1, y2, yp);
  }
}

// measured in-solve on B200 (1M tets): G=8 67.8 us, G=16 71.8 us, G=32 122 us per mat-vec
static int spmv_group() { return options().spmv_g; }

static bool spmv_tma_on() { return options().spmv_tma != 0; }   // 0 selects the register-staged kernel (A/B measurements)

// rows [row0, row1).  layout: bit 0 = x interleaved per node (x[4 node + c]), bit 1 = y interleaved; 0 = the ABI layout.
enum { LAY_ABI = 0, LAY_XAOS = 1, LAY_YAOS = 2 };
int launch_spmv(int row0, int row1, const int* row_ptr, const int* col_ind, const f64* A00, const f64* A01, const f64* A10,
                const f64* A11, f64 alpha, const f64* x, size_t x_poff, f64 beta, f64* y, size_t y_poff, cudaStream_t st,
                int layout = LAY_ABI, const P2PView* pv = nullptr, unsigned hoff = 0, int n_interior = 0, int post_flag = 0,
                const unsigned* run_if = nullptr) {
  if (row1 <= row0) return DFB_OK;
  const i64 rows = row1 - row0;
  if (pv && layout != (LAY_XAOS | LAY_YAOS)) { set_error("launch_spmv: the peer-memory mat-vec runs on interleaved vectors only"); return DFB_ERR_ARG; }
  if (!pv && spmv_tma_on() && layout != LAY_YAOS &&
      (((uintptr_t)row_ptr | (uintptr_t)col_ind | (uintptr_t)A00 | (uintptr_t)A01 | (uintptr_t)A10 | (uintptr_t)A11) & 15u) == 0) {
    constexpr size_t smem = sizeof(SpmvStage) * TS_STAGES + 2 * TS_STAGES * sizeof(uint64_t);
    DFB_CHECK(ensure_dynamic_smem((const void*)k_spmv_tma<2, false, false>, smem));
    DFB_CHECK(ensure_dynamic_smem((const void*)k_spmv_tma<3, false, false>, smem));
    DFB_CHECK(ensure_dynamic_smem((const void*)k_spmv_tma<2, true, true>, smem));
    DFB_CHECK(ensure_dynamic_smem((const void*)k_spmv_tma<3, true, true>, smem));
    const int ntile = ceil_div(rows, TS_TR), grid = std::min(ntile, num_sms());
    const bool two = options().spmv_tma == 2;
#define DFB_TMA(GR, AX, AY)                                                                                                        \
  k_spmv_tma<GR, AX, AY><<<grid, 32 * (TS_CWARPS * GR + 1), smem, st>>>(row0, row1, row_ptr, col_ind, A00, A01, A10, A11, alpha, x, \
                                                                       x_poff, beta, y, y_poff)
    if (layout == LAY_ABI) { if (two) DFB_TMA(2, false, false); else DFB_TMA(3, false, false); }
    else { if (two) DFB_TMA(2, true, true); else DFB_TMA(3, true, true); }
#undef DFB_TMA
    DFB_LAUNCH_CHECK();
    return DFB_OK;
  }
  const int grid8 = ceil_div(rows * 8, 256);
  if (pv) {
    k_spmv_fs<8, true, true, true><<<grid8, 256, 0, st>>>(row0, row1, row_ptr, col_ind, A00, A01, A10, A11, alpha, x, x_poff, beta, y,
                                                         y_poff, pv, hoff, n_interior, post_flag, nullptr);
  } else if (layout == LAY_YAOS) {
    k_spmv_fs<8, false, false, true><<<grid8, 256, 0, st>>>(row0, row1, row_ptr, col_ind, A00, A01, A10, A11, alpha, x, x_poff, beta, y,
                                                           y_poff, nullptr, 0u, 0, 0, run_if);
  } else if (layout == (LAY_XAOS | LAY_YAOS)) {
    k_spmv_fs<8, false, true, true><<<grid8, 256, 0, st>>>(row0, row1, row_ptr, col_ind, A00, A01, A10, A11, alpha, x, x_poff, beta, y,
                                                          y_poff, nullptr, 0u, 0, 0, nullptr);
  } else if (layout == LAY_ABI) {
#define DFB_SPMV(G)                                                                                                              \
  k_spmv_fs<G, false, false, false><<<ceil_div(rows * G, 256), 256, 0, st>>>(row0, row1, row_ptr, col_ind, A00, A01, A10, A11,   \
                                                                             alpha, x, x_poff, beta, y, y_poff, nullptr, 0u, 0, 0, nullptr)
    switch (spmv_group()) {
      case 4: DFB_SPMV(4); break;
      case 32: DFB_SPMV(32); break;
      case 16: DFB_SPMV(16); break;
      default: DFB_SPMV(8); break;
    }
#undef DFB_SPMV
  } else {
    set_error("launch_spmv: unsupported layout %d", layout);
    return DFB_ERR_ARG;
  }
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------------------
// preconditioner
// ------------------------------------------------------------------------------------------------------------
// The reference stores the nodal diagonal block B row-major, then inverts and applies it with column-major
// BLAS (pc.c:75-78,104-112; matrix_impl.cu:663-672): the applied operator is (B^-1)^T (defect D3).
// dinv00[9*i + r + 3*c] is that column-major inverse; y_r = sum_c dinv[r + 3c] x_c.
__global__ void k_pc_setup(int N, const int* __restrict__ row_ptr, const int* __restrict__ col_ind,
                           const f64* __restrict__ A00, const f64* __restrict__ A11, f64* __restrict__ dinv00,
                           f64* __restrict__ dinv11, int* __restrict__ bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int start = row_ptr[i], end = row_ptr[i + 1], len = end - start;
  int lo = start, hi = end;  // columns are sorted: binary search for the diagonal
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (col_ind[mid] < i) lo = mid + 1; else hi = mid;
  }
  const int k = lo - start;
  if (lo >= end || col_ind[lo] != i) {   // no diagonal entry (orphan node in a converted mesh): identity block, flagged
    if (bad) *(volatile int*)bad = 1;
    for (int t = 0; t < 9; t++) dinv00[(size_t)i * 9 + t] = (t % 4 == 0) ? 1.0 : 0.0;
    dinv11[i] = 1.0;
    return;
  }
  // M = column-major view of the row-major block: M(r,c) = B(c,r)
  f64 B[3][3];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) B[r][c] = A00[(size_t)start * 9 + (size_t)k * 3 + (size_t)r * len * 3 + c];
  // inverse of M = (B^T)^-1 = (B^-1)^T ; compute C = B^-1 by cofactors, then store C^T column-major = C row-major
  const f64 c00 = B[1][1] * B[2][2] - B[1][2] * B[2][1], c01 = B[1][2] * B[2][0] - B[1][0] * B[2][2],
            c02 = B[1][0] * B[2][1] - B[1][1] * B[2][0];
  const f64 det = B[0][0] * c00 + B[0][1] * c01 + B[0][2] * c02;
  const f64 a11 = A11[start + k];
  if (det == 0.0 || a11 == 0.0 || !isfinite(det) || !isfinite(a11)) {   // singular diagonal block: identity, flagged (an inf here
    if (bad) *(volatile int*)bad = 2;                                    // would poison the whole Krylov basis)
    for (int t = 0; t < 9; t++) dinv00[(size_t)i * 9 + t] = (t % 4 == 0) ? 1.0 : 0.0;
    dinv11[i] = 1.0;
    return;
  }
  const f64 id = 1.0 / det;
  f64 C[3][3];  // C = B^-1
  C[0][0] = c00 * id; C[1][0] = c01 * id; C[2][0] = c02 * id;
  C[0][1] = (B[0][2] * B[2][1] - B[0][1] * B[2][2]) * id;
  C[1][1] = (B[0][0] * B[2][2] - B[0][2] * B[2][0]) * id;
  C[2][1] = (B[0][1] * B[2][0] - B[0][0] * B[2][1]) * id;
  C[0][2] = (B[0][1] * B[1][2] - B[0][2] * B[1][1]) * id;
  C[1][2] = (B[0][2] * B[1][0] - B[0][0] * B[1][2]) * id;
  C[2][2] = (B[0][0] * B[1][1] - B[0][1] * B[1][0]) * id;
  // Minv(r,c) = C(c,r); column-major storage: dinv[r + 3c] = C(c,r)
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) dinv00[(size_t)i * 9 + r + 3 * c] = C[c][r];
  dinv11[i] = 1.0 / a11;
}

// y = P^-1 x on the live rows (+ optional copy of the phi/T tail when tail_n > 0).
// x: u at x[3i], p at x[x_poff + i];  y likewise with y_poff.
__global__ void k_pc_apply(int n, const f64* __restrict__ dinv00, const f64* __restrict__ dinv11,
                           const f64* __restrict__ x, size_t x_poff, f64* __restrict__ y, size_t y_poff, size_t tail_n,
                           size_t x_toff, size_t y_toff) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const f64* D = dinv00 + (size_t)i * 9;
  const f64 x0 = x[(size_t)i * 3], x1 = x[(size_t)i * 3 + 1], x2 = x[(size_t)i * 3 + 2];
  f64* yu = y + (size_t)i * 3;
  yu[0] = D[0] * x0 + D[3] * x1 + D[6] * x2;
  yu[1] = D[1] * x0 + D[4] * x1 + D[7] * x2;
  yu[2] = D[2] * x0 + D[5] * x1 + D[8] * x2;
  y[y_poff + i] = x[x_poff + i] * dinv11[i];
  for (size_t t = i; t < tail_n; t += n) y[y_toff + t] = x[x_toff + t];
}

// The solver's own copy of the preconditioner: one 96-byte record per node, laid out for the two lanes that share a node in
// the update kernel (interleaved Krylov vectors: lane "half 0" holds (u0,u1) of the node, "half 1" holds (u2,p)):
//   [0..5]  D0 D1 | D3 D4 | D6 D7     -> half 0 forms z0 = D0 x0 + D3 x1 + D6 x2,  z1 = D1 x0 + D4 x1 + D7 x2
//   [6..11] D2 D5 | D8 d11 | 0 0      -> half 1 forms z2 = D2 x0 + D5 x1 + D8 x2,  zp = d11 xp
// (D = dinv00[9 node ..], column-major (B^-1)^T of defect D3; d11 = dinv11[node]).
constexpr int PCREC = 12;
__global__ void k_pc_pack(int n, const f64* __restrict__ dinv00, const f64* __restrict__ dinv11, f64* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const f64* D = dinv00 + (size_t)i * 9;
  f64* r = rec + (size_t)i * PCREC;
  r[0] = D[0]; r[1] = D[1]; r[2] = D[3]; r[3] = D[4]; r[4] = D[6]; r[5] = D[7];
  r[6] = D[2]; r[7] = D[5]; r[8] = D[8]; r[9] = dinv11[i]; r[10] = 0.0; r[11] = 0.0;
}

// one half (two of the four entries) of z = P^-1 x for one node; x0 x1 x2 xp = the node's entries
__device__ __forceinline__ double2 pc_half(const f64* __restrict__ rec, int half, f64 x0, f64 x1, f64 x2, f64 xp) {
  const double2* r2 = reinterpret_cast<const double2*>(rec) + 3 * half;
  const double2 a = __ldg(r2), b = __ldg(r2 + 1);
  if (half == 0) {
    const double2 c = __ldg(r2 + 2);
    return make_double2(a.x * x0 + b.x * x1 + c.x * x2, a.y * x0 + b.y * x1 + c.y * x2);
  }
  return make_double2(a.x * x0 + a.y * x1 + b.x * x2, xp * b.y);
}

// halo push of one half node record into the neighbours' z (peer-memory mode) if node i is boundary-owned; returns whether it
// stored anything.  The caller issues ONE __threadfence_system() after its last push (before the flag election below): a fence
// per push would stall the warp for an NVLink round trip every trip.
__device__ __forceinline__ bool push_half(const P2PView* __restrict__ pv, int tgt_base, int tgt_n, int i, int half, double2 v) {
  const int b = i - tgt_base;
  if (b < 0 || b >= tgt_n) return false;
  const int* tp = pv->tgt_ptr;
  f64* const* ta = pv->tgt_addr;   // resolved at connect time: two dependent loads before the store instead of five
  for (int t = tp[b], te = tp[b + 1]; t < te; t++) *reinterpret_cast<double2*>(ta[t] + 2 * half) = v;
  return true;
}

// raise this mat-vec's halo flag on every neighbour once the whole grid has pushed (last-block election on pv->push_ctr)
__device__ __forceinline__ void push_flag(const P2PView* __restrict__ pv, unsigned hoff) {
  __shared__ bool push_last;
  if (pv->n_nbr == 0) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    push_last = atomicAdd(pv->push_ctr, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!push_last) return;
  if (threadIdx.x == 0) *pv->push_ctr = 0u;
  if ((int)threadIdx.x < pv->n_nbr) {
    __threadfence_system();
    p2p_signal(pv->mbox_peer[pv->nbr[threadIdx.x]] + p2p_h_flag(pv->nranks, pv->rank), pv->seq_base[1] + hoff);
  }
}

// z (interleaved, local numbering) = P^-1 w (interleaved, owned nodes): the first basis vector of a solve (every later one
// leaves the update kernel already preconditioned).  One thread per half node.  pv != NULL: + halo push and flag.
__global__ void __launch_bounds__(256) k_pc_apply_aos(int n, const f64* __restrict__ rec, const f64* __restrict__ w, f64* __restrict__ z,
                                                      const P2PView* __restrict__ pv, unsigned hoff) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(t >> 1), half = (int)(t & 1);
  if (i < n) {
    const double2* w2 = reinterpret_cast<const double2*>(w) + (size_t)i * 2;
    const double2 lo = w2[0], hi = w2[1];
    const double2 v = pc_half(rec + (size_t)i * PCREC, half, lo.x, lo.y, hi.x, hi.y);
    reinterpret_cast<double2*>(z)[t] = v;
    if (pv && push_half(pv, pv->tgt_base, pv->tgt_n, i, half, v)) __threadfence_system();
  }
  if (pv) push_flag(pv, hoff);
}

// x (ABI layout, 6N) += P^-1 t (interleaved, owned nodes)   (krylov.c:313-319: the preconditioner on the combination)
__global__ void k_pc_add_live(int n, const f64* __restrict__ rec, const f64* __restrict__ t, f64* __restrict__ x, size_t poff) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2* t2 = reinterpret_cast<const double2*>(t) + (size_t)i * 2;
  const double2 lo = t2[0], hi = t2[1];
  const double2 a = pc_half(rec + (size_t)i * PCREC, 0, lo.x, lo.y, hi.x, hi.y);
  const double2 b = pc_half(rec + (size_t)i * PCREC, 1, lo.x, lo.y, hi.x, hi.y);
  f64* xu = x + (size_t)i * 3;
  xu[0] += a.x; xu[1] += a.y; xu[2] += b.x;
  x[poff + i] += b.y;
}

// ------------------------------------------------------------------------------------------------------------
// Krylov vector kernels.  A "live" vector has nl = 4*n_own entries, INTERLEAVED per node: v[4 i + c], c = 0..2 velocity,
// c = 3 pressure (whole nodes stay together, so the block-Jacobi preconditioner can be applied where a vector is produced).
//
// The basis is stored UNNORMALISED: column j holds w~_j with v_j = s_j w~_j, s_j = 1/||w~_j|| kept on the device (qs[]).  One
// Arnoldi step is then three kernels instead of four, and no kernel waits for a norm that was only just reduced:
//   mat-vec   w_raw = A z~_j                       z~_j = P^-1 w~_j left behind by the previous update (j = 0: k_pc_apply_aos)
//   multi-dot d_i = w~_i . w_raw, i <= j           raw partial sums; nothing scalar is needed
//   update    h_i = s_i s_j d_i;  w~_{j+1} = s_j w_raw - sum_i (h_i s_i) w~_i;  ||w~_{j+1}||^2;  z~_{j+1} = P^-1 w~_{j+1}
// The scale s_j (the norm reduced at the end of the PREVIOUS update) is first needed in the update's prologue, a mat-vec and a
// multi-dot later: on several GPUs its all-reduce is off the critical path, and the halo of z~ leaves from the update's epilogue.
// In exact arithmetic this is the reference's recurrence (krylov.c:140-290); the roundings differ at the 1e-16 level.
// ------------------------------------------------------------------------------------------------------------
constexpr int NCHUNK = 296;    // 2 x 148 SMs: row chunks (partial-sum slots) of the multi-dot and the plain reductions
constexpr int UCHUNK = 9472;   // 64 x 148 SMs: most blocks of the update kernel (one partial sum each)
constexpr int JT = 8;          // basis columns per block in the multi-dot

__device__ __forceinline__ f64 block_sum_256(f64 v, f64* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = v;
  __syncthreads();
  f64 r = 0.0;
  if (w == 0) {
    r = l < 8 ? sm[l] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(FULLM, r, o);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

// "last block done" election: every block publishes its partial sums, fences, and bumps a counter; the block that sees
// the final count performs the second reduction stage in a FIXED order (results are run-to-run deterministic no matter
// which block is last) and re-arms the counter.
__device__ __forceinline__ bool last_block(unsigned* ctr, unsigned total) {
  __shared__ bool is_last;
  __syncthreads();   // every thread's partial-sum store is ordered before thread 0's fence + counter increment
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(ctr, 1u) == total - 1;
  }
  __syncthreads();
  return is_last;
}


// Stage 1 of the multi-dot for a group of NJ columns: a thread streams 16-byte pairs of rows, U row pairs per trip:
// (NJ + 1) x U independent 16-byte loads per trip (18 for a full group).  ptxas keeps the kernel at 64 registers (4 CTAs of
// 256 threads per SM) and issues them about five at a time; with 1024 resident threads per SM that is still several times the
// bytes in flight the HBM latency-bandwidth product asks for.
template <int NJ>
__device__ __forceinline__ void md_accum(const double2* __restrict__ q2, const double2* __restrict__ w2, size_t np, size_t ld2,
                                         f64 (&acc)[JT]) {
  constexpr int U = NJ <= 2 ? 4 : 2;
  const size_t stride = (size_t)gridDim.x * 256;
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  for (; i + (U - 1) * stride < np; i += U * stride) {
    double2 wv[U], qv[U][NJ];
#pragma unroll
    for (int u = 0; u < U; u++) {
      wv[u] = w2[i + u * stride];
#pragma unroll
      for (int j = 0; j < NJ; j++) qv[u][j] = q2[(size_t)j * ld2 + i + u * stride];
    }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < NJ; j++) acc[j] = fma(qv[u][j].y, wv[u].y, fma(qv[u][j].x, wv[u].x, acc[j]));
  }
  for (; i < np; i += stride) {
    const double2 wi = w2[i];
    double2 qv[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) qv[j] = q2[(size_t)j * ld2 + i];
#pragma unroll
    for (int j = 0; j < NJ; j++) acc[j] = fma(qv[j].y, wi.y, fma(qv[j].x, wi.x, acc[j]));
  }
}

// d[j] = sum_i Q[i, j] * w[i],  j in [0, ncol) (raw dots of the unnormalised columns): stage 1 = per-(chunk, column) partials,
// stage 2 by the last block.
// The columns are dealt EVENLY to the gridDim.y column groups (at most JT each: 20 columns -> 7 + 7 + 6, not 8 + 8 + 4), so
// that every block of the single resident wave streams the same number of bytes.  nl is a multiple of 4 and every column
// starts 32-byte aligned.
__global__ void __launch_bounds__(256) k_multidot(size_t nl, const f64* __restrict__ Q, size_t ldq, int ncol,
                                                  const f64* __restrict__ w, f64* part, f64* __restrict__ h,
                                                  unsigned* ctr, const P2PView* __restrict__ pv, unsigned soff, unsigned soff_prev,
                                                  UpdateScalars U, f64* hcol_prev) {
  __shared__ f64 smj[8][JT];
  __shared__ f64 hsum[P2P_ACAP];
  __shared__ f64 h_s[128], gv_s[256], tc_s[128];   // staged inputs of the previous step's Givens update (peer-memory mode)
  __shared__ f64 slot_s[P2P_MAXR * P2P_ACAP];      // the ranks' partial dots as they arrive in the mailbox
  const int ngrp = gridDim.y, gbase = ncol / ngrp, grem = ncol - gbase * ngrp;
  const int j0 = blockIdx.y * gbase + min((int)blockIdx.y, grem);
  const int nj = gbase + ((int)blockIdx.y < grem ? 1 : 0);
  f64 acc[JT];
#pragma unroll
  for (int j = 0; j < JT; j++) acc[j] = 0.0;
  const double2* q2 = reinterpret_cast<const double2*>(Q + (size_t)j0 * ldq);
  const double2* w2 = reinterpret_cast<const double2*>(w);
  const size_t np = nl >> 1, ld2 = ldq >> 1;
  switch (nj) {
    case 8: md_accum<8>(q2, w2, np, ld2, acc); break;
    case 7: md_accum<7>(q2, w2, np, ld2, acc); break;
    case 6: md_accum<6>(q2, w2, np, ld2, acc); break;
    case 5: md_accum<5>(q2, w2, np, ld2, acc); break;
    case 4: md_accum<4>(q2, w2, np, ld2, acc); break;
    case 3: md_accum<3>(q2, w2, np, ld2, acc); break;
    case 2: md_accum<2>(q2, w2, np, ld2, acc); break;
    default: md_accum<1>(q2, w2, np, ld2, acc); break;
  }
  {  // block reduction of the JT accumulators with a single barrier: warp shuffles, then thread j sums the 8 warps
    const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < JT; j++) {
      f64 v = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
      if (ln == 0) smj[wid][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < nj) {
      f64 r = 0.0;
#pragma unroll
      for (int wv = 0; wv < 8; wv++) r += smj[wv][threadIdx.x];
      part[(size_t)(j0 + threadIdx.x) * NCHUNK + blockIdx.x] = r;
    }
  }
  if (!last_block(ctr, gridDim.x * gridDim.y)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = gridDim.x;
  // The scalar Givens step of the PREVIOUS iteration (nothing has needed it until now: its outputs are the scale and the
  // dead-tail coefficient the coming update reads) runs on one thread of the last warp WHILE the other warps sum the partial
  // dots -- two serial tails of ~1.5 us each side by side instead of one after the other at the end of two kernels.
  const bool prev = soff_prev != 0u;
  const int itp = ncol - 2;   // the previous iteration
  if (prev) {
    gmres_step_stage(itp, hcol_prev, U.gv, U.tailc, h_s, gv_s, tc_s);
    if (pv) {   // the ranks' partial norms of the newest column: published at the end of the previous update, long since arrived
      const unsigned long long sp = pv->seq_base[0] + soff_prev;
      if ((int)threadIdx.x < pv->nranks)
        smj[0][threadIdx.x] = ll_load(pv, pv->mbox_local + p2p_b_ll(pv->nranks, (int)(sp & 1ull), threadIdx.x), (unsigned)sp);
    }
    __syncthreads();
  }
  if (prev && warp == 7) {
    if (lane == 0) {
      if (pv) {
        f64 nrm2 = 0.0;
        for (int r = 0; r < pv->nranks; r++) nrm2 += smj[0][r];
        U.S->nrm2_live = nrm2;
      }
      gmres_step_dev(itp, U.S, hcol_prev, U.gv, U.beta, U.tailc, U.res_hist, U.qs, h_s, gv_s, tc_s);
    }
  } else {
    // column j by warp j mod nw; four columns per trip so that their loads are in flight together (per column the chunks are
    // still added in ascending order by the same lanes: bit-identical to the one-column loop)
    const int nw = prev ? 7 : 8;
    for (int j = warp; j < ncol; j += 4 * nw) {
      f64 sv[4] = {0.0, 0.0, 0.0, 0.0};
      for (int c = lane; c < nchunk; c += 32) {
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (j + u * nw < ncol) sv[u] += __ldcg(part + (size_t)(j + u * nw) * NCHUNK + c);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        f64 v = sv[u];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
        if (lane == 0 && j + u * nw < ncol) {
          if (pv) hsum[j + u * nw] = v; else h[j + u * nw] = v;
        }
      }
    }
  }
  if (pv) {
    // Peer-memory mode: the whole all-reduce finishes HERE, in one block per GPU -- publish this rank's partial to every rank
    // (itself included), wait for every rank's partial in our own mailbox, sum in rank order (bit-identical on all ranks) and
    // leave the result in global memory for the update kernel, whose ~600 blocks then read it like on one GPU.  (Every block
    // of the update polling the mailbox itself -- tens of thousands of uncached loads on a dozen lines -- measured +11 us.)
    __syncthreads();
    const unsigned long long seq = pv->seq_base[0] + soff;
    const int R = pv->nranks, par = (int)(seq & 1ull);
    for (int t = threadIdx.x; t < R * ncol; t += 256) {
      const int r = t / ncol, j = t - r * ncol;
      ll_store(pv->mbox_peer[r] + p2p_a_ll(R, par, pv->rank, j), hsum[j], (unsigned)seq);
    }
    // all slots are polled in parallel (one thread per (rank, column)), then summed in rank order
    for (int t = threadIdx.x; t < R * ncol; t += 256) {
      const int r = t / ncol, j = t - r * ncol;
      slot_s[r * P2P_ACAP + j] = ll_load(pv, pv->mbox_local + p2p_a_ll(R, par, r, j), (unsigned)seq);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < ncol; j += 256) {
      f64 d = 0.0;
      for (int r = 0; r < R; r++) d += slot_s[r * P2P_ACAP + j];
      h[j] = d;
    }
  }
  if (threadIdx.x == 0) *ctr = 0u;
}

// The update of one Arnoldi step (see the header above): w~_{j+1} = s_j w_raw - sum_i (h_i s_i) w~_i with h_i = s_i s_j d_i,
// fused with (a) the sum of squares of the new vector (stage 2 by the last block), (b) z~_{j+1} = P^-1 w~_{j+1}: a node's four
// entries sit in two adjacent lanes, which swap their halves with one shuffle each and write z with coalesced 16-byte stores.
// mode 0: the scalar Arnoldi/Givens step runs here, in the last block (one GPU, DFB_GIVENS_DEFER=0).
// mode 1: the step runs later (the next multi-dot's tail, or a step kernel): block 0 leaves the Hessenberg column in global
//         memory, the last block the sum of squares (one GPU default; the NCCL path, which all-reduces it first).
// mode 2: peer memory: like mode 1, plus the halo stores of z~ into the neighbours in the sweep and this rank's partial norm
//         parked for the next kernel to publish (defer_flag) or published here together with the halo flags.
// One resident wave of blocks (4 per SM) strides over 16-byte row pairs; a thread keeps 8 independent 16-byte loads in flight.
#ifndef UPDATE_REVERSE
#define UPDATE_REVERSE 1
#endif
#ifndef UPDATE_LD
#define UPDATE_LD __ldcs
#endif
__global__ void __launch_bounds__(256) k_update(size_t nl, const f64* __restrict__ Q, size_t ldq, int ncol, const f64* __restrict__ draw,
                                                f64* hcol, f64* hcol_prev, f64* __restrict__ w, f64* part, unsigned* ctr, int mode,
                                                UpdateScalars U, const f64* __restrict__ pcrec, f64* __restrict__ z,
                                                const P2PView* __restrict__ pv, unsigned soff, unsigned hoff, int defer_flag) {
  __shared__ f64 sh[128];    // c_i = h_i s_i: coefficient of the stored (unnormalised) column i
  __shared__ f64 shh[128];   // h_i: the Hessenberg column
  __shared__ f64 sm[8];
  __shared__ f64 sgv[256], stc[128];   // staged inputs of the scalar Givens step (gmres_step_dev)
  const int jc = ncol - 1;   // index of the newest column w~_j (the one A z~ was formed from)
  const unsigned long long seq = pv ? pv->seq_base[0] + soff : 0ull;
  // scale of the newest column (written by the previous step's Givens update): every thread reads it itself -- one broadcast
  // transaction per warp, in flight together with the coefficient loads below instead of a round trip + barrier before them
  const f64 sj = U.qs[jc];
  for (int i = threadIdx.x; i < ncol; i += 256) {
    const f64 d = draw[i];   // raw dots, all-reduced (peer mode: by the multi-dot's last block)
    const f64 si = U.qs[i];
    const f64 h = si * sj * d;
    shh[i] = h;
    sh[i] = h * si;
    if (mode != 0 && blockIdx.x == 0) hcol[i] = h;   // a later kernel's Givens step reads the column from global memory
  }
  __syncthreads();
  const double2* q2 = reinterpret_cast<const double2*>(Q);
  double2* w2 = reinterpret_cast<double2*>(w);
  double2* z2 = reinterpret_cast<double2*>(z);
  const size_t np = nl >> 1, ld2 = ldq >> 1;   // np = 2 n_own: even, so the two halves of a node are always both in range
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * 256 + threadIdx.x) >> 5, nwarp = ((size_t)gridDim.x * 256) >> 5;
  f64 ss = 0.0;
  const int tgt_base = pv ? pv->tgt_base : 0, tgt_n = pv ? pv->tgt_n : 0;
  // The multi-dot has just swept the rows in ASCENDING order, so the highest rows of every column are what the 126 MB L2
  // still holds: sweep DESCENDING here (and leave the lowest rows behind for the next multi-dot's ascending sweep).
  const bool reverse = UPDATE_REVERSE;
  for (size_t base = warp0 * 32; base < np; base += nwarp * 32) {   // warp-uniform trip count
    const size_t ib = base + lane;
    const bool ok = ib < np;
    const size_t i = reverse ? (np - 1 - (ok ? ib : 0)) : (ok ? ib : 0);
    f64 ax[2] = {0.0, 0.0}, ay[2] = {0.0, 0.0};
    double2 wn = make_double2(0.0, 0.0);
    if (ok) {
      const double2* qi = q2 + i;
      for (int j = 0; j < ncol; j += 8) {
        if (j + 8 <= ncol) {
          double2 v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) v[u] = UPDATE_LD(qi + (size_t)(j + u) * ld2);
#pragma unroll
          for (int u = 0; u < 8; u++) {
            ax[u & 1] = fma(v[u].x, sh[j + u], ax[u & 1]);
            ay[u & 1] = fma(v[u].y, sh[j + u], ay[u & 1]);
          }
        } else {
          for (int jj = j; jj < ncol; jj++) {
            const double2 v = UPDATE_LD(qi + (size_t)jj * ld2);
            ax[0] = fma(v.x, sh[jj], ax[0]);
            ay[0] = fma(v.y, sh[jj], ay[0]);
          }
        }
      }
      wn = w2[i];
      wn.x = sj * wn.x - (ax[0] + ax[1]);
      wn.y = sj * wn.y - (ay[0] + ay[1]);
      w2[i] = wn;
      ss = fma(wn.y, wn.y, fma(wn.x, wn.x, ss));
    }
    // z~ = P^-1 w~ for the node this lane shares with its neighbour lane (pair index i: node i >> 1, half i & 1);
    // pcrec == NULL: another preconditioner follows as its own launches (pc2.cu)
    if (!pcrec) continue;   // kernel-uniform
    const f64 ox = __shfl_xor_sync(FULLM, wn.x, 1), oy = __shfl_xor_sync(FULLM, wn.y, 1);
    if (ok) {
      const int half = (int)(i & 1), node = (int)(i >> 1);
      const double2 v = half == 0 ? pc_half(pcrec + (size_t)node * PCREC, 0, wn.x, wn.y, ox, oy)
                                  : pc_half(pcrec + (size_t)node * PCREC, 1, ox, oy, wn.x, wn.y);
      z2[i] = v;
      if (pv) push_half(pv, tgt_base, tgt_n, node, half, v);
    }
  }
  // Halo stores need no fence of their own here (a system-scope fence per pushing thread measured +6 us on the kernel).
  // defer_flag (default): the flag is raised by the first block of the NEXT mat-vec -- the kernel boundary orders every store
  // of this grid before that block's system-scope fence -- because even ONE such fence in this kernel's last block held the
  // kernel's completion, i.e. the start of the mat-vec's interior rows on both sides, for ~9 us (2 GPUs: update 37.9 -> 29.3 us,
  // the single-GPU figure; solve 5.48 -> 5.35 ms).  Otherwise the stores are ordered before this block's arrival in the
  // election below (barrier + gpu-scope fence + atomic, last_block()), and the block that ends up last issues one cumulative
  // system-scope fence before it raises the flags.  ONE election serves the norm reduction and the halo.
  f64 r = block_sum_256(ss, sm);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
  if (!last_block(ctr, gridDim.x)) return;
  if (pv && !defer_flag && (int)threadIdx.x < pv->n_nbr) {   // every block has pushed: this mat-vec's halo is complete on the neighbours
    __threadfence_system();
    p2p_signal(pv->mbox_peer[pv->nbr[threadIdx.x]] + p2p_h_flag(pv->nranks, pv->rank), pv->seq_base[1] + hoff);
  }
  f64 s = 0.0;
  for (int c = threadIdx.x; c < (int)gridDim.x; c += 256) s += __ldcg(part + c);
  s = block_sum_256(s, sm);
  // mode 0 runs the scalar step here: its inputs from shared memory (h itself is in shh already)
  if (mode == 0) {   // kernel-uniform
    gmres_step_stage(jc, nullptr, U.gv, U.tailc, nullptr, sgv, stc);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *ctr = 0u;
    if (pv) {   // the partial sum of squares goes to every rank; a later kernel's Givens step sums the ranks
      if (defer_flag) {
        *pv->nrm_part = s;   // published by the next kernel (the mat-vec's first block, or the step kernel at the end of a chunk)
      } else {
        const int R = pv->nranks, par = (int)(seq & 1ull);
        for (int rr = 0; rr < R; rr++) ll_store(pv->mbox_peer[rr] + p2p_b_ll(R, par, pv->rank), s, (unsigned)seq);
      }
    } else {
      U.S->nrm2_live = s;
      if (mode == 0) gmres_step_dev(jc, U.S, hcol, U.gv, U.beta, U.tailc, U.res_hist, U.qs, shh, sgv, stc);
      // (mode 1: the column is in global memory already, block 0's prologue stored it)
    }
  }
}

// out[i] = sum_j Q[i,j] y[j]   (solution combination, krylov.c:303-311)
__global__ void __launch_bounds__(256) k_combine(size_t nl, const f64* __restrict__ Q, size_t ldq, int ncol,
                                                 const f64* __restrict__ yv, f64* __restrict__ out) {
  __shared__ f64 sh[128];
  for (int j = threadIdx.x; j < ncol; j += 256) sh[j] = yv[j];
  __syncthreads();
  const double2* q2 = reinterpret_cast<const double2*>(Q);
  double2* o2 = reinterpret_cast<double2*>(out);
  const size_t np = nl >> 1, ld2 = ldq >> 1;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < np; i += (size_t)gridDim.x * 256) {
    f64 sx = 0.0, sy = 0.0;
    const double2* qi = q2 + i;
    int j = 0;
    for (; j + 8 <= ncol; j += 8) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = __ldcs(qi + (size_t)(j + u) * ld2);
#pragma unroll
      for (int u = 0; u < 8; u++) { sx = fma(v[u].x, sh[j + u], sx); sy = fma(v[u].y, sh[j + u], sy); }
    }
    for (; j < ncol; j++) {
      const double2 v = __ldcs(qi + (size_t)j * ld2);
      sx = fma(v.x, sh[j], sx); sy = fma(v.y, sh[j], sy);
    }
    o2[i] = make_double2(sx, sy);
  }
}

// part[chunk] = partial sum of squares of v[0:n)
__global__ void __launch_bounds__(256) k_sumsq(size_t n, const f64* __restrict__ v, f64* __restrict__ part) {
  __shared__ f64 sm[8];
  f64 ss = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)NCHUNK * 256) ss = fma(v[i], v[i], ss);
  f64 r = block_sum_256(ss, sm);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}

// gather the live part of an ABI-layout vector (u at v[3i + c], p at v[poff + i]) into the interleaved layout out[4i + c]
// *flag = 1 if any of v[0, n) is not (plus or minus) zero -- NaN counts as nonzero.  The flag starts at 0.
__global__ void __launch_bounds__(256) k_nonzero_flag(size_t n, const f64* __restrict__ v, unsigned* flag) {
  bool any = false;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
    any |= (__double_as_longlong(v[i]) << 1) != 0ll;
  if (__syncthreads_or(any) && threadIdx.x == 0) *flag = 1u;
}

__global__ void k_pack_live(int n_own, const f64* __restrict__ v, size_t poff, f64* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)4 * n_own) return;
  const size_t i = t >> 2;
  const int c = (int)(t & 3);
  out[t] = c < 3 ? v[i * 3 + c] : v[poff + i];
}

// x (ABI layout) += d (interleaved live part)
__global__ void k_add_live_aos(int n_own, const f64* __restrict__ d, f64* __restrict__ x, size_t poff) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)4 * n_own) return;
  const size_t i = t >> 2;
  const int c = (int)(t & 3);
  if (c < 3) x[i * 3 + c] += d[t]; else x[poff + i] += d[t];
}

// x_tail += coef * b_tail
__global__ void k_axpy_dev(size_t n, const f64* __restrict__ coef, const f64* __restrict__ b, f64* __restrict__ x) {
  const f64 a = *coef;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] += a * b[i];
}

// reference BLAS drotg (krylov.c:266 calls cublasDrotg)
__device__ void drotg_dev(f64& a, f64& b, f64& c, f64& s) {
  const f64 roe = fabs(a) > fabs(b) ? a : b;
  const f64 scale = fabs(a) + fabs(b);
  f64 r, z;
  if (scale == 0.0) {
    c = 1.0; s = 0.0; r = 0.0; z = 0.0;
  } else {
    const f64 sa = a / scale, sb = b / scale;
    r = scale * sqrt(sa * sa + sb * sb);
    r = (roe < 0.0 ? -1.0 : 1.0) * r;
    c = a / r;
    s = b / r;
    z = 1.0;
    if (fabs(a) > fabs(b)) z = s;
    if (fabs(b) >= fabs(a) && c != 0.0) z = 1.0 / c;
  }
  a = r;
  b = z;
}

// sum the NCHUNK partials in fixed order (warp 0 of a 32-thread block) -> *out
__global__ void k_final_sum(const f64* __restrict__ part, f64* __restrict__ out) {
  f64 s = 0.0;
  for (int c = threadIdx.x; c < NCHUNK; c += 32) s += part[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULLM, s, o);
  if (threadIdx.x == 0) *out = s;
}

__global__ void k_set_seq_base(unsigned long long* base, unsigned long long seq, unsigned long long hseq) {
  base[0] = seq;
  base[1] = hseq;
}

// start of a solve: beta[0] = ||r0|| including the dead tail; tailc[0] = 1/beta0; inv_norm = 1/beta0
__global__ void k_gmres_begin(GmresScalars* S, f64* beta, f64* tailc, f64* res_hist, f64* qs) {
  const f64 n0 = sqrt(S->nrm2_live + S->tail2);
  S->rnrm_init = n0;
  beta[0] = n0;
  res_hist[0] = n0;
  S->inv_norm = 1.0 / n0;
  qs[0] = 1.0 / n0;
  tailc[0] = 1.0 / n0;
}

// one Arnoldi step's scalar work (krylov.c:229-277 + krylov_util.cu:5-19) for column `it`:
//   hcol[0..it] holds h = Q^T w (already reduced), S->nrm2_live the sum of squares of the updated live w.
// The recurrence reads its inputs -- h_in[0..it], gv_in[0..2 it), tc_in[0..it] -- from SHARED memory, where the calling block
// has staged them with coalesced loads: one thread walking them in global memory pays an L2 round trip per entry (~0.3 us each,
// 40 entries deep at the end of a solve), on the critical path of every Arnoldi step.  Outputs go to the global arrays.
__device__ void gmres_step_dev(int it, GmresScalars* S, f64* hcol, f64* gv, f64* beta, f64* tailc, f64* res_hist, f64* qs,
                               const f64* h_in, const f64* gv_in, const f64* tc_in) {
  // dead tail: every basis vector's rows [4N,6N) equal tailc[j] * b_tail (D4): w_tail = -sum_j h_j tailc[j] b_tail
  f64 cw = 0.0;
  for (int j = 0; j <= it; j++) cw -= h_in[j] * tc_in[j];
  const f64 nrm = sqrt(S->nrm2_live + cw * cw * S->tail2);
  const f64 inv = 1.0 / nrm;
  S->inv_norm = inv;
  qs[it + 1] = inv;             // scale of the (unnormalised) column it + 1
  tailc[it + 1] = cw * inv;
  f64 xx = h_in[0];
  for (int i = 0; i < it; i++) {  // cublasDrot, n = 1 (krylov.c:258-263); the running entry stays in a register
    const f64 c = gv_in[2 * i], s = gv_in[2 * i + 1], yy = h_in[i + 1];
    hcol[i] = c * xx + s * yy;
    xx = c * yy - s * xx;
  }
  f64 a = xx, b = nrm, c, s;
  drotg_dev(a, b, c, s);
  hcol[it] = a;
  hcol[it + 1] = 0.0;  // krylov.c:267
  gv[2 * it] = c;
  gv[2 * it + 1] = s;
  const f64 b0 = beta[it];
  beta[it + 1] = -s * b0;
  beta[it] = b0 * c;
  res_hist[it + 1] = fabs(beta[it + 1]);
}

// stage the inputs of step `it` (block-cooperative; the caller synchronises before thread 0 runs the step)
__device__ __forceinline__ void gmres_step_stage(int it, const f64* hcol, const f64* gv, const f64* tailc, f64* h_s, f64* gv_s, f64* tc_s) {
  for (int i = threadIdx.x; i <= it; i += blockDim.x) {
    if (hcol) h_s[i] = hcol[i];
    tc_s[i] = tailc[i];
  }
  for (int i = threadIdx.x; i < 2 * it; i += blockDim.x) gv_s[i] = gv[i];
}

__global__ void __launch_bounds__(128) k_gmres_step(int it, GmresScalars* S, f64* hcol, f64* gv, f64* beta, f64* tailc, f64* res_hist,
                                                    f64* qs) {
  __shared__ f64 h_s[128], gv_s[256], tc_s[128];
  gmres_step_stage(it, hcol, gv, tailc, h_s, gv_s, tc_s);
  __syncthreads();
  if (threadIdx.x == 0) gmres_step_dev(it, S, hcol, gv, beta, tailc, res_hist, qs, h_s, gv_s, tc_s);
}

// peer-memory mode: fused all-reduce of ||w||^2 (rank order) + the scalar Arnoldi/Givens step; one warp
__global__ void k_gmres_step_peer(int it, GmresScalars* S, f64* hcol, f64* gv, f64* beta, f64* tailc, f64* res_hist, f64* qs,
                                  const P2PView* __restrict__ pv, unsigned soff, int publish) {
  const unsigned long long seq = pv->seq_base[0] + soff;
  const int R = pv->nranks, par = (int)(seq & 1ull);
  __shared__ f64 s_part[P2P_MAXR];
  __shared__ f64 h_s[128], gv_s[256], tc_s[128];
  if (publish && (int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + R)   // the update parked this rank's partial for us
    ll_store(pv->mbox_peer[threadIdx.x - 32] + p2p_b_ll(R, par, pv->rank), *pv->nrm_part, (unsigned)seq);
  gmres_step_stage(it, hcol, gv, tailc, h_s, gv_s, tc_s);
  if ((int)threadIdx.x < R) s_part[threadIdx.x] = ll_load(pv, pv->mbox_local + p2p_b_ll(R, par, threadIdx.x), (unsigned)seq);
  __syncthreads();
  if (threadIdx.x == 0) {
    f64 s = 0.0;
    for (int r = 0; r < R; r++) s += s_part[r];
    S->nrm2_live = s;
    gmres_step_dev(it, S, hcol, gv, beta, tailc, res_hist, qs, h_s, gv_s, tc_s);
  }
}

// back substitution H[0:m,0:m] y = beta (krylov.c:297-301), then tail coefficient sum_j tailc[j] y[j].
// One block of 128 threads (m <= 127): H is staged in shared memory, thread i owns the right-hand side entry i, the
// column-oriented sweep needs one broadcast + one FMA per unknown instead of a serial O(m^2) chain of dependent loads.
__global__ void __launch_bounds__(128) k_gmres_trsv(int m, const f64* __restrict__ H, int ldh, f64* beta, const f64* tailc,
                                                    f64* tail_coef, const f64* __restrict__ qs, f64* __restrict__ ycoef) {
  extern __shared__ f64 hs[];   // [m][m] column-major copy of the triangle
  __shared__ f64 ycur;
  __shared__ f64 red[4];
  const int i = threadIdx.x;
  for (int t = threadIdx.x; t < m * m; t += blockDim.x) {
    const int c = t / m, r = t - c * m;
    hs[t] = r <= c ? H[(size_t)c * ldh + r] : 0.0;
  }
  f64 b = i < m ? beta[i] : 0.0;
  __syncthreads();
  for (int j = m - 1; j >= 0; j--) {
    if (i == j) { b = b / hs[j * m + j]; ycur = b; }
    __syncthreads();
    if (i < j) b -= hs[j * m + i] * ycur;
    __syncthreads();
  }
  if (i < m) {
    beta[i] = b;
    ycoef[i] = b * qs[i];   // coefficient of the stored (unnormalised) column i in the solution combination
  }
  f64 tc = i < m ? tailc[i] * b : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tc += __shfl_xor_sync(FULLM, tc, o);
  if ((i & 31) == 0) red[i >> 5] = tc;
  __syncthreads();
  if (i == 0) *tail_coef = (red[0] + red[1]) + (red[2] + red[3]);
}

}  // namespace dfb

using namespace dfb;

struct dfb_gmres {
  int N = 0, maxit = 0, ldh = 0;
  int n_own = 0;  // rows reduced in the inner products (== N on a single GPU)
  f64 *Q = nullptr, *H = nullptr, *gv = nullptr, *beta = nullptr, *tailc = nullptr;
  // Host-visible status block (pinned, mapped): the residual history [0, maxit] and, behind it, the preconditioner-setup flag.
  // The one thread that runs the scalar Arnoldi step stores each residual straight into it (a posted PCIe write), so the
  // convergence tests cost the host one stream synchronisation and no device-to-host copy.
  f64* h_status = nullptr;    // host address
  f64* res_hist = nullptr;    // the same memory as the device sees it
  int* pc_bad_dev = nullptr;  // device view of the flag word (h_status + maxit + 1)
  f64 *z = nullptr, *t = nullptr, *part = nullptr, *dinv00 = nullptr, *dinv11 = nullptr, *tail_coef = nullptr;
  f64 *qs = nullptr, *draw = nullptr, *ycoef = nullptr, *pcrec = nullptr;   // column scales, raw dots, combination coefficients, packed P^-1
  GmresScalars* S = nullptr;
  unsigned* ctr = nullptr;  // [2] last-block election counters (multi-dot, update)
  size_t bytes = 0;
  int n_interior = 0;
  dfb_parallel_ops par = {0, 0, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool parallel = false;
  dfb_pc2* pc2 = nullptr;   // opt-in two-level Schur-complement preconditioner (pc2.cu); nullptr = the reference's block-Jacobi
  std::string last_profile;   // per-kernel event times of the last solve (DFB_PROFILE != 0), see dfb_gmres_profile
  // CUDA graphs of the chunks of iterations between two convergence tests, valid for one set of matrix pointers and options
  struct GraphKey {
    const void *rp, *ci, *a00, *a01, *a10, *a11, *pv, *pc2;
    int n_own, n_interior, split;
    bool operator==(const GraphKey& o) const {
      return rp == o.rp && ci == o.ci && a00 == o.a00 && a01 == o.a01 && a10 == o.a10 && a11 == o.a11 && pv == o.pv && pc2 == o.pc2 &&
             n_own == o.n_own && n_interior == o.n_interior && split == o.split;
    }
  };
  GraphKey gkey = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0};
  std::vector<cudaGraphExec_t> gexec;
  cudaStream_t cap_stream = nullptr;
  void drop_graphs() {
    for (auto g : gexec) if (g) cudaGraphExecDestroy(g);
    gexec.clear();
  }
};

extern "C" {

int dfb_spmv_fs(int N, const int* d_row_ptr, const int* d_col_ind, const double* d_A00, const double* d_A01,
                const double* d_A10, const double* d_A11, double alpha, const double* d_x, double beta, double* d_y,
                void* stream) {
  if (N <= 0 || !d_row_ptr || !d_col_ind || !d_A00 || !d_A01 || !d_A10 || !d_A11 || !d_x || !d_y) { set_error("dfb_spmv_fs: bad argument"); return DFB_ERR_ARG; }
  return launch_spmv(0, N, d_row_ptr, d_col_ind, d_A00, d_A01, d_A10, d_A11, alpha, d_x, (size_t)3 * N, beta, d_y, (size_t)3 * N,
                     as_stream(stream));
}

int dfb_pc_setup(int N, const int* d_row_ptr, const int* d_col_ind, const double* d_A00, const double* d_A11,
                 double* d_dinv00, double* d_dinv11, void* stream) {
  if (N <= 0 || !d_row_ptr || !d_col_ind || !d_A00 || !d_A11 || !d_dinv00 || !d_dinv11) { set_error("dfb_pc_setup: bad argument"); return DFB_ERR_ARG; }
  k_pc_setup<<<ceil_div(N, 128), 128, 0, as_stream(stream)>>>(N, d_row_ptr, d_col_ind, d_A00, d_A11, d_dinv00, d_dinv11, nullptr);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_pc_apply(int N, const double* d_dinv00, const double* d_dinv11, const double* d_x, double* d_y, void* stream) {
  if (N <= 0 || !d_dinv00 || !d_dinv11 || !d_x || !d_y) { set_error("dfb_pc_apply: bad argument"); return DFB_ERR_ARG; }
  k_pc_apply<<<ceil_div(N, 128), 128, 0, as_stream(stream)>>>(N, d_dinv00, d_dinv11, d_x, (size_t)3 * N, d_y, (size_t)3 * N,
                                                             (size_t)2 * N, (size_t)4 * N, (size_t)4 * N);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_gmres_create(dfb_gmres** out, int N, int maxit) {
  if (!out || N <= 0 || maxit <= 0 || maxit > 127) { set_error("dfb_gmres_create: bad argument (max_iter must be in [1,127])"); return DFB_ERR_ARG; }
  dfb_gmres* w = new dfb_gmres();
  w->N = N; w->n_own = N; w->n_interior = N; w->maxit = maxit; w->ldh = ((maxit + 1 + 31) / 32) * 32;
  const size_t nl = (size_t)4 * N;
  struct { f64** p; size_t n; } allocs[] = {
      {&w->Q, nl * ((size_t)maxit + 1)}, {&w->H, (size_t)w->ldh * maxit}, {&w->gv, (size_t)2 * maxit},
      {&w->beta, (size_t)maxit + 1},    {&w->tailc, (size_t)maxit + 1},
      {&w->z, (size_t)6 * N},           {&w->t, (size_t)6 * N},          {&w->part, (size_t)NCHUNK * (maxit + 2) + UCHUNK},
      {&w->dinv00, (size_t)9 * N},      {&w->dinv11, (size_t)N},         {&w->tail_coef, 8},
      {&w->qs, (size_t)maxit + 2},      {&w->draw, (size_t)maxit + 2},   {&w->ycoef, (size_t)maxit + 2},
      {&w->pcrec, (size_t)PCREC * N}};
  for (auto& a : allocs) {
    if (cudaMalloc(a.p, a.n * sizeof(f64)) != cudaSuccess) {
      set_error("dfb_gmres_create: cudaMalloc of %zu bytes failed", a.n * sizeof(f64));
      dfb_gmres_destroy(w);
      return DFB_ERR_CUDA;
    }
    w->bytes += a.n * sizeof(f64);
  }
  if (cudaHostAlloc(&w->h_status, sizeof(f64) * ((size_t)maxit + 2), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer(&w->res_hist, w->h_status, 0) != cudaSuccess) {
    set_error("dfb_gmres_create: no mapped host memory for the status block");
    dfb_gmres_destroy(w);
    return DFB_ERR_CUDA;
  }
  memset(w->h_status, 0, sizeof(f64) * ((size_t)maxit + 2));
  w->pc_bad_dev = reinterpret_cast<int*>(w->res_hist + maxit + 1);
  DFB_CUDA(cudaMalloc(&w->S, sizeof(GmresScalars)));
  DFB_CUDA(cudaMalloc(&w->ctr, 4 * sizeof(unsigned)));   // [0..1] last-block elections, [3] "x0 is nonzero" flag
  DFB_CUDA(cudaMemset(w->ctr, 0, 4 * sizeof(unsigned)));
  DFB_CUDA(cudaMemset(w->z, 0, sizeof(f64) * 6 * (size_t)N));
  DFB_CUDA(cudaMemset(w->t, 0, sizeof(f64) * 6 * (size_t)N));
  *out = w;
  return DFB_OK;
}

void dfb_gmres_destroy(dfb_gmres* w) {
  if (!w) return;
  cudaFree(w->Q); cudaFree(w->H); cudaFree(w->gv); cudaFree(w->beta); cudaFree(w->tailc);
  if (w->h_status) cudaFreeHost(w->h_status);
  cudaFree(w->z); cudaFree(w->t); cudaFree(w->part); cudaFree(w->dinv00); cudaFree(w->dinv11); cudaFree(w->tail_coef);
  cudaFree(w->qs); cudaFree(w->draw); cudaFree(w->ycoef); cudaFree(w->pcrec);
  cudaFree(w->S); cudaFree(w->ctr);
  w->drop_graphs();
  if (w->cap_stream) cudaStreamDestroy(w->cap_stream);
  delete w;
}

size_t dfb_gmres_bytes(const dfb_gmres* w) { return w ? w->bytes : 0; }

int dfb_gmres_set_pc2(dfb_gmres* w, dfb_pc2* pc) {
  if (!w) { set_error("dfb_gmres_set_pc2: bad argument"); return DFB_ERR_ARG; }
  w->pc2 = pc;
  w->drop_graphs();
  return DFB_OK;
}

int dfb_gmres_profile(const dfb_gmres* w, char* buf, int capacity) {
  if (!w || !buf || capacity <= 0) { set_error("dfb_gmres_profile: bad argument"); return DFB_ERR_ARG; }
  snprintf(buf, (size_t)capacity, "%s", w->last_profile.c_str());
  return DFB_OK;
}

int dfb_gmres_set_parallel(dfb_gmres* w, const dfb_parallel_ops* ops) {
  if (!w) { set_error("dfb_gmres_set_parallel: bad argument"); return DFB_ERR_ARG; }
  if (!ops) { w->parallel = false; w->n_own = w->N; w->n_interior = w->N; return DFB_OK; }
  if (ops->n_own <= 0 || ops->n_own > w->N || ops->n_interior < 0 || ops->n_interior > ops->n_own || !ops->allreduce ||
      !ops->halo_begin || !ops->halo_end) { set_error("dfb_gmres_set_parallel: bad argument"); return DFB_ERR_ARG; }
  w->par = *ops; w->parallel = true; w->n_own = ops->n_own; w->n_interior = ops->n_interior;
  // The sequence numbers that tag the mailbox slots and halo flags live in the communicator (P2PHandle::seq / hseq), not
  // here: every workspace over the same mailbox draws from the same monotonic counters.
  return DFB_OK;
}

int dfb_gmres_solve(dfb_gmres* W, int N, const int* rp, const int* ci, const double* A00, const double* A01,
                    const double* A10, const double* A11, double* d_x, const double* d_b, double atol, double rtol,
                    int* iters, double* res_hist, void* stream) {
  return dfb_gmres_solve_pc(W, N, rp, ci, A00, A01, A10, A11, nullptr, nullptr, d_x, d_b, atol, rtol, iters, res_hist, stream);
}

int dfb_gmres_solve_pc(dfb_gmres* W, int N, const int* rp, const int* ci, const double* A00, const double* A01,
                       const double* A10, const double* A11, const double* ext_dinv00, const double* ext_dinv11,
                       double* d_x, const double* d_b, double atol, double rtol, int* iters, double* res_hist,
                       void* stream) {
  cudaStream_t st = as_stream(stream);
  NvtxRange nvtx("dfb_gmres_solve");
  if (!W || N != W->N || !rp || !ci || !A00 || !A01 || !A10 || !A11 || !d_x || !d_b || !iters) { set_error("dfb_gmres_solve: bad argument"); return DFB_ERR_ARG; }
  const int n_own = W->n_own, maxit = W->maxit, ldh = W->ldh;
  const size_t nl = (size_t)4 * n_own;          // compact live length
  const size_t ldq = nl;
  const size_t poffN = (size_t)3 * N;           // p offset in the 6N / local layout
  const size_t tail_n = (size_t)2 * N;
  f64* Q = W->Q;
#define QCOL(c) (Q + (size_t)(c)*ldq)
#define HCOL(c) (W->H + (size_t)(c)*ldh)
  const int vgrid = ceil_div((i64)nl, 256);
  const int mgrid = std::min(NCHUNK, ceil_div((i64)(nl / 2), 256));   // row chunks of the multi-dot
  // update kernel: one resident wave (4 blocks per SM; measured 27.4 us vs 30.4 us for one row pair per thread)
  const int ugrid = std::min(std::min(UCHUNK, 4 * num_sms()), ceil_div((i64)(nl / 2), 256));
  const int cgrid = std::min(1184, ceil_div((i64)(nl / 2), 256));              // blocks of the combine kernel
  const P2PHandle* ph = W->parallel ? static_cast<const P2PHandle*>(W->par.p2p) : nullptr;
  const P2PView* pv = ph ? ph->dev : nullptr;
  f64* const zvec = ph ? ph->host.z_local : W->z;   // peer-memory mode: z lives in the IPC-shared region (interleaved, 4 N_local)
  if (W->parallel && !pv && !W->par.halo_begin_aos) { set_error("dfb_gmres_solve: the NCCL path needs dfb_parallel_ops.halo_begin_aos"); return DFB_ERR_ARG; }
  if ((!ext_dinv00) != (!ext_dinv11)) { set_error("dfb_gmres_solve_pc: pass both preconditioner arrays or neither"); return DFB_ERR_ARG; }
  const f64 *dinv00 = W->dinv00, *dinv11 = W->dinv11;
  if (ext_dinv00) {  // the caller has run PCSetup already (drop-in layer: the PC tree owns the arrays)
    dinv00 = ext_dinv00; dinv11 = ext_dinv11;
  } else {           // preconditioner setup on every solve, like KrylovSolve -> PCSetup (krylov.c:453)
    *reinterpret_cast<volatile int*>(W->h_status + maxit + 1) = 0;   // (host store: every earlier solve of this workspace ended synchronised)
    k_pc_setup<<<ceil_div(n_own, 128), 128, 0, st>>>(n_own, rp, ci, A00, A11, W->dinv00, W->dinv11, W->pc_bad_dev);
    DFB_LAUNCH_CHECK();
  }
  k_pc_pack<<<ceil_div(n_own, 128), 128, 0, st>>>(n_own, dinv00, dinv11, W->pcrec);
  DFB_LAUNCH_CHECK();
  dfb_pc2* const pc2 = W->pc2;
  if (pc2) {
    if (W->parallel) { set_error("dfb_gmres_solve: the two-level preconditioner runs on one GPU only"); return DFB_ERR_ARG; }
    DFB_CHECK(dfb_pc2_setup(pc2, A00, A01, A10, A11, st));
  }
  DFB_CUDA(cudaMemsetAsync(W->H, 0, sizeof(f64) * (size_t)ldh * maxit, st));
  // r0 = b - A x  (krylov.c:114-118): x in the ABI layout (its ghosts refreshed by the callbacks), r0 interleaved
  k_pack_live<<<vgrid, 256, 0, st>>>(n_own, d_b, poffN, QCOL(0));
  DFB_LAUNCH_CHECK();
  if (W->parallel) {
    DFB_CHECK(W->par.halo_begin(d_x, st, W->par.user));
    DFB_CHECK(launch_spmv(0, W->n_interior, rp, ci, A00, A01, A10, A11, -1.0, d_x, poffN, 1.0, QCOL(0), 0, st, LAY_YAOS));
    DFB_CHECK(W->par.halo_end(d_x, st, W->par.user));
    DFB_CHECK(launch_spmv(W->n_interior, n_own, rp, ci, A00, A01, A10, A11, -1.0, d_x, poffN, 1.0, QCOL(0), 0, st, LAY_YAOS));
  } else {
    // a Newton driver starts every solve from x = 0 (main.c:211): one sweep over x decides whether the mat-vec has anything to do
    DFB_CUDA(cudaMemsetAsync(W->ctr + 3, 0, sizeof(unsigned), st));
    k_nonzero_flag<<<NCHUNK, 256, 0, st>>>((size_t)4 * N, d_x, W->ctr + 3);
    DFB_LAUNCH_CHECK();
    DFB_CHECK(launch_spmv(0, n_own, rp, ci, A00, A01, A10, A11, -1.0, d_x, poffN, 1.0, QCOL(0), 0, st, LAY_YAOS, nullptr, 0u, 0, 0,
                          W->ctr + 3));
  }
  k_sumsq<<<NCHUNK, 256, 0, st>>>(nl, QCOL(0), W->part);
  DFB_LAUNCH_CHECK();
  k_final_sum<<<1, 32, 0, st>>>(W->part, &W->S->nrm2_live);
  DFB_LAUNCH_CHECK();
  // dead tail of b (rows [4N,6N), defect D4): only its norm matters.  In the data-parallel case the tail rows of
  // ghost nodes must not be counted; the parallel driver passes b with a zero tail (the reference driver zeroes
  // it, main.c:63-66), so the local sum is exact.
  k_sumsq<<<NCHUNK, 256, 0, st>>>(tail_n, d_b + (size_t)4 * N, W->part + NCHUNK);
  DFB_LAUNCH_CHECK();
  k_final_sum<<<1, 32, 0, st>>>(W->part + NCHUNK, &W->S->tail2);
  DFB_LAUNCH_CHECK();
  if (W->parallel) DFB_CHECK(W->par.allreduce(&W->S->nrm2_live, 2, st, W->par.user));
  k_gmres_begin<<<1, 1, 0, st>>>(W->S, W->beta, W->tailc, W->res_hist, W->qs);
  DFB_LAUNCH_CHECK();
  DFB_CUDA(cudaMemsetAsync(W->ctr, 0, 2 * sizeof(unsigned), st));
  // z~_0 = P^-1 r0 (+ its halo); every later z~ leaves the update kernel
  // Peer-memory mode: the tags of this solve's fused collectives are {base + offset}; the base goes to the device once, the
  // offsets are fixed per iteration index (all-reduce: iteration + 1; halo of z~_j: j + 1), so the launches replay from graphs.
  if (ph) {
    k_set_seq_base<<<1, 1, 0, st>>>(ph->d_seq_base, *ph->seq, *ph->hseq);
    DFB_LAUNCH_CHECK();
  }
  if (pc2) {
    DFB_CHECK(pc2_apply_aos(pc2, A10, QCOL(0), zvec, st));
  } else {
    k_pc_apply_aos<<<ceil_div((i64)2 * n_own, 256), 256, 0, st>>>(n_own, W->pcrec, QCOL(0), zvec, pv, 1u);
    DFB_LAUNCH_CHECK();
  }

  int iter = 0;
  bool converged = false;
  unsigned peer_err = 0;   // peer-memory mode: raised by a kernel whose bounded wait for another rank ran out
  unsigned pc_bad = 0;     // raised by k_pc_setup: a node row without (1) or with a singular (2) diagonal block
  f64 rnrm_init = 0.0;
  std::vector<f64> hist((size_t)maxit + 1, 0.0);
  SolveProfiler prof;
  const UpdateScalars US = {W->S, W->qs, W->gv, W->beta, W->tailc, W->res_hist};
  // one Arnoldi step = three launches on stream `s` (+ the standalone collectives of the NCCL path)
  const int halo_defer = (pv && options().halo_defer && !options().spmv_peer_split) ? 1 : 0;
  // iterations between two convergence tests: 20 is the reference's rule (krylov.c:281-290) and the default; a shorter interval
  // (DFB_GMRES_CHECK, 1..20) stops closer to the first iteration that meets the tolerance at the price of one host round trip
  // per test -- same arithmetic, the residual history is a prefix of the default's
  const int chk = std::min(20, std::max(1, options().gmres_check));
  // the scalar Givens step of iteration j runs in the tail of iteration j + 1's multi-dot, next to the partial-dot sums
  // (peer-memory mode always; one GPU unless DFB_GIVENS_DEFER=0; the NCCL path keeps its own step kernel)
  const bool givens_defer = pv || (!W->parallel && options().givens_defer);
  auto arnoldi_step = [&](int iter, cudaStream_t s) -> int {
    // w_raw = A z~_iter into column iter + 1 (no scaling: the update applies s_iter)
    f64* w = QCOL(iter + 1);
    prof.begin("spmv", s);
    if (pv && options().spmv_peer_split) {   // interior rows by the plain kernel, the boundary rows (which wait for the halo) after them
      DFB_CHECK(launch_spmv(0, W->n_interior, rp, ci, A00, A01, A10, A11, 1.0, zvec, 0, 0.0, w, 0, s, LAY_XAOS | LAY_YAOS));
      DFB_CHECK(launch_spmv(W->n_interior, n_own, rp, ci, A00, A01, A10, A11, 1.0, zvec, 0, 0.0, w, 0, s, LAY_XAOS | LAY_YAOS, pv, (unsigned)iter + 1u, W->n_interior));
    } else if (pv) {   // ONE launch whose boundary-row blocks (scheduled last) wait for the neighbours' halo flags
      DFB_CHECK(launch_spmv(0, n_own, rp, ci, A00, A01, A10, A11, 1.0, zvec, 0, 0.0, w, 0, s, LAY_XAOS | LAY_YAOS, pv, (unsigned)iter + 1u, W->n_interior,
                            // bit 0: raise the flag of z~_iter's halo (z~_0's left k_pc_apply_aos); bit 1: publish the norm the
                            // previous update parked, unless that iteration closed a chunk (its step kernel did)
                            (halo_defer && iter > 0 ? 1 : 0) | (halo_defer && iter > 0 && iter % chk != 0 ? 2 : 0)));
    } else if (W->parallel) {
      DFB_CHECK(W->par.halo_begin_aos(zvec, s, W->par.user));
      DFB_CHECK(launch_spmv(0, W->n_interior, rp, ci, A00, A01, A10, A11, 1.0, zvec, 0, 0.0, w, 0, s, LAY_XAOS | LAY_YAOS));
      DFB_CHECK(W->par.halo_end(zvec, s, W->par.user));
      DFB_CHECK(launch_spmv(W->n_interior, n_own, rp, ci, A00, A01, A10, A11, 1.0, zvec, 0, 0.0, w, 0, s, LAY_XAOS | LAY_YAOS));
    } else {
      DFB_CHECK(launch_spmv(0, n_own, rp, ci, A00, A01, A10, A11, 1.0, zvec, 0, 0.0, w, 0, s, LAY_XAOS | LAY_YAOS));
    }
    prof.end(s);
    // raw dots d = Q^T w_raw  (krylov.c:166-174)
    const int ncol = iter + 1;
    const unsigned soff = pv ? (unsigned)iter + 1u : 0u;
    // peer-memory mode: the norm + Givens step of the previous iteration is still pending unless it closed a chunk (a convergence test follows a chunk)
    const unsigned soff_prev = (givens_defer && iter > 0 && iter % chk != 0) ? (unsigned)iter : 0u;
    prof.begin("multidot", s);
    const int ny = ceil_div(ncol, JT);
    const int mg = std::max(1, std::min(mgrid, (4 * num_sms()) / ny));   // one resident wave of (row chunk, column group) blocks
    k_multidot<<<dim3(mg, ny), 256, 0, s>>>(nl, Q, ldq, ncol, w, W->part, W->draw, W->ctr, pv, soff, soff_prev, US,
                                            iter ? HCOL(iter - 1) : HCOL(0));
    DFB_LAUNCH_CHECK();
    prof.end(s);
    if (W->parallel && !pv) {
      prof.begin("allreduce d", s);
      DFB_CHECK(W->par.allreduce(W->draw, ncol, s, W->par.user));
      prof.end(s);
    }
    // update + norm + P^-1 (+ Givens step on one GPU; + the fused collectives in peer-memory mode)  (krylov.c:176-183, 229-277)
    prof.begin("update", s);
    k_update<<<ugrid, 256, 0, s>>>(nl, Q, ldq, ncol, W->draw, HCOL(iter), iter ? HCOL(iter - 1) : HCOL(0), w, W->part, W->ctr + 1,
                                   pv ? 2 : ((W->parallel || givens_defer) ? 1 : 0), US, pc2 ? nullptr : W->pcrec, zvec, pv, soff,
                                   (unsigned)iter + 2u /* the halo of z~_{iter+1} leaves from this kernel */, halo_defer);
    DFB_LAUNCH_CHECK();
    prof.end(s);
    if (pc2) {
      prof.begin("pc2 apply", s);
      DFB_CHECK(pc2_apply_aos(pc2, A10, w, zvec, s));
      prof.end(s);
    }
    if (pv) {
      // the norm reduction + Givens step of this iteration are folded into the NEXT update's prologue; they run on their own
      // only when the host needs the residual now (the every-20th test) or the loop ends
      if ((iter + 1) % chk == 0 || iter + 1 == maxit) {
        prof.begin("step (peer sum)", s);
        k_gmres_step_peer<<<1, 128, 0, s>>>(iter, W->S, HCOL(iter), W->gv, W->beta, W->tailc, W->res_hist, W->qs, pv, soff, halo_defer);
        DFB_LAUNCH_CHECK();
        prof.end(s);
      }
    } else if (givens_defer) {   // one GPU: same folding, the step runs on its own where a convergence test (or the end) follows
      if ((iter + 1) % chk == 0 || iter + 1 == maxit) {
        prof.begin("step", s);
        k_gmres_step<<<1, 128, 0, s>>>(iter, W->S, HCOL(iter), W->gv, W->beta, W->tailc, W->res_hist, W->qs);
        DFB_LAUNCH_CHECK();
        prof.end(s);
      }
    } else if (W->parallel) {
      prof.begin("allreduce nrm+step", s);
      DFB_CHECK(W->par.allreduce(&W->S->nrm2_live, 1, s, W->par.user));
      k_gmres_step<<<1, 128, 0, s>>>(iter, W->S, HCOL(iter), W->gv, W->beta, W->tailc, W->res_hist, W->qs);
      DFB_LAUNCH_CHECK();
      prof.end(s);
    }
    return DFB_OK;
  };
  // the reference's only convergence test, after every 20th iteration (krylov.c:281-290)
  auto convergence_test = [&](int done) -> int {
    if (ph) DFB_CUDA(cudaMemcpyAsync(&peer_err, ph->d_err, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    DFB_CUDA(cudaStreamSynchronize(st));
    const volatile f64* hs = W->h_status;   // written by the device, complete once the stream is idle
    pc_bad = ext_dinv00 ? 0u : (unsigned)*reinterpret_cast<const volatile int*>(W->h_status + maxit + 1);
    hist[0] = hs[0];
    hist[(size_t)done] = hs[done];
    if (pc_bad) {
      set_error("dfb_gmres_solve: the block-Jacobi setup met a %s (node row without / with a singular diagonal block)",
                pc_bad == 1 ? "missing diagonal entry" : "singular diagonal block");
      return DFB_ERR_ARG;
    }
    if (peer_err) return DFB_OK;
    rnrm_init = hist[0];
    const f64 rnrm = hist[done];
    if (rnrm < atol || rnrm < (rnrm_init + 1e-16) * rtol) converged = true;
    return DFB_OK;
  };
  // The iterations between two convergence tests (20: 60 launches with fixed arguments for a given workspace and
  // matrix) are captured once into a CUDA graph per chunk and replayed -- the launch gaps shrink, the host does one call.
  const bool use_graph = (!W->parallel || pv) && !prof.on && options().graph != 0;
  const dfb_gmres::GraphKey gkey = {rp, ci, A00, A01, A10, A11, pv, pc2, n_own, W->n_interior, options().spmv_peer_split | (options().halo_defer << 1) | (options().givens_defer << 2) | (chk << 3)};
  while (!converged && iter < maxit && !peer_err) {
    if (use_graph && iter % chk == 0 && iter + chk <= maxit) {
      const size_t chunk = (size_t)iter / chk;
      if (!(W->gkey == gkey)) { W->drop_graphs(); W->gkey = gkey; }
      if (W->gexec.size() <= chunk) W->gexec.resize(chunk + 1, nullptr);
      if (!W->gexec[chunk]) {
        if (!W->cap_stream) DFB_CUDA(cudaStreamCreateWithFlags(&W->cap_stream, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        DFB_CUDA(cudaStreamBeginCapture(W->cap_stream, cudaStreamCaptureModeThreadLocal));
        int rc = DFB_OK;
        for (int k = 0; k < chk && rc == DFB_OK; k++) rc = arnoldi_step(iter + k, W->cap_stream);
        const cudaError_t ce = cudaStreamEndCapture(W->cap_stream, &graph);
        if (rc != DFB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        DFB_CUDA(ce);
        const cudaError_t ie = cudaGraphInstantiate(&W->gexec[chunk], graph, 0);
        cudaGraphDestroy(graph);
        DFB_CUDA(ie);
      } else {
        count_launch(chk * (pc2 ? 20 : (options().spmv_peer_split && pv ? 4 : 3)) + (givens_defer ? 1 : 0));   // (pc2: 3 + 5 + its Chebyshev steps, approx.)
      }
      DFB_CUDA(cudaGraphLaunch(W->gexec[chunk], st));
      iter += chk;
      DFB_CHECK(convergence_test(iter));
      continue;
    }
    DFB_CHECK(arnoldi_step(iter, st));
    iter++;
    if (iter % chk == 0) DFB_CHECK(convergence_test(iter));
  }
  if (ph) {   // identical control flow on every rank: the communicator's counters advance by what this solve used
    *ph->seq += (unsigned long long)iter;
    *ph->hseq += (unsigned long long)iter + 1ull;
  }
  if (iter && !peer_err) {
    prof.begin("trsv..axpy", st);
    DFB_CHECK(ensure_dynamic_smem((const void*)k_gmres_trsv, sizeof(f64) * 127 * 127));
    k_gmres_trsv<<<1, 128, sizeof(f64) * (size_t)iter * iter, st>>>(iter, W->H, ldh, W->beta, W->tailc, W->tail_coef, W->qs, W->ycoef);
    DFB_LAUNCH_CHECK();
    k_combine<<<cgrid, 256, 0, st>>>(nl, Q, ldq, iter, W->ycoef, W->t);
    DFB_LAUNCH_CHECK();
    // x += P^-1 (combination)  (krylov.c:313-319)
    if (pc2) {
      DFB_CHECK(pc2_apply_aos(pc2, A10, W->t, zvec, st));
      k_add_live_aos<<<vgrid, 256, 0, st>>>(n_own, zvec, d_x, poffN);
    } else {
      k_pc_add_live<<<ceil_div(n_own, 128), 128, 0, st>>>(n_own, W->pcrec, W->t, d_x, poffN);
    }
    DFB_LAUNCH_CHECK();
    k_axpy_dev<<<ceil_div((i64)tail_n, 256), 256, 0, st>>>(tail_n, W->tail_coef, d_b + (size_t)4 * N, d_x + (size_t)4 * N);
    DFB_LAUNCH_CHECK();
    prof.end(st);
    if (W->parallel) {  // leave the ghosts of the solution consistent
      DFB_CHECK(W->par.halo_begin(d_x, st, W->par.user));
      DFB_CHECK(W->par.halo_end(d_x, st, W->par.user));
    }
  }
  if (ph && !peer_err) DFB_CUDA(cudaMemcpyAsync(&peer_err, ph->d_err, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  DFB_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k <= iter; k++) hist[(size_t)k] = const_cast<const volatile f64*>(W->h_status)[k];
  if (!ext_dinv00) pc_bad = (unsigned)*reinterpret_cast<const volatile int*>(W->h_status + maxit + 1);
  if (pc_bad) {
    set_error("dfb_gmres_solve: the block-Jacobi setup met a %s (node row without / with a singular diagonal block)",
              pc_bad == 1 ? "missing diagonal entry" : "singular diagonal block");
    return DFB_ERR_ARG;
  }
  if (peer_err) {
    set_error("dfb_gmres_solve: a peer-memory wait timed out (a rank died or left the solve early); destroy the communicator");
    return DFB_ERR_PEER;
  }
  if (res_hist)
    for (int k = 0; k <= iter; k++) res_hist[k] = hist[k];
  *iters = iter;
  W->last_profile = prof.report();
#undef QCOL
#undef HCOL
  return DFB_OK;
}

}  // extern "C"
