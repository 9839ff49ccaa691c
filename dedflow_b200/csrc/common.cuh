// common.cuh -- shared host/device helpers of libdedflow_b200 (sm_100a only, no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/dedflow_b200.h"

typedef int32_t i32;
typedef int64_t i64;
typedef uint32_t u32;
typedef uint8_t u8;
typedef double f64;

namespace dfb {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace dfb
struct dfb_pc2;
// one application of the two-level Schur-complement preconditioner (pc2.cu) on interleaved vectors: z = P2^-1 w
int pc2_apply_aos(const dfb_pc2* P, const double* A10, const double* w, double* z, cudaStream_t st);
namespace dfb {

#define DFB_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      dfb::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,  \
                     cudaGetErrorString(_e));                                                 \
      return DFB_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define DFB_LAUNCH_CHECK()                                                                    \
  do {                                                                                        \
    dfb::count_launch();                                                                      \
    DFB_CUDA(cudaGetLastError());                                                             \
  } while (0)

#define DFB_CHECK(expr)                 \
  do {                                  \
    int _s = (expr);                    \
    if (_s != DFB_OK) return _s;        \
  } while (0)

inline int ceil_div(i64 a, i64 b) { return (int)((a + b - 1) / b); }

// Device scratch that is released on every return path (the DFB_CUDA / DFB_CHECK macros return early on errors).
// release() hands the pointer over to a longer-lived owner.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t count) {
    if (p) { cudaFree(p); p = nullptr; }
    const cudaError_t e = cudaMalloc(&p, sizeof(T) * (count ? count : 1));
    if (e != cudaSuccess) {
      p = nullptr;
      set_error("cudaMalloc of %zu bytes failed: %s", sizeof(T) * count, cudaGetErrorString(e));
      return DFB_ERR_CUDA;
    }
    return DFB_OK;
  }
  T* release() { T* q = p; p = nullptr; return q; }
  operator T*() const { return p; }
};
inline cudaStream_t as_stream(void* s) { return (cudaStream_t)s; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is process-wide state of a kernel: raise it under a lock and only ever upwards
// (entry points may be called from several host threads; no function-local static caches).
int ensure_dynamic_smem(const void* func, size_t bytes);

// NVTX range of one phase of the hot path (the reference has no tracing at all); a no-op unless a tool
// (nsys, ncu --nvtx) is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name);
  ~NvtxRange();
};

// number of SMs of the current device (B200: 148); cached
int num_sms();

// Variant switches of the library.  Read ONCE from the environment (first use) and afterwards only changed through
// dfb_set_option(): no entry point calls getenv on its launch path, and tests / A-B scripts switch variants inside one process
// without touching the environment.  Keys are the environment names (DFB_J_VARIANT=pairs|pull|fused, DFB_J_PAIR_ROWS=8|16, DFB_J_PAIR_NT=96|128|192|224,
// DFB_J_PAIR_ORDER=morton|natural, DFB_J_PULL_PLAIN=0|1, DFB_F_VARIANT=patch|pipe|scratch, DFB_F_PATCH_CTAS=2|3, DFB_SPMV_G=4|8|16|32, DFB_SPMV_TMA=0|1, DFB_SPMV_PEER_SPLIT=0|1, DFB_HALO_DEFER=0|1,
// DFB_GRAPH=0|1, DFB_GMRES_CHECK=1..20, DFB_GIVENS_DEFER=0|1, DFB_PROFILE=0|1|2, DFB_ASSEMBLE_MODE=gather|atomic|colored, DFB_VERBOSE=0|1, DFB_PC=jacobi|schur2, DFB_PC_AGG=2..16, DFB_PC_DEGREE=1..64).
struct Options {
  int j_variant = 2;        // 0 pull, 1 fused, 2 pairs
  int j_pair_rows = 8;
  int j_pair_natural = 0;
  int j_pair_nt = 96;       // threads per CTA of k_pairJ (96 | 128 for 8-row groups, 192 | 224 for 16-row groups)
  int j_pull_plain = 0;
  int f_variant = 1;        // 0 scratch (k_elemF + k_gatherF), 1 patch (k_patchF), 2 pipe (k_patchF_pipe: persistent, double-buffered)
  int f_patch_ctas = 2;     // register budget of k_patchF: resident CTAs per SM (2: 232 registers, measured faster; 3: 168 + spills)
  int spmv_g = 8;
  int spmv_tma = 0;        // 0 register-staged k_spmv_fs (default: measured faster in-solve), 1/2 TMA ring with 3/2 consumer groups
  int spmv_peer_split = 0;  // peer-memory mat-vec as two launches (interior rows by the plain kernel)
  int gmres_check = 20;     // iterations between two convergence tests (the reference: 20; DFB_GMRES_CHECK=1..20)
  int givens_defer = 1;     // one GPU: the scalar Givens step runs in the next multi-dot's tail (DFB_GIVENS_DEFER=0: in the update's)
  int halo_defer = 1;       // peer-memory mode: the mat-vec's first block raises the halo flag of the update before it (DFB_HALO_DEFER=0: the update's last block does)
  int graph = 1;
  int profile = 0;
  int assemble_mode = DFB_MODE_GATHER;
  int verbose = 0;
  int pc = 0;               // drop-in KrylovSolve: 0 the reference's block-Jacobi, 1 two-level Schur complement (pc2.cu)
  int pc_agg = 4, pc_degree = 10;
};
Options& options();

// physics / time-integration constants: reference src/assemble.cu:23-40, src/main.c:23-27
constexpr f64 kRHOC = 0.5;
constexpr f64 kDT = 5e-2;
constexpr f64 kALPHAM = (3.0 - kRHOC) / (1.0 + kRHOC);
constexpr f64 kALPHAF = 1.0 / (1.0 + kRHOC);
constexpr f64 kGAMMA = 0.5 + kALPHAM - kALPHAF;
constexpr f64 kRHO = 1.0e3;
constexpr f64 kCP = 1.0;
constexpr f64 kKAPPA = 0.66;
constexpr f64 kMU = 10.0 / 3.0;
// 4-point tet rule, reference src/assemble.cu:43-47
constexpr f64 kGW = 0.0416666666666667;
constexpr f64 kSHA = 0.5854101966249685;
constexpr f64 kSHB = 0.1381966011250105;

// ------------------------------------------------------------------------------------------------------------
// Peer-memory view of the data-parallel communicator (dist.cu builds it, the Krylov kernels in solve.cu consume it): the
// collectives of a GMRES iteration are FUSED into the compute kernels -- partial sums and halo values are stored straight
// into the peers' memory over NVLink (CUDA IPC mappings), flags carry monotonically increasing sequence numbers.
// The small all-reduce payloads (multi-dot coefficients, norms) travel in "LL" slots: a double is sent as two 8-byte words
// {32 data bits | 32-bit sequence tag}.  8-byte stores are single-copy atomic, so a word validates itself: the sender needs
// no fence and no separate flag, the receiver polls the slot until both tags match -- ONE NVLink traversal per exchange
// instead of three (data, fence round trip, flag).  Slots are double-buffered by the parity of the sequence number.
// Mailbox layout in 8-byte words, R = nranks:   A: multi-dot partials [2][R][128][2]     B: norm partials [2][R][2]
//                                               H: halo flags [R] (bulk data: stores + fence + flag)
// ------------------------------------------------------------------------------------------------------------
constexpr int P2P_MAXR = 8;
constexpr int P2P_ACAP = 128;
constexpr size_t P2P_MBOX_WORDS = 8192;   // mailbox size (words); the shared z vector follows it in the same allocation

struct P2PView {
  int rank, nranks;
  unsigned long long* mbox_local;
  unsigned long long* mbox_peer[P2P_MAXR];
  f64* z_local;
  f64* z_peer[P2P_MAXR];
  int n_nbr;
  int nbr[P2P_MAXR];
  int send_off[P2P_MAXR + 1];
  unsigned long long nbr_poff[P2P_MAXR];   // 3 * N_local of the neighbour (offset of p inside its z)
  const int* send_nodes;                   // [n_send] my local ids
  const int* remote_nodes;                 // [n_send] the same nodes in the neighbour's local numbering (its ghosts)
  unsigned* push_ctr;                      // last-block counter of the kernel that pushes the halo
  // the send lists inverted: boundary-owned node i (tgt_base <= i < tgt_base + tgt_n) goes to the targets
  // [tgt_ptr[i - tgt_base], tgt_ptr[i - tgt_base + 1]) = (neighbour slot q, local id on that neighbour)
  int tgt_base, tgt_n;
  const int* tgt_ptr;
  const int* tgt_q;
  const int* tgt_rid;
  f64* nrm_part;                           // this rank's partial norm of the newest column, parked by the update for a later kernel to publish
  f64* const* tgt_addr;                    // the same targets resolved once: address of the node's four entries in that neighbour's z
  // Bounded waits: every poll loop below gives up after `timeout_ns` (a peer died or returned early from the solve) and raises
  // *err; once *err is set every later wait of every later kernel returns at once, so the stream drains and the host sees
  // the word at its next synchronisation point (dfb_gmres_solve_pc returns DFB_ERR_PEER).  A hung peer can therefore cost
  // at most ~timeout per kernel in flight, never a GPU that has to be reset.
  unsigned* err;
  unsigned long long timeout_ns;
  // {all-reduce sequence, halo sequence} at the start of the current solve (device memory, set by the host once per solve):
  // the kernels of a solve receive OFFSETS from these (fixed for a given iteration index, so the launches can be replayed
  // from a CUDA graph) and form the tags as base + offset.
  const unsigned long long* seq_base;
};

// what dfb_comm_p2p_view() returns.  The sequence counters of the fused collectives belong to the COMMUNICATOR (one owner per
// dfb_comm): every workspace that runs over the same mailbox draws from the same monotonically increasing counters, so a
// mailbox tag or a halo flag left behind by an earlier solve (of this or of another workspace) can never satisfy a later wait.
struct P2PHandle {
  P2PView host;
  const P2PView* dev;
  unsigned long long* seq;    // all-reduce sequence (multi-dot / norm slots)
  unsigned long long* hseq;   // halo sequence
  unsigned* d_err;            // device error word (== host.err)
  unsigned long long* d_seq_base;   // device copy of {seq, hseq} at the start of the running solve (== host.seq_base)
};

__host__ __device__ inline size_t p2p_a_ll(int R, int par, int r, int j) { return (((size_t)par * R + r) * P2P_ACAP + j) * 2; }
__host__ __device__ inline size_t p2p_b_ll(int R, int par, int r) { return (size_t)4 * R * P2P_ACAP + ((size_t)par * R + r) * 2; }
__host__ __device__ inline size_t p2p_h_flag(int R, int r) { return (size_t)4 * R * P2P_ACAP + (size_t)4 * R + r; }
static_assert((size_t)4 * P2P_MAXR * P2P_ACAP + 5 * P2P_MAXR <= P2P_MBOX_WORDS, "mailbox too small");

#ifdef __CUDACC__
__device__ __forceinline__ void p2p_signal(unsigned long long* flag, unsigned long long seq) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
}
__device__ __forceinline__ void ll_store(unsigned long long* slot, f64 v, unsigned tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"((b & 0xffffffffull) | t), "l"((b >> 32) | t) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// true once the wait has to be abandoned: the error word is already raised, or this wait ran out of time (then it raises it)
__device__ __forceinline__ bool p2p_give_up(const P2PView* pv, unsigned long long t0) {
  if (*(volatile unsigned*)pv->err) return true;
  if (global_ns() - t0 > pv->timeout_ns) { atomicExch(pv->err, 1u); return true; }
  return false;
}
__device__ __forceinline__ f64 ll_load(const P2PView* pv, const unsigned long long* slot, unsigned tag) {
  unsigned long long w0, w1;
  unsigned long long t0 = 0;
  for (unsigned spin = 0;; spin++) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
    if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag) break;
    if ((spin & 255u) == 0u) {          // the clock and the error word are looked at every 256th poll only
      if (spin == 0u) t0 = global_ns();
      else if (p2p_give_up(pv, t0)) return 0.0;
    }
    __nanosleep(20);
  }
  return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}
__device__ __forceinline__ void p2p_wait(const P2PView* pv, const unsigned long long* flag, unsigned long long seq) {
  unsigned long long v;
  unsigned long long t0 = 0;
  for (unsigned spin = 0;; spin++) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= seq) break;
    if ((spin & 255u) == 0u) {
      if (spin == 0u) t0 = global_ns();
      else if (p2p_give_up(pv, t0)) return;
    }
    __nanosleep(32);
  }
}
#endif

}  // namespace dfb
