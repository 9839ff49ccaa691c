// newton.cu -- the vector work of DEDFlow's Newton / generalised-alpha driver around the hot path, fused.
//
// the caller side of the path.  Replaces (reference paths relative to /root/reference/src):
//   main.c:107-118, 232-246  alpha-level states: memset + 2 Daxpy + Dcopy (dwgalpha), Dcopy + 2 Daxpy + memset (wgalpha)
//                            -> ONE kernel (k_genalpha), 5 vector reads + 2 writes instead of 14 passes
//   main.c:127-130, 262-265  four cublasDnrm2 block norms, each a blocking host round trip -> ONE kernel (k_block_sumsq)
//   main.c:226               dwg -= dx                                                       -> k_newton_update
//   main.c:544-545           predictor: two Dscal                                            -> k_predict
//   main.c:559-563           corrector: four Daxpy + Dcopy                                   -> k_correct
// Every 6N vector is [u: N x 3 | p: N | phi: N | T: N]; the pressure slot follows the reference's special rules (defect D6:
// p lives in dwg, wgalpha's slot is zero, the predictor / corrector skip it).
#include <algorithm>

#include "common.cuh"

namespace dfb {

constexpr int NB_CHUNK = 296;

__global__ void __launch_bounds__(256) k_genalpha(size_t N, const f64* __restrict__ wgold, const f64* __restrict__ dwgold,
                                                  const f64* __restrict__ dwg, f64* __restrict__ wgalpha,
                                                  f64* __restrict__ dwgalpha) {
  const f64 f1a = 1.0 - kALPHAM, f1b = kALPHAM;
  const f64 f2a = kDT * kALPHAF * (1.0 - kGAMMA), f2b = kDT * kALPHAF * kGAMMA;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 6 * N; i += (size_t)gridDim.x * blockDim.x) {
    const f64 d = dwg[i], dold = dwgold[i];
    const bool pslot = i >= 3 * N && i < 4 * N;
    dwgalpha[i] = pslot ? d : fma(f1b, d, f1a * dold);                 // main.c:107-112
    wgalpha[i] = pslot ? 0.0 : fma(f2b, d, fma(f2a, dold, wgold[i]));   // main.c:114-118
  }
}

__global__ void __launch_bounds__(256) k_newton_update(size_t n, const f64* __restrict__ dx, f64* __restrict__ dwg) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dwg[i] -= dx[i];
}

__global__ void __launch_bounds__(256) k_predict(size_t N, f64* __restrict__ dwg) {
  const f64 fac = (kGAMMA - 1.0) / kGAMMA;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 6 * N; i += (size_t)gridDim.x * blockDim.x)
    if (i < 3 * N || i >= 4 * N) dwg[i] *= fac;
}

__global__ void __launch_bounds__(256) k_correct(size_t N, f64* __restrict__ wgold, f64* __restrict__ dwgold,
                                                 const f64* __restrict__ dwg) {
  const f64 c0 = kDT * (1.0 - kGAMMA), c1 = kDT * kGAMMA;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 6 * N; i += (size_t)gridDim.x * blockDim.x) {
    const f64 d = dwg[i];
    if (i < 3 * N || i >= 4 * N) wgold[i] = fma(c1, d, fma(c0, dwgold[i], wgold[i]));   // main.c:559-562 (two Daxpy)
    dwgold[i] = d;                                                                         // main.c:563
  }
}

// sums of squares of the four blocks [u | p | phi | T] over the first n_own nodes; two-stage, fixed order, one kernel
__global__ void __launch_bounds__(256) k_block_sumsq(size_t N, size_t n_own, const f64* __restrict__ F, f64* __restrict__ part,
                                                     unsigned* __restrict__ ctr, f64* __restrict__ out) {
  __shared__ f64 sm[8];
  __shared__ bool is_last;
  f64 acc[4] = {0.0, 0.0, 0.0, 0.0};
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = t0; i < 3 * n_own; i += stride) acc[0] = fma(F[i], F[i], acc[0]);
#pragma unroll
  for (int b = 1; b < 4; b++)
    for (size_t i = t0; i < n_own; i += stride) {
      const f64 v = F[(size_t)(2 + b) * N + i];
      acc[b] = fma(v, v, acc[b]);
    }
#pragma unroll
  for (int b = 0; b < 4; b++) {
    f64 v = acc[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      f64 r = 0.0;
      for (int w = 0; w < 8; w++) r += sm[w];
      part[(size_t)b * NB_CHUNK + blockIdx.x] = r;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(ctr, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 4) {
    f64 s = 0.0;
    for (int c = lane; c < (int)gridDim.x; c += 32) s += __ldcg(part + (size_t)warp * NB_CHUNK + c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[warp] = s;
  }
  if (threadIdx.x == 0) *ctr = 0u;
}

struct NormScratch {
  f64* part = nullptr;
  unsigned* ctr = nullptr;
  f64* out = nullptr;
};
static NormScratch g_norm;   // one device per process (one process per GPU)

static int norm_scratch() {
  if (g_norm.part) return DFB_OK;
  DFB_CUDA(cudaMalloc(&g_norm.part, sizeof(f64) * 4 * NB_CHUNK));
  DFB_CUDA(cudaMalloc(&g_norm.ctr, sizeof(unsigned)));
  DFB_CUDA(cudaMemset(g_norm.ctr, 0, sizeof(unsigned)));
  DFB_CUDA(cudaMalloc(&g_norm.out, sizeof(f64) * 4));
  return DFB_OK;
}

static int vec_grid(size_t n) { return (int)std::min<size_t>((n + 255) / 256, (size_t)8 * num_sms()); }

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_genalpha_stage(int N, const double* d_wgold, const double* d_dwgold, const double* d_dwg, double* d_wgalpha,
                       double* d_dwgalpha, void* stream) {
  if (N <= 0 || !d_wgold || !d_dwgold || !d_dwg || !d_wgalpha || !d_dwgalpha) { set_error("dfb_genalpha_stage: bad argument"); return DFB_ERR_ARG; }
  k_genalpha<<<vec_grid((size_t)6 * N), 256, 0, as_stream(stream)>>>((size_t)N, d_wgold, d_dwgold, d_dwg, d_wgalpha, d_dwgalpha);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_newton_update(int N, const double* d_dx, double* d_dwg, void* stream) {
  if (N <= 0 || !d_dx || !d_dwg) { set_error("dfb_newton_update: bad argument"); return DFB_ERR_ARG; }
  k_newton_update<<<vec_grid((size_t)6 * N), 256, 0, as_stream(stream)>>>((size_t)6 * N, d_dx, d_dwg);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_genalpha_predict(int N, double* d_dwg, void* stream) {
  if (N <= 0 || !d_dwg) { set_error("dfb_genalpha_predict: bad argument"); return DFB_ERR_ARG; }
  k_predict<<<vec_grid((size_t)6 * N), 256, 0, as_stream(stream)>>>((size_t)N, d_dwg);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_genalpha_correct(int N, double* d_wgold, double* d_dwgold, const double* d_dwg, void* stream) {
  if (N <= 0 || !d_wgold || !d_dwgold || !d_dwg) { set_error("dfb_genalpha_correct: bad argument"); return DFB_ERR_ARG; }
  k_correct<<<vec_grid((size_t)6 * N), 256, 0, as_stream(stream)>>>((size_t)N, d_wgold, d_dwgold, d_dwg);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_block_sumsq(int N, int n_own, const double* d_F, double* d_out4, void* stream) {
  if (N <= 0 || n_own <= 0 || n_own > N || !d_F || !d_out4) { set_error("dfb_block_sumsq: bad argument"); return DFB_ERR_ARG; }
  DFB_CHECK(norm_scratch());
  k_block_sumsq<<<NB_CHUNK, 256, 0, as_stream(stream)>>>((size_t)N, (size_t)n_own, d_F, g_norm.part, g_norm.ctr, d_out4);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_block_norms(int N, const double* d_F, double* h_out4, void* stream) {
  if (!h_out4) { set_error("dfb_block_norms: bad argument"); return DFB_ERR_ARG; }
  DFB_CHECK(norm_scratch());
  DFB_CHECK(dfb_block_sumsq(N, N, d_F, g_norm.out, stream));
  f64 h[4];
  DFB_CUDA(cudaMemcpyAsync(h, g_norm.out, sizeof(h), cudaMemcpyDeviceToHost, as_stream(stream)));
  DFB_CUDA(cudaStreamSynchronize(as_stream(stream)));
  for (int b = 0; b < 4; b++) h_out4[b] = sqrt(h[b]);
  return DFB_OK;
}

}  // extern "C"
