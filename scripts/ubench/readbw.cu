// readbw.cu -- what a read-only streaming kernel can reach on this GPU, as a function of footprint and mechanism:
//   (a) register loads: grid-stride 16-byte loads, U independent loads per thread per trip;
//   (b) TMA ring: one producer thread per CTA issues 1-D bulk copies into a ring of shared-memory stages, consumers only touch
//       one word per stage (the bytes never pass through registers).
// The result is the ceiling the SpMV / multi-dot kernels are judged against beside MEASURED_PEAKS.json's copy bandwidth.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o readbw readbw.cu && ./readbw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include <algorithm>
#include "../../dedflow_b200/csrc/tma.cuh"
using namespace dfb;

template <int U>
__global__ void __launch_bounds__(256) k_ldg(const double2* __restrict__ p, size_t n2, double* out) {
  const size_t stride = (size_t)gridDim.x * 256;
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  double acc = 0.0;
  for (; i + (U - 1) * stride < n2; i += U * stride) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; u++) acc += v[u].x + v[u].y;
  }
  if (acc == 1.2345e-300) out[0] = acc;
}

template <int STAGES, int STAGE_BYTES>
__global__ void __launch_bounds__(160, 1) k_tma(const char* __restrict__ p, size_t bytes, double* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { tma::mbar_init(full + s, 1); tma::mbar_init(empty + s, 4); }
    tma::fence_barrier_init();
  }
  __syncthreads();
  const size_t ntile = bytes / STAGE_BYTES;
  if (warp == 4) {
    if (lane) return;
    size_t t = blockIdx.x;
    for (int j = 0; t < ntile; j++, t += gridDim.x) {
      const int s = j % STAGES;
      tma::mbar_wait(empty + s, ((j / STAGES) & 1) ^ 1);
      tma::mbar_arrive_expect_tx(full + s, STAGE_BYTES);
      tma::bulk_g2s(sm + (size_t)s * STAGE_BYTES, p + t * STAGE_BYTES, STAGE_BYTES, full + s);
    }
    return;
  }
  double acc = 0.0;
  size_t t = blockIdx.x;
  for (int j = 0; t < ntile; j++, t += gridDim.x) {
    const int s = j % STAGES;
    tma::mbar_wait(full + s, (j / STAGES) & 1);
    acc += reinterpret_cast<const double*>(sm + (size_t)s * STAGE_BYTES)[threadIdx.x];
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(empty + s);
  }
  if (acc == 1.2345e-300) out[0] = acc;
}

static float time_it(void (*fn)(void*), void* ctx, char* flush, size_t flush_bytes, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  std::vector<float> ts;
  for (int r = 0; r < reps; r++) {
    if (flush) cudaMemsetAsync(flush, r, flush_bytes);
    cudaEventRecord(a);
    fn(ctx);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    ts.push_back(ms);
  }
  std::sort(ts.begin(), ts.end());
  return ts[ts.size() / 2];
}

struct Ctx { const char* p; size_t bytes; double* out; int grid; };

int main() {
  const size_t sizes[] = {44u << 20, 350u << 20, 1400u << 20};
  char* buf; double* out; char* flush;
  const size_t maxb = 1400u << 20, fb = 512u << 20;
  cudaMalloc(&buf, maxb); cudaMalloc(&out, 64); cudaMalloc(&flush, fb);
  cudaMemset(buf, 1, maxb);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (size_t bytes : sizes) {
    Ctx c{buf, bytes, out, 0};
    printf("---- %zu MB (L2 flushed before every launch) ----\n", bytes >> 20);
    for (int occ : {2, 4, 8}) {
      c.grid = sms * occ;
      float t1 = time_it([](void* v) { Ctx* c = (Ctx*)v; k_ldg<1><<<c->grid, 256>>>((const double2*)c->p, c->bytes / 16, c->out); }, &c, flush, fb, 15);
      float t4 = time_it([](void* v) { Ctx* c = (Ctx*)v; k_ldg<4><<<c->grid, 256>>>((const double2*)c->p, c->bytes / 16, c->out); }, &c, flush, fb, 15);
      float t8 = time_it([](void* v) { Ctx* c = (Ctx*)v; k_ldg<8><<<c->grid, 256>>>((const double2*)c->p, c->bytes / 16, c->out); }, &c, flush, fb, 15);
      printf("ldg   %d CTAs/SM x 256 thr: U=1 %7.1f us %6.0f GB/s | U=4 %7.1f us %6.0f GB/s | U=8 %7.1f us %6.0f GB/s\n", occ, t1 * 1e3,
             bytes / t1 / 1e6, t4 * 1e3, bytes / t4 / 1e6, t8 * 1e3, bytes / t8 / 1e6);
    }
    c.grid = sms;
    {
      constexpr int SB = 32768, ST = 6;
      cudaFuncSetAttribute(k_tma<ST, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * SB + 128);
      float t = time_it([](void* v) { Ctx* c = (Ctx*)v; k_tma<ST, SB><<<c->grid, 160, ST * SB + 128>>>(c->p, c->bytes, c->out); }, &c, flush, fb, 15);
      printf("tma   6 x 32 KB stages/SM: %7.1f us %6.0f GB/s\n", t * 1e3, bytes / t / 1e6);
    }
    {
      constexpr int SB = 65536, ST = 3;
      cudaFuncSetAttribute(k_tma<ST, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * SB + 128);
      float t = time_it([](void* v) { Ctx* c = (Ctx*)v; k_tma<ST, SB><<<c->grid, 160, ST * SB + 128>>>(c->p, c->bytes, c->out); }, &c, flush, fb, 15);
      printf("tma   3 x 64 KB stages/SM: %7.1f us %6.0f GB/s\n", t * 1e3, bytes / t / 1e6);
    }
    {
      constexpr int SB = 8192, ST = 8;
      cudaFuncSetAttribute(k_tma<ST, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * SB + 128);
      float t = time_it([](void* v) { Ctx* c = (Ctx*)v; k_tma<ST, SB><<<c->grid, 160, ST * SB + 128>>>(c->p, c->bytes, c->out); }, &c, flush, fb, 15);
      printf("tma   8 x  8 KB stages/SM: %7.1f us %6.0f GB/s\n", t * 1e3, bytes / t / 1e6);
    }
    {
      constexpr int SB = 16384, ST = 12;
      cudaFuncSetAttribute(k_tma<ST, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * SB + 256);
      float t = time_it([](void* v) { Ctx* c = (Ctx*)v; k_tma<ST, SB><<<c->grid, 160, ST * SB + 256>>>(c->p, c->bytes, c->out); }, &c, flush, fb, 15);
      printf("tma  12 x 16 KB stages/SM: %7.1f us %6.0f GB/s\n", t * 1e3, bytes / t / 1e6);
    }
    // back to back (no flush): what L2 residency gives at this footprint
    c.grid = sms * 8;
    float tb = time_it([](void* v) { Ctx* c = (Ctx*)v; k_ldg<4><<<c->grid, 256>>>((const double2*)c->p, c->bytes / 16, c->out); }, &c, nullptr, 0, 15);
    printf("ldg U=4 8 CTAs/SM, back to back (no flush): %7.1f us %6.0f GB/s\n", tb * 1e3, bytes / tb / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
