// color.cu -- element graph coloring (Jones-Plassmann-Luby) and color batches on the device.
//
// Replaces (reference paths relative to /root/reference/src):
//   color_impl.cu:185-192,225-237  weights = cuRAND XORWOW(seed 1234) u32 % (INT_MAX/2)
//   color_impl.cu:63-183           ColorElementJPLGPU: per round 4 kernels + CUB reduce + blocking 1-byte D2H
//                                  over ALL elements
//   color.c:63-67, color_impl.cu:240-255  GetMaxColor
//   Mesh.c:165-206, indexing.cu:92-103    per color thrust::count + thrust::copy_if (2*num_color passes)
//
// B200 design: the result of JPL on tie-free weights is the longest-path depth in the weight-oriented
// conflict DAG: color(e) = 1 + max{color(j) : j shares a vertex with e, w_j > w_e}.  A round therefore only
// needs the still-uncolored elements; we keep a compacted work list that shrinks geometrically, use immutable
// weights plus a separate color array (no in-place race, defect D2 -> ties are broken by element id so the
// coloring is always valid and deterministic), and read the remaining count once per round.
// Batches are one stable 8..9-bit radix sort of (color, element id).
#define QUALIFIERS static __forceinline__ __host__ __device__
#include <curand_kernel.h>
#include <cub/cub.cuh>

#include "common.cuh"
#include "plan.cuh"

namespace dfb {

constexpr int XORWOW_STREAMS = 4096;  // cuRAND default ordering: output n comes from subsequence n % 4096

__global__ void k_weights(int E, unsigned long long seed, int* __restrict__ weight) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= XORWOW_STREAMS || t >= E) return;
  curandStateXORWOW_t st;
  curand_init(seed, (unsigned long long)t, 0ULL, &st);
  const u32 ub = (u32)(2147483647 / 2);  // COLOR_RANDOM_UB - COLOR_RANDOM_LB, color_impl.cu:9-10
  for (i64 n = t; n < E; n += XORWOW_STREAMS) weight[n] = (int)(curand(&st) % ub);
}

__global__ void k_init_list(int E, int* __restrict__ list, int* __restrict__ color) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E) return;
  list[i] = i;
  color[i] = -1;
}

// one JPL round over the active list.  An element j blocks i iff j is uncoloured at the START of the round
// (color[j] < 0, or coloured in this very round: color[j] == c) and (w_j, j) > (w_i, i).
__global__ void k_jpl_round(int n_active, const int* __restrict__ list, int* __restrict__ next_list,
                            int* __restrict__ next_count, const int* __restrict__ ien,
                            const int* __restrict__ v2c_ptr, const int* __restrict__ v2c,
                            const int* __restrict__ weight, int* __restrict__ color, int c) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_active) return;
  int i = list[t];
  int wi = weight[i];
  bool is_max = true;
  int4 nd = *reinterpret_cast<const int4*>(ien + (size_t)i * 4);
  int nodes[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
  for (int a = 0; a < 4 && is_max; a++) {
    int s = v2c_ptr[nodes[a]], e = v2c_ptr[nodes[a] + 1];
    for (int k = s; k < e; k++) {
      int j = v2c[k] >> 2;
      if (j == i) continue;
      int cj = __ldcg(color + j);
      if (cj >= 0 && cj != c) continue;  // coloured in an earlier round: never blocks (color_impl.cu:87, values < 0)
      int wj = weight[j];
      if (wj > wi || (wj == wi && j > i)) { is_max = false; break; }
    }
  }
  if (is_max) {
    color[i] = c;
  } else {
    int pos = atomicAdd(next_count, 1);
    next_list[pos] = i;
  }
}

__global__ void k_iota(int n, int* __restrict__ a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_color_weights(int E, unsigned long long seed, int* d_weight, void* stream) {
  cudaStream_t st = as_stream(stream);
  if (E <= 0 || !d_weight) { set_error("dfb_color_weights: bad argument"); return DFB_ERR_ARG; }
  k_weights<<<XORWOW_STREAMS / 128, 128, 0, st>>>(E, seed, d_weight);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int dfb_color_jpl(int N, int E, const int* d_ien, const int* d_weight, int max_color, int* d_color, int* num_color,
                  void* stream) {
  cudaStream_t st = as_stream(stream);
  if (N <= 0 || E <= 0 || !d_ien || !d_weight || !d_color || !num_color) { set_error("dfb_color_jpl: bad argument"); return DFB_ERR_ARG; }
  // scoped device scratch: released on every return path (the DFB_CUDA / DFB_CHECK macros return early on errors)
  DevBuf<int> ptr, v2c, lists, counts;
  {
    int *p0 = nullptr, *p1 = nullptr;
    DFB_CHECK(build_v2c(N, E, d_ien, &p0, &p1, st));
    ptr.p = p0; v2c.p = p1;
  }
  DFB_CHECK(lists.alloc((size_t)E * 2));
  DFB_CHECK(counts.alloc((size_t)max_color + 1));
  DFB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * ((size_t)max_color + 1), st));
  k_init_list<<<ceil_div(E, 256), 256, 0, st>>>(E, lists, d_color);
  DFB_LAUNCH_CHECK();
  int n_active = E, c = 0;
  int* cur = lists;
  int* nxt = lists.p + E;
  for (; c < max_color && n_active > 0; c++) {
    k_jpl_round<<<ceil_div(n_active, 256), 256, 0, st>>>(n_active, cur, nxt, counts.p + c, d_ien, ptr, v2c, d_weight,
                                                        d_color, c);
    DFB_LAUNCH_CHECK();
    DFB_CUDA(cudaMemcpyAsync(&n_active, counts.p + c, sizeof(int), cudaMemcpyDeviceToHost, st));
    DFB_CUDA(cudaStreamSynchronize(st));
    int* t = cur; cur = nxt; nxt = t;
  }
  if (n_active > 0) { set_error("dfb_color_jpl: %d elements uncoloured after %d rounds", n_active, max_color); return DFB_ERR_COLOR; }
  *num_color = c;
  return DFB_OK;
}

int dfb_color_batches(int E, const int* d_color, int num_color, int* h_batch_offset, int* d_batch_ind, void* stream) {
  cudaStream_t st = as_stream(stream);
  if (E <= 0 || num_color <= 0 || !d_color || !h_batch_offset || !d_batch_ind) { set_error("dfb_color_batches: bad argument"); return DFB_ERR_ARG; }
  DevBuf<int> keys_out, ids, hist;
  DevBuf<char> tmp;
  DFB_CHECK(keys_out.alloc((size_t)E));
  DFB_CHECK(ids.alloc((size_t)E));
  DFB_CHECK(hist.alloc((size_t)num_color + 1));
  k_iota<<<ceil_div(E, 256), 256, 0, st>>>(E, ids);
  DFB_LAUNCH_CHECK();
  int bits = 1;
  while ((1 << bits) < num_color) bits++;
  size_t tmp_bytes = 0, tmp2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_color, keys_out.p, ids.p, d_batch_ind, E, 0, bits, st);
  cub::DeviceHistogram::HistogramEven(nullptr, tmp2, d_color, hist.p, num_color + 1, 0, num_color, E, st);
  if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
  DFB_CHECK(tmp.alloc(tmp_bytes));
  cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, d_color, keys_out.p, ids.p, d_batch_ind, E, 0, bits, st);  // stable
  DFB_LAUNCH_CHECK();
  cub::DeviceHistogram::HistogramEven(tmp.p, tmp_bytes, d_color, hist.p, num_color + 1, 0, num_color, E, st);
  DFB_LAUNCH_CHECK();
  DFB_CUDA(cudaMemcpyAsync(h_batch_offset + 1, hist.p, sizeof(int) * (size_t)num_color, cudaMemcpyDeviceToHost, st));
  DFB_CUDA(cudaStreamSynchronize(st));
  h_batch_offset[0] = 0;
  for (int c = 0; c < num_color; c++) h_batch_offset[c + 1] += h_batch_offset[c];
  return DFB_OK;
}

}  // extern "C"
