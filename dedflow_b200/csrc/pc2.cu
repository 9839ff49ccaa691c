// pc2.cu -- a stronger preconditioner for the field-split flow system: block lower-triangular ("SIMPLE"-type) with an
// ADDITIVE TWO-LEVEL approximation of the pressure Schur complement.  This fills the slot the reference reserves for AMGX on the
// pressure block (src/pc.c:160-235, src/krylov.c:392-453: `PCCreateAMGX(A11, cfg)` is commented out and compiled out there), so
// it is NEW WORK, opt-in (dfb_gmres_set_pc2 / DFB_PC=schur2): the default stays the reference's block-Jacobi and every parity
// test is untouched.  Checked against a numpy/scipy restatement (oracle/pc2_oracle.py).
//
//   u = D^-1 r_u                                   D = the 3x3 diagonal blocks of A00 (true inverse, not defect D3's transpose)
//   r~ = r_p - A10 u                               lower-triangular coupling: one pass over A10 (3 of the 16 value streams)
//   p = w dS^-1 r~  +  P Cheb_m(Sc, P^T r~)        S = A11 - A10 D^-1 A01 (never formed), dS = diag(S),
//                                                  P = piecewise-constant prolongation from aggregates of ~a^3 nodes (coordinate
//                                                  cells), Sc = P^T S P (Galerkin, rebuilt per solve), m Chebyshev steps on it
// Why this shape (prototype on the first Newton system of a time step, m=28: 24,389 nodes): block-Jacobi 84 GMRES iterations to
// rtol 1e-4, exact A11 solve 116+ (A11 alone is the wrong pressure operator), Chebyshev on S unstable (S is non-symmetric), this
// preconditioner 20-24 -- with NO fine-level application of S: per application it reads A10 once plus a coarse matrix of
// ~125 entries per aggregate.  Single GPU (the halo of the coarse correction is not implemented).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace dfb {

constexpr int PC2_MAXAGG = 32;   // most distinct aggregates among a node and its neighbours

__device__ __forceinline__ int find_col(const int* __restrict__ ci, int start, int len, int target) {
  int lo = 0, hi = len;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ci[start + mid] < target) lo = mid + 1; else hi = mid;
  }
  return (lo < len && ci[start + lo] == target) ? lo : -1;
}

// Dinv[9 i + 3 r + c] = (B_i^-1)(r, c), B_i = diagonal block of A00 (row-major)
__global__ void k_pc2_dinv(int N, const int* __restrict__ rp, const int* __restrict__ ci, const f64* __restrict__ A00,
                           f64* __restrict__ Dinv, int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int start = rp[i], len = rp[i + 1] - start;
  const int k = find_col(ci, start, len, i);
  f64* out = Dinv + (size_t)i * 9;
  if (k < 0) { atomicExch(bad, 1); for (int t = 0; t < 9; t++) out[t] = (t % 4 == 0) ? 1.0 : 0.0; return; }
  f64 B[3][3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) B[r][c] = A00[(size_t)start * 9 + (size_t)k * 3 + (size_t)r * len * 3 + c];
  const f64 c00 = B[1][1] * B[2][2] - B[1][2] * B[2][1], c01 = B[1][2] * B[2][0] - B[1][0] * B[2][2],
            c02 = B[1][0] * B[2][1] - B[1][1] * B[2][0];
  const f64 det = B[0][0] * c00 + B[0][1] * c01 + B[0][2] * c02;
  if (det == 0.0 || !isfinite(det)) { atomicExch(bad, 1); for (int t = 0; t < 9; t++) out[t] = (t % 4 == 0) ? 1.0 : 0.0; return; }
  const f64 id = 1.0 / det;
  out[0] = c00 * id; out[3] = c01 * id; out[6] = c02 * id;
  out[1] = (B[0][2] * B[2][1] - B[0][1] * B[2][2]) * id;
  out[4] = (B[0][0] * B[2][2] - B[0][2] * B[2][0]) * id;
  out[7] = (B[0][1] * B[2][0] - B[0][0] * B[2][1]) * id;
  out[2] = (B[0][1] * B[1][2] - B[0][2] * B[1][1]) * id;
  out[5] = (B[0][2] * B[1][0] - B[0][0] * B[1][2]) * id;
  out[8] = (B[0][0] * B[1][1] - B[0][1] * B[1][0]) * id;
}

// dSinv[i] = 1 / (A11_ii - sum_k A10[i, k] D_k^-1 A01[k, i])
__global__ void k_pc2_diagS(int N, const int* __restrict__ rp, const int* __restrict__ ci, const f64* __restrict__ A01,
                            const f64* __restrict__ A10, const f64* __restrict__ A11, const f64* __restrict__ Dinv,
                            f64* __restrict__ dSinv, int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int start = rp[i], len = rp[i + 1] - start;
  f64 s = 0.0, diag = 0.0;
  for (int p = 0; p < len; p++) {
    const int k = ci[start + p];
    if (k == i) diag = A11[start + p];
    const int sk = rp[k], lk = rp[k + 1] - sk;
    const int pos = find_col(ci, sk, lk, i);
    if (pos < 0) continue;   // structurally non-symmetric entry: no transpose partner
    const f64* a10 = A10 + (size_t)start * 3 + (size_t)p * 3;
    const f64* D = Dinv + (size_t)k * 9;
    f64 a01[3];
    for (int r = 0; r < 3; r++) a01[r] = A01[(size_t)sk * 3 + (size_t)r * lk + pos];
    for (int r = 0; r < 3; r++) s += a10[r] * (D[3 * r] * a01[0] + D[3 * r + 1] * a01[1] + D[3 * r + 2] * a01[2]);
  }
  const f64 dS = diag - s;
  if (dS == 0.0 || !isfinite(dS)) { atomicExch(bad, 1); dSinv[i] = 1.0; return; }
  dSinv[i] = 1.0 / dS;
}

__device__ __forceinline__ int coarse_slot(const int* __restrict__ crp, const int* __restrict__ cci, int I, int J) {
  const int s = crp[I];
  const int p = find_col(cci, s, crp[I + 1] - s, J);
  return p < 0 ? -1 : s + p;
}

// Sc += P^T A11 P
__global__ void k_pc2_coarse_a11(int N, const int* __restrict__ rp, const int* __restrict__ ci, const f64* __restrict__ A11,
                                 const int* __restrict__ agg, const int* __restrict__ crp, const int* __restrict__ cci,
                                 f64* __restrict__ Sc, int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int I = agg[i];
  for (int p = rp[i]; p < rp[i + 1]; p++) {
    const int slot = coarse_slot(crp, cci, I, agg[ci[p]]);
    if (slot < 0) { atomicExch(bad, 2); continue; }
    atomicAdd(Sc + slot, A11[p]);
  }
}

// Sc -= (P^T A10) D^-1 (A01 P): fine node k contributes -h_I . D_k^-1 g_J for every pair of aggregates (I, J) among its neighbours,
// h_I = sum_{i in I} A10[i, k]  (row i, column block k),  g_J = sum_{j in J} A01[k, j]  (row block k, column j)
__global__ void __launch_bounds__(128) k_pc2_coarse_schur(int N, const int* __restrict__ rp, const int* __restrict__ ci,
                                                          const f64* __restrict__ A01, const f64* __restrict__ A10,
                                                          const f64* __restrict__ Dinv, const int* __restrict__ agg,
                                                          const int* __restrict__ crp, const int* __restrict__ cci,
                                                          f64* __restrict__ Sc, int* __restrict__ bad) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  int ids[PC2_MAXAGG];
  f64 h[PC2_MAXAGG][3], g[PC2_MAXAGG][3];
  int na = 0;
  const int sk = rp[k], lk = rp[k + 1] - sk;
  for (int p = 0; p < lk; p++) {
    const int i = ci[sk + p];
    const int I = agg[i];
    int idx = 0;
    while (idx < na && ids[idx] != I) idx++;
    if (idx == na) {
      if (na == PC2_MAXAGG) { atomicExch(bad, 3); continue; }
      ids[na] = I;
      for (int r = 0; r < 3; r++) { h[na][r] = 0.0; g[na][r] = 0.0; }
      na++;
    }
    for (int r = 0; r < 3; r++) g[idx][r] += A01[(size_t)sk * 3 + (size_t)r * lk + p];
    const int si = rp[i], li = rp[i + 1] - si;
    const int pos = find_col(ci, si, li, k);
    if (pos >= 0)
      for (int r = 0; r < 3; r++) h[idx][r] += A10[(size_t)si * 3 + (size_t)pos * 3 + r];
  }
  const f64* D = Dinv + (size_t)k * 9;
  for (int b = 0; b < na; b++) {
    f64 v[3];
    for (int r = 0; r < 3; r++) v[r] = D[3 * r] * g[b][0] + D[3 * r + 1] * g[b][1] + D[3 * r + 2] * g[b][2];
    for (int a = 0; a < na; a++) {
      const int slot = coarse_slot(crp, cci, ids[a], ids[b]);
      if (slot < 0) { atomicExch(bad, 2); continue; }
      atomicAdd(Sc + slot, -(h[a][0] * v[0] + h[a][1] * v[1] + h[a][2] * v[2]));
    }
  }
}

// coarse diagonal + Gershgorin bound of the Jacobi-scaled coarse operator (an upper bound of its spectral radius: the
// Chebyshev interval is [bound / ratio, bound])
__global__ void k_pc2_coarse_diag(int Nc, const int* __restrict__ crp, const int* __restrict__ cci, const f64* __restrict__ Sc,
                                  f64* __restrict__ dScinv, unsigned long long* __restrict__ lam_bits, int* __restrict__ bad) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= Nc) return;
  f64 diag = 0.0, sum = 0.0;
  for (int p = crp[I]; p < crp[I + 1]; p++) {
    const f64 v = Sc[p];
    sum += fabs(v);
    if (cci[p] == I) diag = v;
  }
  if (diag == 0.0 || !isfinite(diag)) { atomicExch(bad, 4); dScinv[I] = 1.0; return; }
  dScinv[I] = 1.0 / diag;
  atomicMax(lam_bits, (unsigned long long)__double_as_longlong(sum / fabs(diag)));   // positive doubles order like their bits
}

// ---------------------------------------------------------------------------------------------------------------- apply
// z_u = D^-1 w_u  (interleaved vectors: v[4 i + c])
__global__ void k_pc2_u(int N, const f64* __restrict__ Dinv, const f64* __restrict__ w, f64* __restrict__ z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double2* w2 = reinterpret_cast<const double2*>(w) + (size_t)i * 2;
  const double2 lo = w2[0], hi = w2[1];
  const f64* D = Dinv + (size_t)i * 9;
  f64* zi = z + (size_t)i * 4;
  zi[0] = D[0] * lo.x + D[1] * lo.y + D[2] * hi.x;
  zi[1] = D[3] * lo.x + D[4] * lo.y + D[5] * hi.x;
  zi[2] = D[6] * lo.x + D[7] * lo.y + D[8] * hi.x;
}

// rt_i = w_p(i) - (A10 z_u)_i.   8 lanes per nodal row over its 3 len contiguous A10 values.
__global__ void __launch_bounds__(256) k_pc2_rp(int N, const int* __restrict__ rp, const int* __restrict__ ci,
                                                const f64* __restrict__ A10, const f64* __restrict__ w, const f64* __restrict__ z,
                                                f64* __restrict__ rt) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(gt >> 3), lane = (int)(threadIdx.x & 7);
  f64 s = 0.0;
  if (row < N) {
    const int start = rp[row], len3 = 3 * (rp[row + 1] - start);
    const f64* a = A10 + (size_t)start * 3;
    for (int t = lane; t < len3; t += 8) {
      const int p = t / 3, jj = t - 3 * p;
      s = fma(__ldcs(a + t), __ldg(z + (size_t)__ldg(ci + start + p) * 4 + jj), s);
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (row < N && lane == 0) {
    const f64 r = w[(size_t)row * 4 + 3] - s;
    rt[row] = r;
  }
}
// z_p(i) = omega rt_i / dS_i (its own pass: the kernel above is still gathering z of other rows)
__global__ void k_pc2_zp(int N, const f64* __restrict__ rt, const f64* __restrict__ dSinv, f64 omega, f64* __restrict__ z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) z[(size_t)i * 4 + 3] = omega * rt[i] * dSinv[i];
}

struct ChebCoef { f64 theta, delta, sigma; };
__device__ __forceinline__ ChebCoef cheb_coef(const unsigned long long* lam_bits, f64 ratio) {
  const f64 lmx = __longlong_as_double((long long)*lam_bits), lmn = lmx / ratio;
  ChebCoef c;
  c.theta = 0.5 * (lmx + lmn);
  c.delta = 0.5 * (lmx - lmn);
  c.sigma = c.theta / c.delta;
  return c;
}

// rc = P^T rt (fixed member order: deterministic); Chebyshev start: res = rc, ec = 0, d0 = dSc^-1 rc / theta
__global__ void k_pc2_restrict(int Nc, const int* __restrict__ agg_ptr, const int* __restrict__ agg_nodes, const f64* __restrict__ rt,
                               const f64* __restrict__ dScinv, const unsigned long long* __restrict__ lam_bits, f64 ratio,
                               f64* __restrict__ res, f64* __restrict__ ec, f64* __restrict__ d0) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= Nc) return;
  f64 s = 0.0;
  for (int p = agg_ptr[I]; p < agg_ptr[I + 1]; p++) s += rt[agg_nodes[p]];
  const ChebCoef c = cheb_coef(lam_bits, ratio);
  res[I] = s;
  ec[I] = 0.0;
  d0[I] = dScinv[I] * s / c.theta;
}

// Chebyshev step k -> k + 1 on the coarse system: ec += d_k; res -= Sc d_k; d_{k+1} = rho_{k+1} rho_k d_k + 2 rho_{k+1}/delta dSc^-1 res
__global__ void k_pc2_sweep(int Nc, int k, const int* __restrict__ crp, const int* __restrict__ cci, const f64* __restrict__ Sc,
                            const f64* __restrict__ dScinv, const unsigned long long* __restrict__ lam_bits, f64 ratio,
                            f64* __restrict__ res, f64* __restrict__ ec, const f64* __restrict__ d_old, f64* __restrict__ d_new) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= Nc) return;
  const ChebCoef c = cheb_coef(lam_bits, ratio);
  f64 rho = 1.0 / c.sigma;
  for (int t = 0; t < k; t++) rho = 1.0 / (2.0 * c.sigma - rho);
  const f64 rho_new = 1.0 / (2.0 * c.sigma - rho);
  f64 t = 0.0;
  for (int p = crp[I]; p < crp[I + 1]; p++) t = fma(Sc[p], d_old[cci[p]], t);
  const f64 dI = d_old[I];
  const f64 r = res[I] - t;
  res[I] = r;
  ec[I] += dI;
  d_new[I] = rho_new * rho * dI + 2.0 * rho_new / c.delta * (dScinv[I] * r);
}

// z_p += (P ec)_i  (ec + the last Chebyshev direction)
__global__ void k_pc2_prolong(int N, const int* __restrict__ agg, const f64* __restrict__ ec, const f64* __restrict__ d_last,
                              f64* __restrict__ z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int I = agg[i];
  z[(size_t)i * 4 + 3] += ec[I] + d_last[I];
}

// ABI-layout (6N) <-> interleaved, for the standalone entry point
__global__ void k_pc2_pack(int N, const f64* __restrict__ x, f64* __restrict__ w) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)4 * N) return;
  const size_t i = t >> 2;
  const int c = (int)(t & 3);
  w[t] = c < 3 ? x[i * 3 + c] : x[(size_t)3 * N + i];
}
__global__ void k_pc2_unpack(int N, const f64* __restrict__ z, f64* __restrict__ y) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)4 * N) return;
  const size_t i = t >> 2;
  const int c = (int)(t & 3);
  if (c < 3) y[i * 3 + c] = z[t]; else y[(size_t)3 * N + i] = z[t];
}

}  // namespace dfb

using namespace dfb;

struct dfb_pc2 {
  int N = 0, Nc = 0, cnnz = 0, deg = 10;
  f64 ratio = 30.0, omega = 0.7;
  const int *rp = nullptr, *ci = nullptr;   // borrowed nodal pattern
  int *agg = nullptr, *agg_ptr = nullptr, *agg_nodes = nullptr, *crp = nullptr, *cci = nullptr, *bad = nullptr;
  f64 *Dinv = nullptr, *dSinv = nullptr, *Sc = nullptr, *dScinv = nullptr, *rt = nullptr, *res = nullptr, *ec = nullptr,
      *d0 = nullptr, *d1 = nullptr, *wtmp = nullptr, *ztmp = nullptr;
  unsigned long long* lam = nullptr;
  bool ready = false;
};

// launches of one application on interleaved vectors (w: input, z: output, both 4N); used by the solver and the entry point
int pc2_apply_aos(const dfb_pc2* P, const f64* A10, const f64* w, f64* z, cudaStream_t st) {
  const int N = P->N, Nc = P->Nc;
  k_pc2_u<<<ceil_div(N, 128), 128, 0, st>>>(N, P->Dinv, w, z);
  DFB_LAUNCH_CHECK();
  k_pc2_rp<<<ceil_div((i64)N * 8, 256), 256, 0, st>>>(N, P->rp, P->ci, A10, w, z, P->rt);
  DFB_LAUNCH_CHECK();
  k_pc2_zp<<<ceil_div(N, 256), 256, 0, st>>>(N, P->rt, P->dSinv, P->omega, z);
  DFB_LAUNCH_CHECK();
  k_pc2_restrict<<<ceil_div(Nc, 128), 128, 0, st>>>(Nc, P->agg_ptr, P->agg_nodes, P->rt, P->dScinv, P->lam, P->ratio, P->res, P->ec, P->d0);
  DFB_LAUNCH_CHECK();
  f64 *dold = P->d0, *dnew = P->d1;
  for (int k = 0; k + 1 < P->deg; k++) {
    k_pc2_sweep<<<ceil_div(Nc, 128), 128, 0, st>>>(Nc, k, P->crp, P->cci, P->Sc, P->dScinv, P->lam, P->ratio, P->res, P->ec, dold, dnew);
    DFB_LAUNCH_CHECK();
    std::swap(dold, dnew);
  }
  k_pc2_prolong<<<ceil_div(N, 256), 256, 0, st>>>(N, P->agg, P->ec, dold, z);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" {

void dfb_pc2_destroy(dfb_pc2* P) {
  if (!P) return;
  cudaFree(P->agg); cudaFree(P->agg_ptr); cudaFree(P->agg_nodes); cudaFree(P->crp); cudaFree(P->cci); cudaFree(P->bad);
  cudaFree(P->Dinv); cudaFree(P->dSinv); cudaFree(P->Sc); cudaFree(P->dScinv); cudaFree(P->rt); cudaFree(P->res); cudaFree(P->ec);
  cudaFree(P->d0); cudaFree(P->d1); cudaFree(P->wtmp); cudaFree(P->ztmp); cudaFree(P->lam);
  delete P;
}

// Static part, once per mesh: aggregates = cells of agg_cells^3 average node spacings (any mesh, any numbering), their member
// lists, and the coarse pattern (pairs of aggregates within two fine edges: the square of the aggregated adjacency).
int dfb_pc2_create(dfb_pc2** out, int N, const int* d_row_ptr, const int* d_col_ind, const double* d_xg, int agg_cells,
                   int cheb_degree, void* stream) {
  cudaStream_t st = as_stream(stream);
  if (!out || N <= 0 || !d_row_ptr || !d_col_ind || !d_xg || agg_cells < 2 || agg_cells > 16 || cheb_degree < 1 || cheb_degree > 64) {
    set_error("dfb_pc2_create: bad argument");
    return DFB_ERR_ARG;
  }
  std::vector<int> rp((size_t)N + 1);
  DFB_CUDA(cudaMemcpyAsync(rp.data(), d_row_ptr, sizeof(int) * ((size_t)N + 1), cudaMemcpyDeviceToHost, st));
  DFB_CUDA(cudaStreamSynchronize(st));
  const size_t Z = (size_t)rp[N];
  std::vector<int> ci(Z);
  std::vector<f64> xg((size_t)3 * N);
  DFB_CUDA(cudaMemcpyAsync(ci.data(), d_col_ind, sizeof(int) * Z, cudaMemcpyDeviceToHost, st));
  DFB_CUDA(cudaMemcpyAsync(xg.data(), d_xg, sizeof(f64) * 3 * (size_t)N, cudaMemcpyDeviceToHost, st));
  DFB_CUDA(cudaStreamSynchronize(st));
  // ---- aggregates
  f64 lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < N; i++)
    for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], xg[(size_t)3 * i + d]); hi[d] = std::max(hi[d], xg[(size_t)3 * i + d]); }
  f64 vol = 1.0;
  int nd = 0;
  for (int d = 0; d < 3; d++)
    if (hi[d] > lo[d]) { vol *= hi[d] - lo[d]; nd++; }
  const f64 h = nd ? pow(vol / (f64)N, 1.0 / nd) : 1.0;
  const f64 cell = agg_cells * h;
  long long nc[3];
  for (int d = 0; d < 3; d++) nc[d] = std::max(1ll, (long long)ceil((hi[d] - lo[d]) / cell - 1e-9));
  std::vector<long long> key((size_t)N);
  for (int i = 0; i < N; i++) {
    long long c[3];
    for (int d = 0; d < 3; d++) c[d] = std::min(nc[d] - 1, (long long)floor((xg[(size_t)3 * i + d] - lo[d]) / cell));
    key[i] = c[0] + nc[0] * (c[1] + nc[1] * c[2]);
  }
  std::vector<long long> uniq(key);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  const int Nc = (int)uniq.size();
  std::vector<int> agg((size_t)N), agg_ptr((size_t)Nc + 1, 0), agg_nodes((size_t)N);
  for (int i = 0; i < N; i++) {
    agg[i] = (int)(std::lower_bound(uniq.begin(), uniq.end(), key[i]) - uniq.begin());
    agg_ptr[(size_t)agg[i] + 1]++;
  }
  for (int I = 0; I < Nc; I++) agg_ptr[(size_t)I + 1] += agg_ptr[I];
  {
    std::vector<int> fill(agg_ptr.begin(), agg_ptr.end() - 1);
    for (int i = 0; i < N; i++) agg_nodes[(size_t)fill[agg[i]]++] = i;   // ascending node id inside an aggregate
  }
  // ---- coarse pattern: C1 = aggregated adjacency, pattern = C1 * C1
  std::vector<std::vector<int>> c1((size_t)Nc);
  for (int i = 0; i < N; i++) {
    std::vector<int>& row = c1[agg[i]];
    for (int p = rp[i]; p < rp[i + 1]; p++) {
      const int J = agg[ci[p]];
      if (row.empty() || row.back() != J) row.push_back(J);   // cheap de-duplication of runs; exact one below
    }
  }
  for (auto& row : c1) { std::sort(row.begin(), row.end()); row.erase(std::unique(row.begin(), row.end()), row.end()); }
  std::vector<int> crp((size_t)Nc + 1, 0), cci;
  {
    std::vector<int> tmp;
    for (int I = 0; I < Nc; I++) {
      tmp.clear();
      for (int K : c1[I]) tmp.insert(tmp.end(), c1[K].begin(), c1[K].end());
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      cci.insert(cci.end(), tmp.begin(), tmp.end());
      crp[(size_t)I + 1] = (int)cci.size();
    }
  }
  dfb_pc2* P = new dfb_pc2();
  P->N = N; P->Nc = Nc; P->cnnz = (int)cci.size(); P->deg = cheb_degree;
  P->rp = d_row_ptr; P->ci = d_col_ind;
  struct { void** p; size_t bytes; const void* src; } up[] = {
      {(void**)&P->agg, sizeof(int) * (size_t)N, agg.data()},
      {(void**)&P->agg_ptr, sizeof(int) * ((size_t)Nc + 1), agg_ptr.data()},
      {(void**)&P->agg_nodes, sizeof(int) * (size_t)N, agg_nodes.data()},
      {(void**)&P->crp, sizeof(int) * ((size_t)Nc + 1), crp.data()},
      {(void**)&P->cci, sizeof(int) * cci.size(), cci.data()},
      {(void**)&P->bad, sizeof(int), nullptr},
      {(void**)&P->Dinv, sizeof(f64) * 9 * (size_t)N, nullptr},
      {(void**)&P->dSinv, sizeof(f64) * (size_t)N, nullptr},
      {(void**)&P->Sc, sizeof(f64) * cci.size(), nullptr},
      {(void**)&P->dScinv, sizeof(f64) * (size_t)Nc, nullptr},
      {(void**)&P->rt, sizeof(f64) * (size_t)N, nullptr},
      {(void**)&P->res, sizeof(f64) * (size_t)Nc, nullptr},
      {(void**)&P->ec, sizeof(f64) * (size_t)Nc, nullptr},
      {(void**)&P->d0, sizeof(f64) * (size_t)Nc, nullptr},
      {(void**)&P->d1, sizeof(f64) * (size_t)Nc, nullptr},
      {(void**)&P->wtmp, sizeof(f64) * 4 * (size_t)N, nullptr},
      {(void**)&P->ztmp, sizeof(f64) * 4 * (size_t)N, nullptr},
      {(void**)&P->lam, sizeof(unsigned long long), nullptr}};
  for (auto& u : up) {
    if (cudaMalloc(u.p, u.bytes ? u.bytes : 1) != cudaSuccess) {
      set_error("dfb_pc2_create: cudaMalloc of %zu bytes failed", u.bytes);
      dfb_pc2_destroy(P);
      return DFB_ERR_CUDA;
    }
    if (u.src) cudaMemcpy(*u.p, u.src, u.bytes, cudaMemcpyHostToDevice);
  }
  *out = P;
  return DFB_OK;
}

int dfb_pc2_info(const dfb_pc2* P, int* num_aggregates, int* coarse_nnz) {
  if (!P) { set_error("dfb_pc2_info: bad argument"); return DFB_ERR_ARG; }
  if (num_aggregates) *num_aggregates = P->Nc;
  if (coarse_nnz) *coarse_nnz = P->cnnz;
  return DFB_OK;
}

// Numeric part, once per solve (the Jacobian changes every Newton iteration): D^-1, diag(S), the Galerkin coarse matrix.
int dfb_pc2_setup(dfb_pc2* P, const double* A00, const double* A01, const double* A10, const double* A11, void* stream) {
  cudaStream_t st = as_stream(stream);
  if (!P || !A00 || !A01 || !A10 || !A11) { set_error("dfb_pc2_setup: bad argument"); return DFB_ERR_ARG; }
  const int N = P->N, Nc = P->Nc;
  DFB_CUDA(cudaMemsetAsync(P->bad, 0, sizeof(int), st));
  DFB_CUDA(cudaMemsetAsync(P->Sc, 0, sizeof(f64) * (size_t)P->cnnz, st));
  DFB_CUDA(cudaMemsetAsync(P->lam, 0, sizeof(unsigned long long), st));
  k_pc2_dinv<<<ceil_div(N, 128), 128, 0, st>>>(N, P->rp, P->ci, A00, P->Dinv, P->bad);
  DFB_LAUNCH_CHECK();
  k_pc2_diagS<<<ceil_div(N, 128), 128, 0, st>>>(N, P->rp, P->ci, A01, A10, A11, P->Dinv, P->dSinv, P->bad);
  DFB_LAUNCH_CHECK();
  k_pc2_coarse_a11<<<ceil_div(N, 128), 128, 0, st>>>(N, P->rp, P->ci, A11, P->agg, P->crp, P->cci, P->Sc, P->bad);
  DFB_LAUNCH_CHECK();
  k_pc2_coarse_schur<<<ceil_div(N, 128), 128, 0, st>>>(N, P->rp, P->ci, A01, A10, P->Dinv, P->agg, P->crp, P->cci, P->Sc, P->bad);
  DFB_LAUNCH_CHECK();
  k_pc2_coarse_diag<<<ceil_div(Nc, 128), 128, 0, st>>>(Nc, P->crp, P->cci, P->Sc, P->dScinv, P->lam, P->bad);
  DFB_LAUNCH_CHECK();
  int bad = 0;
  DFB_CUDA(cudaMemcpyAsync(&bad, P->bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  DFB_CUDA(cudaStreamSynchronize(st));
  if (bad) {
    static const char* why[] = {"", "singular or missing diagonal block", "coarse pattern miss", "more than 32 aggregates around a node",
                                "singular coarse diagonal"};
    set_error("dfb_pc2_setup: %s", why[bad < 5 ? bad : 0]);
    P->ready = false;
    return DFB_ERR_ARG;
  }
  P->ready = true;
  return DFB_OK;
}

// y = P2^-1 x on ABI-layout vectors (rows [4N, 6N) copied through, like the reference's PCNone sections)
int dfb_pc2_apply(dfb_pc2* P, const double* A10, const double* d_x, double* d_y, void* stream) {
  cudaStream_t st = as_stream(stream);
  if (!P || !P->ready || !A10 || !d_x || !d_y) { set_error("dfb_pc2_apply: bad argument (run dfb_pc2_setup first)"); return DFB_ERR_ARG; }
  const int N = P->N;
  k_pc2_pack<<<ceil_div((i64)4 * N, 256), 256, 0, st>>>(N, d_x, P->wtmp);
  DFB_LAUNCH_CHECK();
  DFB_CHECK(pc2_apply_aos(P, A10, P->wtmp, P->ztmp, st));
  k_pc2_unpack<<<ceil_div((i64)4 * N, 256), 256, 0, st>>>(N, P->ztmp, d_y);
  DFB_LAUNCH_CHECK();
  DFB_CUDA(cudaMemcpyAsync(d_y + (size_t)4 * N, d_x + (size_t)4 * N, sizeof(f64) * 2 * (size_t)N, cudaMemcpyDeviceToDevice, st));
  return DFB_OK;
}

}  // extern "C"
