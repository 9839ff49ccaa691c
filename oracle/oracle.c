/*
 * oracle.c -- CPU restatement of DEDFlow's FEM linear-system hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker / the reported CPU baseline.
 *
 * Parity status: the reference ships no golden vectors or tests.  This
 * restatement is pinned against the reference's own CUDA build (oracle/_ref, built from
 * the sources under /root/reference by oracle/ref/Makefile) run on a B200; the outputs of
 * that run are committed as tests/golden/ref_*.npz together with the generating script
 * (oracle/ref/run_ref.py).  See DESIGN.md "Oracle".
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/src).  Plain C99 + OpenMP; compiled with -ffp-contract=off so every
 * floating point operation is the IEEE operation written here.
 *
 * Layouts (identical to the reference):
 *   xg[3N] interleaved, ien[4E], every state vector is 6N: [u: N x 3 | p: N | phi: N | T: N]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int32_t i32;
typedef int64_t i64;
typedef uint32_t u32;
typedef double f64;

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------- */
/* constants: assemble.cu:23-118, main.c:23-27                                             */
/* ------------------------------------------------------------------------------------- */
#define kRHOC (0.5)
#define kDT (5e-2)
#define kALPHAM ((3.0 - kRHOC) / (1.0 + kRHOC))
#define kALPHAF (1.0 / (1.0 + kRHOC))
#define kGAMMA (0.5 + kALPHAM - kALPHAF)
#define kRHO (1.0e3)
#define kCP (1.0)
#define kKAPPA (0.66)
#define kMU (10.0 / 3.0)
#define NSHL 4
#define NQR 4
#define NQRB 3
#define BS 6

static const f64 c_fb[3] = {0.0, 0.0, -9.81 * 0.0};
static const f64 c_gw[4] = {0.0416666666666667, 0.0416666666666667, 0.0416666666666667, 0.0416666666666667};
#define SA 0.5854101966249685
#define SB 0.1381966011250105
static const f64 c_shlu[16] = {SA, SB, SB, SB, SB, SA, SB, SB, SB, SB, SA, SB, SB, SB, SB, SA};
static const f64 c_gwb[3] = {0.1666666666666667, 0.1666666666666667, 0.1666666666666667};
#define T6 0.1666666666666667
#define T3 0.6666666666666667
/* c_shlub[forn*12 + iq*4 + a]  (assemble.cu:87-102) */
static const f64 c_shlub[48] = {
    0.0, T6, T6, T3, 0.0, T6, T3, T6, 0.0, T3, T6, T6,
    T6, 0.0, T6, T3, T6, 0.0, T3, T6, T3, 0.0, T6, T6,
    T3, T6, 0.0, T6, T6, T3, 0.0, T6, T6, T6, 0.0, T3,
    T6, T3, T6, 0.0, T6, T6, T3, 0.0, T3, T6, T6, 0.0};
/* reference-face normals, assemble.cu:114-118 */
static const f64 c_nv2[12] = {1.0, 1.0, 1.0, -1.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, -1.0};

ORC_API int orc_version(void) { return 1; }
ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ===================================================================================== */
/* a1. scalar nodal sparsity pattern -- csr.c:57-79 (sorted insert), :81-133, :143-190    */
/* ===================================================================================== */
#define PREALLOC_SIZE 64

static i32 *lower_bound_i32(i32 *first, i32 *last, i32 value) {
  i64 count = last - first;
  while (count > 0) {
    i64 step = count / 2;
    i32 *it = first + step;
    if (*it < value) {
      first = it + 1;
      count -= step + 1;
    } else {
      count = step;
    }
  }
  return first;
}

/* returns nnz, or -1 if a row would exceed PREALLOC_SIZE (the reference ASSERTs, csr.c:64) */
ORC_API i64 orc_nodal_pattern(i32 num_node, i32 num_tet, const i32 *ien, i32 *row_ptr, i32 *col_ind_cap /* N*64 */) {
  i32 *buff = (i32 *)calloc((size_t)num_node * PREALLOC_SIZE, sizeof(i32));
  i32 *row_len = (i32 *)calloc((size_t)num_node, sizeof(i32));
  i64 nnz = 0;
  for (i32 k = 0; k < num_tet; k++) {
    const i32 *elem = ien + (size_t)k * 4;
    for (int i = 0; i < 4; i++) {
      for (int jj = -1; jj < 4; jj++) { /* push (i,i) first, then all j != i : csr.c:101-106 */
        int j = jj < 0 ? i : jj;
        if (jj >= 0 && j == i) continue;
        i32 key = elem[i], value = elem[j];
        i32 *row = buff + (size_t)key * PREALLOC_SIZE;
        i32 len = row_len[key];
        i32 *it = lower_bound_i32(row, row + len, value);
        if (it == row + len || *it != value) {
          if (len >= PREALLOC_SIZE) {
            free(buff);
            free(row_len);
            return -1;
          }
          memmove(it + 1, it, (size_t)(row + len - it) * sizeof(i32));
          *it = value;
          row_len[key]++;
        }
      }
    }
  }
  row_ptr[0] = 0;
  for (i32 i = 0; i < num_node; i++) {
    row_ptr[i + 1] = row_ptr[i] + row_len[i];
    memcpy(col_ind_cap + row_ptr[i], buff + (size_t)i * PREALLOC_SIZE, sizeof(i32) * (size_t)row_len[i]);
    nnz += row_len[i];
  }
  free(buff);
  free(row_len);
  return nnz;
}

/* a2. blocked scalar-CSR expansion -- csr_impl.cu:24-59.
 * The reference never writes new_row_ptr[num_row*block_row] (defect D1); fix_last!=0 writes nnz there,
 * fix_last==0 leaves the zero of the zero-filling allocator (alloc.c:23-30). */
ORC_API void orc_expand_block(i32 num_row, const i32 *row_ptr, const i32 *col_ind, i32 br, i32 bc, i32 *new_row_ptr,
                              i32 *new_col_ind, int fix_last) {
  for (i32 i = 0; i < num_row; i++) {
    i32 start = row_ptr[i], len = row_ptr[i + 1] - start;
    for (i32 j = 0; j < br; j++) new_row_ptr[i * br + j] = start * br * bc + j * bc * len;
  }
  new_row_ptr[(size_t)num_row * br] = fix_last ? row_ptr[num_row] * br * bc : 0;
  for (i32 i = 0; i < num_row; i++) {
    i32 start = row_ptr[i], len = row_ptr[i + 1] - start;
    for (i32 j = 0; j < br; j++) {
      i32 row = i * br + j;
      for (i32 k = 0; k < len; k++) {
        i32 col = col_ind[start + k];
        for (i32 l = 0; l < bc; l++) new_col_ind[new_row_ptr[row] + k * bc + l] = col * bc + l;
      }
    }
  }
}

/* ===================================================================================== */
/* a3. coloring -- color_impl.cu:17-61 (v2e map), :185-192,225-237 (weights), :63-183 (JPL)*/
/* ===================================================================================== */
/* weights from raw XORWOW u32 draws: val % (INT_MAX/2 - 0) + 0  (color_impl.cu:9-13,185-192) */
ORC_API void orc_weights_from_u32(i64 n, const u32 *raw, i32 *weight) {
  const u32 ub = (u32)(2147483647 / 2);
  for (i64 i = 0; i < n; i++) weight[i] = (i32)(raw[i] % ub);
}

/* vertex->element CSR; columns in ascending element order (the reference's order is atomic-arrival,
 * color_impl.cu:38-48, and does not influence any result). */
ORC_API void orc_v2e(i32 num_node, i32 num_elem, const i32 *ien, i32 *row_ptr, i32 *col_ind) {
  memset(row_ptr, 0, sizeof(i32) * ((size_t)num_node + 1));
  for (i64 i = 0; i < (i64)num_elem * 4; i++) row_ptr[ien[i] + 1]++;
  for (i32 i = 0; i < num_node; i++) row_ptr[i + 1] += row_ptr[i];
  i32 *cnt = (i32 *)calloc((size_t)num_node, sizeof(i32));
  for (i32 e = 0; e < num_elem; e++)
    for (int j = 0; j < 4; j++) {
      i32 node = ien[(size_t)e * 4 + j];
      col_ind[row_ptr[node] + cnt[node]++] = e;
    }
  free(cnt);
}

/* Jones-Plassmann-Luby element coloring.  val[] holds the weight (>=0) of an uncolored element and
 * -1-c once colored (color_impl.cu:87,120-126,165-171).  Round c marks every uncolored element i with no
 * vertex-sharing elem != i such that val[i] < val[elem] (strict, :87).  The reference marks in place
 * (benign race); on tie-free inputs that equals the snapshot semantics used here.  *n_tie_pairs counts
 * ordered (i,elem) incidences with equal weights among vertex-sharing elements (defect D2): bit-exact
 * colors are only defined when it is 0.  Returns the number of rounds executed; color[] in [0,rounds). */
ORC_API i32 orc_color_jpl(i32 num_elem, i32 num_node, const i32 *ien, const i32 *weight, i32 max_color, i32 *color,
                          i64 *n_tie_pairs) {
  i32 *v2e_ptr = (i32 *)malloc(sizeof(i32) * ((size_t)num_node + 1));
  i32 *v2e_col = (i32 *)malloc(sizeof(i32) * (size_t)num_elem * 4);
  orc_v2e(num_node, num_elem, ien, v2e_ptr, v2e_col);
  i32 *val = (i32 *)malloc(sizeof(i32) * (size_t)num_elem);
  unsigned char *mark = (unsigned char *)malloc((size_t)num_elem);
  memcpy(val, weight, sizeof(i32) * (size_t)num_elem);
  i64 ties = 0;
#pragma omp parallel for reduction(+ : ties) schedule(static)
  for (i32 i = 0; i < num_elem; i++)
    for (int j = 0; j < 4; j++) {
      i32 node = ien[(size_t)i * 4 + j];
      for (i32 k = v2e_ptr[node]; k < v2e_ptr[node + 1]; k++) {
        i32 e = v2e_col[k];
        if (e != i && weight[e] == weight[i]) ties++;
      }
    }
  if (n_tie_pairs) *n_tie_pairs = ties;
  i32 c = 0;
  i64 left = num_elem;
  for (; c < max_color && left; c++) {
#pragma omp parallel for schedule(dynamic, 4096)
    for (i32 i = 0; i < num_elem; i++) {
      mark[i] = 0;
      i32 ec = val[i];
      if (ec < 0) continue;
      int found_max = 1;
      for (int j = 0; j < 4; j++) {
        i32 node = ien[(size_t)i * 4 + j];
        for (i32 k = v2e_ptr[node]; k < v2e_ptr[node + 1]; k++) {
          i32 e = v2e_col[k];
          if (ec < val[e] && e != i) found_max = 0;
        }
      }
      mark[i] = (unsigned char)found_max;
    }
    left = 0;
#pragma omp parallel for reduction(+ : left) schedule(static)
    for (i32 i = 0; i < num_elem; i++) {
      if (mark[i]) val[i] = -1 - c;
      left += val[i] >= 0;
    }
  }
  for (i32 i = 0; i < num_elem; i++) color[i] = val[i] * (-1) - 1; /* RecoverColorKernel :136-141 */
  free(v2e_ptr);
  free(v2e_col);
  free(val);
  free(mark);
  return c;
}

/* a4. color batches -- Mesh.c:165-206, indexing.cu:92-103: ascending element ids per color */
ORC_API void orc_color_batches(i32 num_elem, const i32 *color, i32 num_color, i32 *batch_offset, i32 *batch_ind) {
  memset(batch_offset, 0, sizeof(i32) * ((size_t)num_color + 1));
  for (i32 i = 0; i < num_elem; i++) batch_offset[color[i] + 1]++;
  for (i32 c = 0; c < num_color; c++) batch_offset[c + 1] += batch_offset[c];
  i32 *cur = (i32 *)malloc(sizeof(i32) * (size_t)num_color);
  memcpy(cur, batch_offset, sizeof(i32) * (size_t)num_color);
  for (i32 i = 0; i < num_elem; i++) batch_ind[cur[color[i]]++] = i;
  free(cur);
}

/* ===================================================================================== */
/* a6. element geometry                                                                    */
/* ===================================================================================== */
/* J = [x1-x0 | x2-x0 | x3-x0] column-major (assemble.cu:321-348); inverse by partially pivoted LU as the
 * batched getrf/getri pair does (assemble.cu:1275-1279); detJ = |U00*U11*U22| (assemble.cu:350-357).
 * inv[i + 3j] = Jinv(i,j). */
static void geom_invJ(const f64 *x0, const f64 *x1, const f64 *x2, const f64 *x3, f64 inv[9], f64 *detJ) {
  f64 A[9];
  int piv[3];
  for (int i = 0; i < 3; i++) {
    A[0 + i] = x1[i] - x0[i];
    A[3 + i] = x2[i] - x0[i];
    A[6 + i] = x3[i] - x0[i];
  }
  for (int k = 0; k < 3; k++) {
    int p = k;
    f64 amax = fabs(A[k + 3 * k]);
    for (int i = k + 1; i < 3; i++)
      if (fabs(A[i + 3 * k]) > amax) {
        amax = fabs(A[i + 3 * k]);
        p = i;
      }
    piv[k] = p;
    if (p != k)
      for (int j = 0; j < 3; j++) {
        f64 t = A[k + 3 * j];
        A[k + 3 * j] = A[p + 3 * j];
        A[p + 3 * j] = t;
      }
    for (int i = k + 1; i < 3; i++) {
      A[i + 3 * k] /= A[k + 3 * k];
      for (int j = k + 1; j < 3; j++) A[i + 3 * j] -= A[i + 3 * k] * A[k + 3 * j];
    }
  }
  *detJ = fabs(A[0] * A[4] * A[8]);
  /* solve A X = I column by column with the row permutation applied */
  for (int c = 0; c < 3; c++) {
    f64 b[3] = {0.0, 0.0, 0.0};
    b[c] = 1.0;
    for (int k = 0; k < 3; k++)
      if (piv[k] != k) {
        f64 t = b[k];
        b[k] = b[piv[k]];
        b[piv[k]] = t;
      }
    for (int i = 1; i < 3; i++)
      for (int j = 0; j < i; j++) b[i] -= A[i + 3 * j] * b[j];
    for (int i = 2; i >= 0; i--) {
      for (int j = i + 1; j < 3; j++) b[i] -= A[i + 3 * j] * b[j];
      b[i] /= A[i + 3 * i];
    }
    inv[0 + 3 * c] = b[0];
    inv[1 + 3 * c] = b[1];
    inv[2 + 3 * c] = b[2];
  }
}

/* shgrad[a*3 + d] = dN_a/dx_d : assemble.cu:1308-1328 */
static void shape_grad(const f64 inv[9], f64 sh[12]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) sh[i * 3 + j + 3] = inv[i + j * 3];
  sh[0] = -sh[3] - sh[6] - sh[9];
  sh[1] = -sh[4] - sh[7] - sh[10];
  sh[2] = -sh[5] - sh[8] - sh[11];
}

/* G = S^T S over shgrad+3 (assemble.cu:1586-1593): G[i + 3j] = sum_d sh[(i+1)*3+d] * sh[(j+1)*3+d] */
static void metric_G(const f64 sh[12], f64 G[9]) {
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++) {
      f64 s = 0.0;
      for (int d = 0; d < 3; d++) s += sh[(i + 1) * 3 + d] * sh[(j + 1) * 3 + d];
      G[i + 3 * j] = s;
    }
}

/* gather + interpolation: assemble.cu:135-154 (LoadElementValueKernel), :1601-1693 */
static void gather_interp(const i32 *nodes, i32 num_node, const f64 *wg, const f64 *dwg, const f64 sh[12], f64 qr_wg[24],
                          f64 qr_dwg[24], f64 grad[18]) {
  f64 buf[24];
  for (int a = 0; a < 4; a++) {
    i32 n = nodes[a];
    buf[0 * 4 + a] = wg[(size_t)n * 3 + 0];
    buf[1 * 4 + a] = wg[(size_t)n * 3 + 1];
    buf[2 * 4 + a] = wg[(size_t)n * 3 + 2];
    buf[3 * 4 + a] = dwg[(size_t)num_node * 3 + n]; /* pressure from the increment vector: defect D6 */
    buf[4 * 4 + a] = wg[(size_t)num_node * 4 + n];
    buf[5 * 4 + a] = wg[(size_t)num_node * 5 + n];
  }
  for (int comp = 0; comp < 6; comp++) {
    for (int d = 0; d < 3; d++) {
      f64 s = 0.0;
      for (int a = 0; a < 4; a++) s += sh[a * 3 + d] * buf[comp * 4 + a];
      grad[comp * 3 + d] = s;
    }
    for (int q = 0; q < 4; q++) {
      f64 s = 0.0;
      for (int a = 0; a < 4; a++) s += c_shlu[q + 4 * a] * buf[comp * 4 + a];
      qr_wg[comp * 4 + q] = s;
    }
  }
  if (qr_dwg) {
    for (int a = 0; a < 4; a++) {
      i32 n = nodes[a];
      buf[0 * 4 + a] = dwg[(size_t)n * 3 + 0];
      buf[1 * 4 + a] = dwg[(size_t)n * 3 + 1];
      buf[2 * 4 + a] = dwg[(size_t)n * 3 + 2];
      buf[3 * 4 + a] = dwg[(size_t)num_node * 3 + n];
      buf[4 * 4 + a] = dwg[(size_t)num_node * 4 + n];
      buf[5 * 4 + a] = dwg[(size_t)num_node * 5 + n];
    }
    for (int comp = 0; comp < 6; comp++)
      for (int q = 0; q < 4; q++) {
        f64 s = 0.0;
        for (int a = 0; a < 4; a++) s += c_shlu[a + 4 * q] * buf[comp * 4 + a];
        qr_dwg[comp * 4 + q] = s;
      }
  }
}

/* GetStabTau: assemble.cu:444-484 */
static void stab_tau(const f64 *G, const f64 *uadv, f64 rho, f64 cp, f64 mu, f64 kappa, f64 dt, f64 *tau) {
  f64 t0 = 4.0 / (dt * dt), t1 = 0.0, t2 = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      t1 += G[i * 3 + j] * uadv[i] * uadv[j];
      t2 += G[i * 3 + j] * G[i * 3 + j];
    }
  mu /= rho;
  kappa /= rho * cp;
  tau[0] = (1.0 / sqrt(t0 + t1 + 3.0 * mu * mu * t2)) / rho;
  tau[1] = sqrt(t1 + 3.0 * mu * mu * t2) / (G[0] + G[4] + G[8]);
  tau[2] = 1.0 / sqrt(t0 + t1);
  tau[3] = (1.0 / sqrt(t0 + t1 + 3.0 * kappa * kappa * t2)) / (rho * cp);
}

/* element residual: AssembleWeakFormKernel<.,.,4,1>, assemble.cu:761-924.  elem_F[a*6 + ii] */
static void elem_residual(const f64 *G, f64 detJ, const f64 *sh, const f64 *qr_wg, const f64 *qr_dwg, const f64 *grad,
                          f64 *elem_F) {
  for (int i = 0; i < 24; i++) elem_F[i] = 0.0;
  f64 tau[4] = {0, 0, 0, 0};
  f64 divu = grad[0] + grad[4] + grad[8];
  f64 rLi[3], uadv[3], shconv[4];
  for (int iq = 0; iq < NQR; iq++) {
    uadv[0] = qr_wg[NQR * 0 + iq];
    uadv[1] = qr_wg[NQR * 1 + iq];
    uadv[2] = qr_wg[NQR * 2 + iq];
    for (int i = 0; i < 3; i++) {
      rLi[i] = 0.0;
      rLi[i] += kRHO * (qr_dwg[NQR * i + iq] - c_fb[i]);
      rLi[i] += kRHO * uadv[0] * grad[3 * i + 0];
      rLi[i] += kRHO * uadv[1] * grad[3 * i + 1];
      rLi[i] += kRHO * uadv[2] * grad[3 * i + 2];
      rLi[i] += grad[3 * 3 + i];
    }
    stab_tau(G, uadv, kRHO, kCP, kMU, kKAPPA, kDT, tau);
    for (int aa = 0; aa < 4; aa++) {
      shconv[aa] = 0.0;
      shconv[aa] += uadv[0] * sh[aa * 3 + 0];
      shconv[aa] += uadv[1] * sh[aa * 3 + 1];
      shconv[aa] += uadv[2] * sh[aa * 3 + 2];
    }
    f64 tmp0[3], tmp1[9];
    for (int i = 0; i < 3; i++) {
      tmp0[i] = 0.0;
      tmp0[i] += kRHO * (qr_dwg[NQR * i + iq] - c_fb[i]);
      tmp0[i] += kRHO * (uadv[0] - tau[0] * rLi[0]) * grad[3 * i + 0];
      tmp0[i] += kRHO * (uadv[1] - tau[0] * rLi[1]) * grad[3 * i + 1];
      tmp0[i] += kRHO * (uadv[2] - tau[0] * rLi[2]) * grad[3 * i + 2];
    }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        tmp1[i * 3 + j] = 0.0;
        tmp1[i * 3 + j] += kMU * (grad[3 * i + j] + grad[3 * j + i]);
        tmp1[i * 3 + j] += kRHO * tau[0] * rLi[i] * uadv[j];
        tmp1[i * 3 + j] -= kRHO * tau[0] * tau[0] * rLi[i] * rLi[j];
      }
    for (int i = 0; i < 3; i++) tmp1[i * 3 + i] += -qr_wg[NQR * 3 + iq] + kRHO * tau[1] * divu;
    for (int aa = 0; aa < 4; aa++)
      for (int ii = 0; ii < 3; ii++) {
        f64 bm = 0.0;
        bm += c_shlu[aa * NQR + iq] * tmp0[ii];
        bm += sh[aa * 3 + 0] * tmp1[ii * 3 + 0];
        bm += sh[aa * 3 + 1] * tmp1[ii * 3 + 1];
        bm += sh[aa * 3 + 2] * tmp1[ii * 3 + 2];
        elem_F[aa * BS + ii] += bm * c_gw[iq] * detJ;
      }
    for (int aa = 0; aa < 4; aa++) {
      f64 bc = 0.0;
      bc += c_shlu[aa * NQR + iq] * divu;
      bc += tau[0] * rLi[0] * sh[aa * 3 + 0];
      bc += tau[0] * rLi[1] * sh[aa * 3 + 1];
      bc += tau[0] * rLi[2] * sh[aa * 3 + 2];
      elem_F[aa * BS + 3] += bc * c_gw[iq] * detJ;
    }
    for (int aa = 0; aa < 4; aa++) {
      f64 bp = qr_dwg[NQR * 4 + iq] + uadv[0] * grad[3 * 4 + 0] + uadv[1] * grad[3 * 4 + 1] + uadv[2] * grad[3 * 4 + 2];
      elem_F[aa * BS + 4] += bp * (c_shlu[aa * NQR + iq] + tau[2] * shconv[aa]) * c_gw[iq] * detJ;
    }
    for (int aa = 0; aa < 4; aa++) {
      f64 bt = kRHO * kCP *
               (qr_dwg[NQR * 5 + iq] + uadv[0] * grad[3 * 5 + 0] + uadv[1] * grad[3 * 5 + 1] + uadv[2] * grad[3 * 5 + 2]) *
               (c_shlu[aa * NQR + iq] + kRHO * kCP * tau[3] * shconv[aa]);
      bt += kKAPPA * (grad[3 * 5 + 0] * sh[aa * 3 + 0] + grad[3 * 5 + 1] * sh[aa * 3 + 1] + grad[3 * 5 + 2] * sh[aa * 3 + 2]);
      elem_F[aa * BS + 5] += bt * c_gw[iq] * detJ;
    }
  }
}

/* element Jacobian: AssembleWeakFormLHSKernel, assemble.cu:495-759 (the live kernel; u,p 4x4 only).
 * elem_J[(a*4+b)*36 + ii*6 + jj].  Note the tau of this kernel uses sum_{a=1..3} (u.gradN_a)^2
 * (assemble.cu:592-602), not u.G.u (defect D5).  Entries (4,4),(5,5) receive (a==b) (assemble.cu:757-758). */
static void elem_jacobian(const f64 *G, f64 detJ, const f64 *sh, const f64 *qr_wg, f64 *elem_J) {
  const f64 fact1 = kALPHAM, fact2 = kDT * kALPHAF * kGAMMA;
  const f64 knu = kMU / kRHO;
  f64 gg = 0.0, tr = 0.0;
  for (int i = 0; i < 9; i++) {
    f64 gij = G[i];
    gg += gij * gij;
    tr += gij * (f64)(!(i & 0x3));
  }
  f64 itr = 1.0 / tr;
  f64 shconv[4][4], tauM[4], tauC[4]; /* [iq][a] */
  for (int iq = 0; iq < NQR; iq++) {
    for (int a = 0; a < 4; a++) {
      f64 s = 0.0;
      s += sh[a * 3 + 0] * qr_wg[0 * NQR + iq];
      s += sh[a * 3 + 1] * qr_wg[1 * NQR + iq];
      s += sh[a * 3 + 2] * qr_wg[2 * NQR + iq];
      shconv[iq][a] = s;
    }
    f64 tmp = 0;
    tmp += shconv[iq][1] * shconv[iq][1];
    tmp += shconv[iq][2] * shconv[iq][2];
    tmp += shconv[iq][3] * shconv[iq][3];
    tauM[iq] = (1.0 / sqrt(4.0 / (kDT * kDT) + tmp + 3.0 * knu * knu * gg)) / kRHO;
    tauC[iq] = sqrt(tmp + 3.0 * knu * knu * gg) * itr;
  }
  memset(elem_J, 0, sizeof(f64) * 576);
  for (int aa = 0; aa < 4; aa++)
    for (int bb = 0; bb < 4; bb++) {
      f64 b[16];
      for (int i = 0; i < 16; i++) b[i] = 0.0;
      for (int iq = 0; iq < NQR; iq++) {
        const f64 *sc = shconv[iq];
        f64 tau0 = tauM[iq], tau1 = tauC[iq];
        f64 eK = sh[aa * 3 + 0] * sh[bb * 3 + 0] + sh[aa * 3 + 1] * sh[bb * 3 + 1] + sh[aa * 3 + 2] * sh[bb * 3 + 2];
        f64 detJgw = detJ * c_gw[iq];
        f64 tmp = 0.0;
        tmp += fact1 * kRHO * c_shlu[aa * NQR + iq] * c_shlu[bb * NQR + iq];
        tmp += fact1 * kRHO * kRHO * tau0 * sc[aa] * c_shlu[bb * NQR + iq];
        tmp += fact2 * c_shlu[aa * NQR + iq] * kRHO * sc[bb];
        tmp += fact2 * tau0 * kRHO * sc[aa] * kRHO * sc[bb];
        tmp += fact2 * kMU * eK;
        b[0 * 4 + 0] += tmp * detJgw;
        b[1 * 4 + 1] += tmp * detJgw;
        b[2 * 4 + 2] += tmp * detJgw;
        for (int ii = 0; ii < 3; ii++)
          for (int jj = 0; jj < 3; jj++) {
            b[ii * 4 + jj] += fact2 * kMU * sh[aa * 3 + jj] * sh[bb * 3 + ii] * detJgw;
            b[ii * 4 + jj] += fact2 * kRHO * tau1 * sh[aa * 3 + ii] * sh[bb * 3 + jj] * detJgw;
          }
        for (int ii = 0; ii < 3; ii++) {
          b[ii * 4 + 3] -= sh[aa * 3 + ii] * c_shlu[bb * NQR + iq] * detJgw;
          b[ii * 4 + 3] += kRHO * tau0 * sc[aa] * sh[bb * 3 + ii] * detJgw;
        }
        for (int ii = 0; ii < 3; ii++) {
          b[3 * 4 + ii] += fact1 * kRHO * tau0 * sh[aa * 3 + ii] * c_shlu[bb * NQR + iq] * detJgw;
          b[3 * 4 + ii] += fact2 * c_shlu[aa * NQR + iq] * sh[bb * 3 + ii] * detJgw;
          b[3 * 4 + ii] += fact2 * tau0 * sh[aa * 3 + ii] * kRHO * sc[bb] * detJgw;
        }
        b[3 * 4 + 3] += tau0 * eK * detJgw;
      }
      f64 *dst = elem_J + (aa * 4 + bb) * 36;
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) dst[i * 6 + j] += b[i * 4 + j];
      dst[4 * 6 + 4] += (f64)(aa == bb);
      dst[5 * 6 + 5] += (f64)(aa == bb);
    }
}

/* one element: geometry -> gather/interpolate -> weak forms (the per-batch pipeline of assemble.cu:1559-1705) */
static void tet_element_one(i32 e, i32 num_node, const i32 *ien, const f64 *xg, const f64 *wg, const f64 *dwg, f64 *eF,
                            f64 *eJ, f64 *metric, f64 *shgrad) {
  const i32 *nodes = ien + (size_t)e * 4;
  f64 inv[9], detJ, sh[12], G[9], qr_wg[24], qr_dwg[24], grad[18];
  geom_invJ(xg + (size_t)nodes[0] * 3, xg + (size_t)nodes[1] * 3, xg + (size_t)nodes[2] * 3, xg + (size_t)nodes[3] * 3, inv,
            &detJ);
  shape_grad(inv, sh);
  metric_G(sh, G);
  gather_interp(nodes, num_node, wg, dwg, sh, qr_wg, qr_dwg, grad);
  if (eF) elem_residual(G, detJ, sh, qr_wg, qr_dwg, grad, eF);
  if (eJ) elem_jacobian(G, detJ, sh, qr_wg, eJ);
  if (metric) {
    memcpy(metric, G, sizeof(f64) * 9);
    metric[9] = detJ;
  }
  if (shgrad) memcpy(shgrad, sh, sizeof(f64) * 12);
}

/* Per-element checkpoints (the arrays the reference author dumps): elem_F[n*24], elem_J[n*576],
 * plus optional geometry (G col-major + detJ, shape gradients).  elem_ids == NULL means elements 0..n-1. */
ORC_API void orc_tet_elements(i32 n, const i32 *elem_ids, i32 num_node, const i32 *ien, const f64 *xg, const f64 *wg,
                              const f64 *dwg, f64 *elem_F, f64 *elem_J, f64 *elem_metric /* n*10: G,detJ */,
                              f64 *elem_shgrad /* n*12 */) {
#pragma omp parallel for schedule(static)
  for (i32 t = 0; t < n; t++) {
    i32 e = elem_ids ? elem_ids[t] : t;
    tet_element_one(e, num_node, ien, xg, wg, dwg, elem_F ? elem_F + (size_t)t * 24 : NULL,
                    elem_J ? elem_J + (size_t)t * 576 : NULL, elem_metric ? elem_metric + (size_t)t * 10 : NULL,
                    elem_shgrad ? elem_shgrad + (size_t)t * 12 : NULL);
  }
}

/* ElemRHSLocal2GlobalKernel x4: assemble.cu:188-208,1709-1724 */
static void scatter_F_one(const i32 *nodes, i32 num_node, const f64 *eF, f64 *F) {
  for (int a = 0; a < 4; a++) {
    i32 n = nodes[a];
    F[(size_t)n * 3 + 0] += eF[a * 6 + 0];
    F[(size_t)n * 3 + 1] += eF[a * 6 + 1];
    F[(size_t)n * 3 + 2] += eF[a * 6 + 2];
    F[(size_t)num_node * 3 + n] += eF[a * 6 + 3];
    F[(size_t)num_node * 4 + n] += eF[a * 6 + 4];
    F[(size_t)num_node * 5 + n] += eF[a * 6 + 5];
  }
}

/* SetBlockValueToSubmatKernel: matrix_impl.cu:370-453 with offset={0,3,4,5,6}, live blocks (0,0),(0,1),(1,0),(1,1)
 * (main.c:375-404); alpha=beta=1 (matrix.c:574-592). */
static void scatter_J_one(const i32 *nodes, const i32 *row_ptr, const i32 *col_ind, const f64 *eJ, f64 *A00, f64 *A01,
                          f64 *A10, f64 *A11) {
  for (int aa = 0; aa < 4; aa++)
    for (int bb = 0; bb < 4; bb++) {
      i32 row = nodes[aa], col = nodes[bb];
      i32 start = row_ptr[row], end = row_ptr[row + 1], len = end - start, k;
      for (k = start; k < end; k++)
        if (col_ind[k] == col) break;
      const f64 *val = eJ + (aa * 4 + bb) * 36;
      size_t s = (size_t)start, ko = (size_t)(k - start);
      for (int ii = 0; ii < 3; ii++)
        for (int jj = 0; jj < 3; jj++) A00[s * 9 + ko * 3 + (size_t)ii * len * 3 + jj] += val[ii * 6 + jj];
      for (int ii = 0; ii < 3; ii++) A01[s * 3 + ko + (size_t)ii * len] += val[ii * 6 + 3];
      for (int jj = 0; jj < 3; jj++) A10[s * 3 + ko * 3 + jj] += val[3 * 6 + jj];
      A11[s + ko] += val[3 * 6 + 3];
    }
}

/* AssembleSystemTet: assemble.cu:1467-1762.  Color batches in order, elements ascending inside a batch, so the
 * floating-point accumulation order is the reference's.  F and/or the four sub-block value arrays may be NULL. */
ORC_API void orc_assemble_tet(i32 num_node, const i32 *ien, const f64 *xg, i32 num_batch, const i32 *batch_offset,
                              const i32 *batch_ind, const f64 *wg, const f64 *dwg, f64 *F, const i32 *row_ptr,
                              const i32 *col_ind, f64 *A00, f64 *A01, f64 *A10, f64 *A11) {
  int doJ = A00 != NULL;
  for (i32 b = 0; b < num_batch; b++) {
    i32 bs = batch_offset[b + 1] - batch_offset[b];
    if (bs == 0) break; /* assemble.cu:1565-1567 */
    const i32 *ids = batch_ind + batch_offset[b];
#pragma omp parallel
    {
      f64 *eJ = doJ ? (f64 *)malloc(sizeof(f64) * 576) : NULL;
      f64 eF[24];
#pragma omp for schedule(static)
      for (i32 t = 0; t < bs; t++) {
        tet_element_one(ids[t], num_node, ien, xg, wg, dwg, F ? eF : NULL, eJ, NULL, NULL);
        const i32 *nodes = ien + (size_t)ids[t] * 4;
        if (F) scatter_F_one(nodes, num_node, eF, F);
        if (doJ) scatter_J_one(nodes, row_ptr, col_ind, eJ, A00, A01, A10, A11);
      }
      free(eJ);
    }
  }
}

/* ===================================================================================== */
/* a8. boundary faces -- assemble.cu:1764-1964, kernels :279-319, :1038-1214               */
/* ===================================================================================== */
static void face_element(i32 e, i32 iorn, i32 num_node, const i32 *ien, const f64 *xg, const f64 *wg, const f64 *dwg,
                         f64 *elem_F, f64 *elem_J) {
  const i32 *nodes = ien + (size_t)e * 4;
  f64 inv[9], detJ, sh[12], nv[3];
  geom_invJ(xg + (size_t)nodes[0] * 3, xg + (size_t)nodes[1] * 3, xg + (size_t)nodes[2] * 3, xg + (size_t)nodes[3] * 3, inv,
            &detJ);
  shape_grad(inv, sh);
  { /* GetElemFaceNVKernel :305-317 (Nanson) */
    f64 b[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++)
      for (int n = 0; n < 3; n++) b[n] += inv[n * 3 + k] * c_nv2[iorn * 3 + k];
    nv[0] = b[0] * detJ;
    nv[1] = b[1] * detJ;
    nv[2] = b[2] * detJ;
  }
  /* gather: only u (wgalpha) and p (dwgalpha+3N) are loaded (:1841-1848); comps 4,5 of the zero-filled buffer stay 0 */
  f64 buf[24];
  memset(buf, 0, sizeof(buf));
  for (int a = 0; a < 4; a++) {
    i32 n = nodes[a];
    buf[0 * 4 + a] = wg[(size_t)n * 3 + 0];
    buf[1 * 4 + a] = wg[(size_t)n * 3 + 1];
    buf[2 * 4 + a] = wg[(size_t)n * 3 + 2];
    buf[3 * 4 + a] = dwg[(size_t)num_node * 3 + n];
  }
  f64 grad[18], qr[18]; /* qr[comp*3 + q] */
  for (int comp = 0; comp < 6; comp++) {
    for (int d = 0; d < 3; d++) {
      f64 s = 0.0;
      for (int a = 0; a < 4; a++) s += sh[a * 3 + d] * buf[comp * 4 + a];
      grad[comp * 3 + d] = s;
    }
    for (int q = 0; q < NQRB; q++) {
      f64 s = 0.0;
      for (int a = 0; a < 4; a++) s += c_shlub[iorn * 12 + q * 4 + a] * buf[comp * 4 + a];
      qr[comp * 3 + q] = s;
    }
  }
  /* FaceAssemblyKernel :1054-1064 */
  f64 hinv = 0.0, detJb = 0.0, uadv[3];
  for (int i = 0; i < 3; i++) {
    uadv[i] = inv[i + 3 * 0] * nv[0] + inv[i + 3 * 1] * nv[1] + inv[i + 3 * 2] * nv[2];
    hinv += uadv[i] * uadv[i];
    detJb += nv[i] * nv[i];
  }
  detJb = sqrt(detJb);
  (void)detJb;
  hinv = sqrt(hinv);
  f64 tau_b = 4.0 * kMU * hinv;
  const f64 *shl = c_shlub + NQRB * NSHL * iorn;
  if (elem_F) {
    for (int i = 0; i < 24; i++) elem_F[i] = 0.0;
    f64 tmp0[3], tmp1[9];
    for (int iq = 0; iq < NQRB; iq++) {
      uadv[0] = qr[NQRB * 0 + iq];
      uadv[1] = qr[NQRB * 1 + iq];
      uadv[2] = qr[NQRB * 2 + iq];
      f64 unor = uadv[0] * nv[0] + uadv[1] * nv[1] + uadv[2] * nv[2];
      f64 uneg = (unor - fabs(unor)) * 0.5;
      for (int i = 0; i < 3; i++) {
        tmp0[i] = 0.0;
        tmp0[i] += nv[i] * qr[NQRB * 3 + iq];
        tmp0[i] -= kMU * (nv[0] * grad[3 * i + 0] + nv[1] * grad[3 * i + 1] + nv[2] * grad[3 * i + 2]);
        tmp0[i] -= kMU * (nv[0] * grad[3 * 0 + i] + nv[1] * grad[3 * 1 + i] + nv[2] * grad[3 * 2 + i]);
        tmp0[i] -= kRHO * uneg * uadv[i];
        tmp0[i] += tau_b * uadv[i];
      }
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) tmp1[i * 3 + j] = -kMU * (nv[i] * uadv[j] + nv[j] * uadv[i]);
      for (int aa = 0; aa < 4; aa++) {
        for (int ii = 0; ii < 3; ii++) {
          f64 bm = 0.0;
          bm += shl[iq * NSHL + aa] * tmp0[ii];
          bm += sh[aa * 3 + 0] * tmp1[ii * 3 + 0];
          bm += sh[aa * 3 + 1] * tmp1[ii * 3 + 1];
          bm += sh[aa * 3 + 2] * tmp1[ii * 3 + 2];
          elem_F[aa * BS + ii] += bm * c_gwb[iq];
        }
        elem_F[aa * BS + 3] -= shl[iq * NSHL + aa] * unor * c_gwb[iq];
      }
    }
  }
  if (elem_J) {
    const f64 fact2 = kDT * kALPHAF * kGAMMA;
    memset(elem_J, 0, sizeof(f64) * 576);
    f64 shnorm[4];
    for (int aa = 0; aa < 4; aa++) {
      shnorm[aa] = 0.0;
      shnorm[aa] += sh[aa * 3 + 0] * nv[0];
      shnorm[aa] += sh[aa * 3 + 1] * nv[1];
      shnorm[aa] += sh[aa * 3 + 2] * nv[2];
    }
#define M4(aa, bb, ii, jj) elem_J[((aa)*4 + (bb)) * 36 + (ii)*6 + (jj)]
    for (int iq = 0; iq < NQRB; iq++) {
      uadv[0] = qr[NQRB * 0 + iq];
      uadv[1] = qr[NQRB * 1 + iq];
      uadv[2] = qr[NQRB * 2 + iq];
      f64 unor = uadv[0] * nv[0] + uadv[1] * nv[1] + uadv[2] * nv[2];
      f64 uneg = (unor - fabs(unor)) * 0.5;
      for (int aa = 0; aa < 4; aa++)
        for (int bb = 0; bb < 4; bb++) {
          f64 Na = shl[iq * NSHL + aa], Nb = shl[iq * NSHL + bb];
          f64 t0 = 0.0;
          t0 -= kMU * (shnorm[bb] * Na + shnorm[aa] * Nb);
          t0 -= kRHO * Na * Nb * uneg;
          t0 += tau_b * Na * Nb;
          M4(aa, bb, 0, 0) += fact2 * t0 * c_gwb[iq];
          M4(aa, bb, 1, 1) += fact2 * t0 * c_gwb[iq];
          M4(aa, bb, 2, 2) += fact2 * t0 * c_gwb[iq];
          for (int ii = 0; ii < 3; ii++)
            for (int jj = 0; jj < 3; jj++) {
              t0 = 0.0;
              t0 -= kMU * Na * sh[bb * 3 + ii] * nv[jj];
              t0 -= kMU * Nb * sh[aa * 3 + jj] * nv[ii];
              M4(aa, bb, ii, jj) += fact2 * t0 * c_gwb[iq];
            }
          t0 = Na * Nb;
          for (int ii = 0; ii < 3; ii++) {
            M4(aa, bb, 3, ii) -= fact2 * t0 * nv[ii] * c_gwb[iq];
            M4(aa, bb, ii, 3) += t0 * nv[ii] * c_gwb[iq];
          }
        }
    }
#undef M4
  }
}

ORC_API void orc_face_elements(i32 num_face, const i32 *f2e, const i32 *forn, i32 num_node, const i32 *ien, const f64 *xg,
                               const f64 *wg, const f64 *dwg, f64 *elem_F, f64 *elem_J) {
#pragma omp parallel for schedule(static)
  for (i32 f = 0; f < num_face; f++)
    face_element(f2e[f], forn[f], num_node, ien, xg, wg, dwg, elem_F ? elem_F + (size_t)f * 24 : NULL,
                 elem_J ? elem_J + (size_t)f * 576 : NULL);
}

/* AssembleSystemTetFace for ONE boundary group (the driver runs id 4 only, assemble.cu:1825-1828, defect D13).
 * Scatter order: colors 0..num_color-1, faces ascending inside a color (the mask loop :1916-1945). */
ORC_API void orc_assemble_face(i32 num_face, const i32 *f2e, const i32 *forn, i32 num_node, const i32 *ien, const f64 *xg,
                               const i32 *color, i32 num_color, const f64 *wg, const f64 *dwg, f64 *F, const i32 *row_ptr,
                               const i32 *col_ind, f64 *A00, f64 *A01, f64 *A10, f64 *A11) {
  int doJ = A00 != NULL;
  f64 *eF = F ? (f64 *)malloc(sizeof(f64) * 24 * (size_t)num_face) : NULL;
  f64 *eJ = doJ ? (f64 *)malloc(sizeof(f64) * 576 * (size_t)num_face) : NULL;
  orc_face_elements(num_face, f2e, forn, num_node, ien, xg, wg, dwg, eF, eJ);
  for (i32 c = 0; c < num_color; c++)
    for (i32 f = 0; f < num_face; f++) {
      if (color[f2e[f]] != c) continue;
      const i32 *nodes = ien + (size_t)f2e[f] * 4;
      if (F) scatter_F_one(nodes, num_node, eF + (size_t)f * 24, F);
      if (doJ) scatter_J_one(nodes, row_ptr, col_ind, eJ + (size_t)f * 576, A00, A01, A10, A11);
    }
  free(eF);
  free(eJ);
}

/* ===================================================================================== */
/* a9. Dirichlet -- dirichlet.c:31-61, dirichlet_impl.cu:15-37, matrix_impl.cu:6-23, matrix.c:449-469 */
/* ===================================================================================== */
ORC_API void orc_dirichlet_vec(i32 nb, const i32 *bnode, i32 shape, const i32 *bctype, f64 *b) {
  for (i32 ic = 0; ic < shape; ic++)
    if (bctype[ic] == 1)
      for (i32 i = 0; i < nb; i++) b[(size_t)bnode[i] * shape + ic] = 0.0;
}

/* rows node*shape+ic of A00 -> unit rows, of A01 -> zero rows.  Sections >= 1 are skipped because the reference
 * passes a negative count (defect D8). */
ORC_API void orc_dirichlet_mat(i32 nb, const i32 *bnode, i32 shape, const i32 *bctype, i32 num_node, const i32 *row_ptr,
                               const i32 *col_ind, f64 *A00, f64 *A01) {
  for (i32 ic = 0; ic < shape; ic++) {
    if (bctype[ic] != 1) continue;
    for (i32 i = 0; i < nb; i++) {
      i32 ir = bnode[i] * shape + ic;
      if (ir < 0 || ir >= 3 * num_node) continue;
      i32 node = ir / 3, ii = ir % 3;
      i32 start = row_ptr[node], len = row_ptr[node + 1] - start;
      /* scalar row ir of the 3x3 pattern: entries start*9 + ii*3*len + [0, 3*len), column = col*3 + l */
      for (i32 k = 0; k < len; k++)
        for (int l = 0; l < 3; l++)
          A00[(size_t)start * 9 + (size_t)ii * 3 * len + (size_t)k * 3 + l] = 1.0 * (f64)(col_ind[start + k] * 3 + l == ir);
      for (i32 k = 0; k < len; k++) A01[(size_t)start * 3 + (size_t)ii * len + k] = 0.0;
    }
  }
}

/* ===================================================================================== */
/* a10. field-split mat-vec -- matrix.c:471-524.  y[0:4N) = beta*y + alpha*sum_j A_ij x_j; y[4N:6N) untouched (D4) */
/* ===================================================================================== */
ORC_API void orc_fs_amvpby(i32 N, const i32 *row_ptr, const i32 *col_ind, const f64 *A00, const f64 *A01, const f64 *A10,
                           const f64 *A11, f64 alpha, const f64 *x, f64 beta, f64 *y) {
#pragma omp parallel for schedule(static)
  for (i32 i = 0; i < N; i++) {
    i32 start = row_ptr[i], len = row_ptr[i + 1] - start;
    size_t s = (size_t)start;
    const f64 *xu = x, *xp = x + (size_t)3 * N;
    for (int ii = 0; ii < 3; ii++) {
      f64 acc = (beta == 0.0) ? 0.0 : beta * y[(size_t)i * 3 + ii];
      f64 s00 = 0.0, s01 = 0.0;
      for (i32 k = 0; k < len; k++) {
        i32 c = col_ind[start + k];
        for (int l = 0; l < 3; l++) s00 += A00[s * 9 + (size_t)ii * 3 * len + (size_t)k * 3 + l] * xu[(size_t)c * 3 + l];
      }
      acc += alpha * s00; /* block (0,0) first, then (0,1): matrix.c:484-495 loop order */
      for (i32 k = 0; k < len; k++) s01 += A01[s * 3 + (size_t)ii * len + k] * xp[col_ind[start + k]];
      acc += alpha * s01;
      y[(size_t)i * 3 + ii] = acc;
    }
    {
      f64 acc = (beta == 0.0) ? 0.0 : beta * y[(size_t)3 * N + i];
      f64 s10 = 0.0, s11 = 0.0;
      for (i32 k = 0; k < len; k++) {
        i32 c = col_ind[start + k];
        for (int l = 0; l < 3; l++) s10 += A10[s * 3 + (size_t)k * 3 + l] * xu[(size_t)c * 3 + l];
      }
      acc += alpha * s10;
      for (i32 k = 0; k < len; k++) s11 += A11[s + k] * xp[col_ind[start + k]];
      acc += alpha * s11;
      y[(size_t)3 * N + i] = acc;
    }
  }
}

/* ===================================================================================== */
/* a11. preconditioner -- pc.c:44-147, matrix_impl.cu:642-683                              */
/* ===================================================================================== */
/* dinv00[9N]: per node the inverse of the 3x3 diagonal block stored ROW-major and inverted as if it were
 * COLUMN-major, i.e. the buffer holds (B^T)^{-1} column-major (defect D3); dinv11[N] = 1/diag(A11). */
ORC_API void orc_pc_setup(i32 N, const i32 *row_ptr, const i32 *col_ind, const f64 *A00, const f64 *A11, f64 *dinv00,
                          f64 *dinv11) {
#pragma omp parallel for schedule(static)
  for (i32 i = 0; i < N; i++) {
    i32 start = row_ptr[i], end = row_ptr[i + 1], len = end - start, k;
    for (k = start; k < end; k++)
      if (col_ind[k] == i) break;
    f64 M[9]; /* M[r*3+c] = B(r,c) row-major == column-major storage of B^T */
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) M[r * 3 + c] = A00[(size_t)start * 9 + (size_t)(k - start) * 3 + (size_t)r * len * 3 + c];
    /* invert the column-major matrix stored in M with partially pivoted LU (getrf/getri, pc.c:75-77) */
    f64 A[9], inv[9];
    int piv[3];
    memcpy(A, M, sizeof(A));
    for (int kk = 0; kk < 3; kk++) {
      int p = kk;
      f64 amax = fabs(A[kk + 3 * kk]);
      for (int r = kk + 1; r < 3; r++)
        if (fabs(A[r + 3 * kk]) > amax) {
          amax = fabs(A[r + 3 * kk]);
          p = r;
        }
      piv[kk] = p;
      if (p != kk)
        for (int j = 0; j < 3; j++) {
          f64 t = A[kk + 3 * j];
          A[kk + 3 * j] = A[p + 3 * j];
          A[p + 3 * j] = t;
        }
      for (int r = kk + 1; r < 3; r++) {
        A[r + 3 * kk] /= A[kk + 3 * kk];
        for (int j = kk + 1; j < 3; j++) A[r + 3 * j] -= A[r + 3 * kk] * A[kk + 3 * j];
      }
    }
    for (int c = 0; c < 3; c++) {
      f64 b[3] = {0, 0, 0};
      b[c] = 1.0;
      for (int kk = 0; kk < 3; kk++)
        if (piv[kk] != kk) {
          f64 t = b[kk];
          b[kk] = b[piv[kk]];
          b[piv[kk]] = t;
        }
      for (int r = 1; r < 3; r++)
        for (int j = 0; j < r; j++) b[r] -= A[r + 3 * j] * b[j];
      for (int r = 2; r >= 0; r--) {
        for (int j = r + 1; j < 3; j++) b[r] -= A[r + 3 * j] * b[j];
        b[r] /= A[r + 3 * r];
      }
      inv[0 + 3 * c] = b[0];
      inv[1 + 3 * c] = b[1];
      inv[2 + 3 * c] = b[2];
    }
    memcpy(dinv00 + (size_t)i * 9, inv, sizeof(inv));
    dinv11[i] = 1.0 / A11[k];
  }
}

/* PCDecompositionApply: Jacobi(bs=3) on u via column-major gemv (pc.c:104-112), Jacobi(bs=1) on p (pc.c:101),
 * copies for phi and T (pc.c:26). */
ORC_API void orc_pc_apply(i32 N, const f64 *dinv00, const f64 *dinv11, const f64 *x, f64 *y) {
#pragma omp parallel for schedule(static)
  for (i32 i = 0; i < N; i++) {
    const f64 *D = dinv00 + (size_t)i * 9;
    const f64 *xi = x + (size_t)i * 3;
    for (int r = 0; r < 3; r++) y[(size_t)i * 3 + r] = D[r + 0] * xi[0] + D[r + 3] * xi[1] + D[r + 6] * xi[2];
    y[(size_t)3 * N + i] = x[(size_t)3 * N + i] * dinv11[i];
    y[(size_t)4 * N + i] = x[(size_t)4 * N + i];
    y[(size_t)5 * N + i] = x[(size_t)5 * N + i];
  }
}

/* ===================================================================================== */
/* a12. GMRES -- krylov.c:56-334, krylov_util.cu:5-19                                      */
/* ===================================================================================== */
/* deterministic blocked dot product (fixed 4096-element blocks; independent of the thread count) */
static f64 ddot(i64 n, const f64 *a, const f64 *b) {
  const i64 BLK = 4096;
  i64 nb = (n + BLK - 1) / BLK;
  f64 *part = (f64 *)malloc(sizeof(f64) * (size_t)(nb > 0 ? nb : 1));
#pragma omp parallel for schedule(static)
  for (i64 ib = 0; ib < nb; ib++) {
    i64 lo = ib * BLK, hi = lo + BLK < n ? lo + BLK : n;
    f64 s = 0.0;
    for (i64 i = lo; i < hi; i++) s += a[i] * b[i];
    part[ib] = s;
  }
  f64 s = 0.0;
  for (i64 ib = 0; ib < nb; ib++) s += part[ib];
  free(part);
  return s;
}

/* reference BLAS drotg */
static void drotg(f64 *a, f64 *b, f64 *c, f64 *s) {
  f64 roe = fabs(*a) > fabs(*b) ? *a : *b;
  f64 scale = fabs(*a) + fabs(*b);
  f64 r, z;
  if (scale == 0.0) {
    *c = 1.0;
    *s = 0.0;
    r = 0.0;
    z = 0.0;
  } else {
    f64 sa = *a / scale, sb = *b / scale;
    r = scale * sqrt(sa * sa + sb * sb);
    r = (roe < 0.0 ? -1.0 : 1.0) * r;
    *c = *a / r;
    *s = *b / r;
    z = 1.0;
    if (fabs(*a) > fabs(*b)) z = *s;
    if (fabs(*b) >= fabs(*a) && *c != 0.0) z = 1.0 / *c;
  }
  *a = r;
  *b = z;
}

/* Right-preconditioned, un-restarted GMRES(maxit) with single-pass classical Gram-Schmidt and Givens rotations;
 * convergence is tested only when (iter+1)%20==0 (krylov.c:281-290, defect D10).  All vectors have n = 6N entries
 * (matrix.c:408-409); the mat-vec only touches [0,4N) (D4).  res_hist[k] = |beta[k]| after iteration k-1
 * (res_hist[0] = ||r0||).  Returns the number of iterations performed. */
ORC_API i32 orc_gmres(i32 N, const i32 *row_ptr, const i32 *col_ind, const f64 *A00, const f64 *A01, const f64 *A10,
                      const f64 *A11, i32 maxit, f64 atol, f64 rtol, f64 *x, const f64 *b, f64 *res_hist /* maxit+1 */) {
  i64 n = (i64)6 * N;
  i32 ldh = ((maxit + 1 + 31) / 32) * 32;
  f64 *dinv00 = (f64 *)malloc(sizeof(f64) * 9 * (size_t)N), *dinv11 = (f64 *)malloc(sizeof(f64) * (size_t)N);
  orc_pc_setup(N, row_ptr, col_ind, A00, A11, dinv00, dinv11);
  f64 *Q = (f64 *)calloc((size_t)n * (maxit + 1), sizeof(f64));
  f64 *H = (f64 *)calloc((size_t)ldh * maxit, sizeof(f64));
  f64 *tmp = (f64 *)calloc((size_t)n * 2, sizeof(f64));
  f64 *gv = (f64 *)calloc((size_t)2 * maxit, sizeof(f64));
  f64 *beta = (f64 *)calloc((size_t)maxit + 1, sizeof(f64));
#define QCOL(c) (Q + (size_t)(c)*n)
#define HCOL(c) (H + (size_t)(c)*ldh)
  memcpy(QCOL(0), b, sizeof(f64) * (size_t)n);
  orc_fs_amvpby(N, row_ptr, col_ind, A00, A01, A10, A11, -1.0, x, 1.0, QCOL(0));
  f64 rnrm_init = sqrt(ddot(n, QCOL(0), QCOL(0)));
  beta[0] = rnrm_init;
  if (res_hist) res_hist[0] = rnrm_init;
  f64 rnrm = 1.0 / rnrm_init;
  for (i64 i = 0; i < n; i++) QCOL(0)[i] *= rnrm;
  i32 iter = 0;
  int converged = 0;
  while (!converged && iter < maxit) {
    orc_pc_apply(N, dinv00, dinv11, QCOL(iter), tmp);
    orc_fs_amvpby(N, row_ptr, col_ind, A00, A01, A10, A11, 1.0, tmp, 0.0, QCOL(iter + 1));
    f64 *w = QCOL(iter + 1), *h = HCOL(iter);
    for (i32 j = 0; j <= iter; j++) h[j] = ddot(n, QCOL(j), w); /* Dgemv(T) krylov.c:166-174 */
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; i++) { /* Dgemv(N) krylov.c:176-183 */
      f64 s = 0.0;
      for (i32 j = 0; j <= iter; j++) s += QCOL(j)[i] * h[j];
      w[i] -= s;
    }
    h[iter + 1] = sqrt(ddot(n, w, w));
    rnrm = 1.0 / h[iter + 1];
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; i++) w[i] *= rnrm;
    for (i32 i = 0; i < iter; i++) { /* Drot n=1 krylov.c:258-263 */
      f64 c = gv[2 * i], s = gv[2 * i + 1], xx = h[i], yy = h[i + 1];
      h[i] = c * xx + s * yy;
      h[i + 1] = c * yy - s * xx;
    }
    drotg(h + iter, h + iter + 1, gv + 2 * iter, gv + 2 * iter + 1);
    h[iter + 1] = 0.0;
    { /* krylov_util.cu:5-19 */
      f64 b0 = beta[iter];
      beta[iter + 1] = -gv[2 * iter + 1] * b0;
      beta[iter] = b0 * gv[2 * iter];
    }
    if (res_hist) res_hist[iter + 1] = fabs(beta[iter + 1]);
    if ((iter + 1) % 20 == 0) {
      rnrm = fabs(beta[iter + 1]);
      if (rnrm < atol || rnrm < (rnrm_init + 1e-16) * rtol) converged = 1;
    }
    iter++;
  }
  if (iter) {
    for (i32 i = iter - 1; i >= 0; i--) { /* Dtrsv upper, non-unit krylov.c:297-301 */
      f64 s = beta[i];
      for (i32 j = i + 1; j < iter; j++) s -= HCOL(j)[i] * beta[j];
      beta[i] = s / HCOL(i)[i];
    }
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; i++) {
      f64 s = 0.0;
      for (i32 j = 0; j < iter; j++) s += QCOL(j)[i] * beta[j];
      tmp[i] = s;
    }
    orc_pc_apply(N, dinv00, dinv11, tmp, tmp + n);
    for (i64 i = 0; i < n; i++) x[i] += tmp[n + i];
  }
#undef QCOL
#undef HCOL
  free(dinv00);
  free(dinv11);
  free(Q);
  free(H);
  free(tmp);
  free(gv);
  free(beta);
  return iter;
}
