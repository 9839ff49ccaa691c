// CPU probe of dedflow_b200/csrc/elem_math.cuh: evaluates the SAME __host__ __device__ element arithmetic the
// kernels use, on the host, so that the hoisted Jacobian / residual formulas can be checked against the oracle
// without a GPU (pytest -m "not gpu").  Test infrastructure only; never part of libdedflow_b200.so.
#include <cstdint>
#include <cstddef>
#include "../../dedflow_b200/csrc/elem_math.cuh"
using namespace dfb::em;

static void load(int e, int N, const int32_t* ien, const double* xg, const double* wg, const double* dwg,
                 double x[4][3], double val[6][4], double dval[6][4]) {
  for (int a = 0; a < 4; a++) {
    int n = ien[(size_t)e * 4 + a];
    for (int d = 0; d < 3; d++) { x[a][d] = xg[(size_t)n * 3 + d]; val[d][a] = wg[(size_t)n * 3 + d]; dval[d][a] = dwg[(size_t)n * 3 + d]; }
    val[3][a] = dwg[(size_t)3 * N + n]; val[4][a] = wg[(size_t)4 * N + n]; val[5][a] = wg[(size_t)5 * N + n];
    dval[3][a] = dwg[(size_t)3 * N + n]; dval[4][a] = dwg[(size_t)4 * N + n]; dval[5][a] = dwg[(size_t)5 * N + n];
  }
}

extern "C" void probe_tet(int n, int N, const int32_t* ien, const double* xg, const double* wg, const double* dwg,
                          double* eF /* n*24 */, double* eJ /* n*16*16: [a][b][ii*4+jj] */) {
  for (int e = 0; e < n; e++) {
    double x[4][3], val[6][4], dval[6][4], u[4][3];
    load(e, N, ien, xg, wg, dwg, x, val, dval);
    Geom g; geometry(x, g);
    if (eF) { double f[4][6]; residual(g, val, dval, f); for (int a = 0; a < 4; a++) for (int i = 0; i < 6; i++) eF[(size_t)e * 24 + a * 6 + i] = f[a][i]; }
    if (eJ) {
      for (int a = 0; a < 4; a++) for (int d = 0; d < 3; d++) u[a][d] = val[d][a];
      JPrep p; jac_prep(g, u, p);
      for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) jac_block(g, p, a, b, eJ + ((size_t)e * 16 + a * 4 + b) * 16);
    }
  }
}

// the pair formulation of the Jacobian (k_pairJ): element record -> pair accumulators -> the same [a][b][ii*4+jj] layout
extern "C" void probe_tet_pairs(int n, int N, const int32_t* ien, const double* xg, const double* wg, const double* dwg,
                                double* eJ /* n*16*16 */) {
  for (int e = 0; e < n; e++) {
    double x[4][3], val[6][4], dval[6][4], u[4][3], rec[JREC];
    load(e, N, ien, xg, wg, dwg, x, val, dval);
    Geom g; geometry(x, g);
    for (int a = 0; a < 4; a++) for (int d = 0; d < 3; d++) u[a][d] = val[d][a];
    JPrep p; jac_prep(g, u, p);
    jrec_store(g, p, rec);
    for (int a = 0; a < 4; a++) {
      double* daa = eJ + ((size_t)e * 16 + a * 4 + a) * 16;
      for (int v = 0; v < 16; v++) daa[v] = 0.0;
      jrec_diag(rec + a * 10, rec + 40, a, daa);
      for (int b = a + 1; b < 4; b++) {
        double acc[24];
        for (int v = 0; v < 24; v++) acc[v] = 0.0;
        jrec_pair(rec + a * 10, rec + b * 10, rec + 40, a, b, acc);
        jrec_pair_blocks(acc, eJ + ((size_t)e * 16 + a * 4 + b) * 16, eJ + ((size_t)e * 16 + b * 4 + a) * 16);
      }
    }
  }
}

extern "C" void probe_face(int nf, const int32_t* f2e, const int32_t* forn, int N, const int32_t* ien, const double* xg,
                           const double* wg, const double* dwg, double* eF, double* eJ) {
  for (int f = 0; f < nf; f++) {
    double x[4][3], val[6][4], dval[6][4], u[4][3], v4[4][4];
    load(f2e[f], N, ien, xg, wg, dwg, x, val, dval);
    Geom g; geometry(x, g);
    FacePrep fp; face_prep(g, forn[f], fp);
    for (int c = 0; c < 4; c++) for (int a = 0; a < 4; a++) v4[c][a] = val[c][a];
    for (int a = 0; a < 4; a++) for (int d = 0; d < 3; d++) u[a][d] = val[d][a];
    if (eF) { double r[4][6]; face_residual(g, fp, forn[f], v4, r); for (int a = 0; a < 4; a++) for (int i = 0; i < 6; i++) eF[(size_t)f * 24 + a * 6 + i] = r[a][i]; }
    if (eJ) for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) face_block(g, fp, forn[f], u, a, b, eJ + ((size_t)f * 16 + a * 4 + b) * 16);
  }
}
