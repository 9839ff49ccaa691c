// elem_math.cuh -- per-element arithmetic of the VMS Navier-Stokes weak form on linear tets, written for
// registers: closed-form geometry, gather, interpolation, residual (F) and the hoisted Jacobian (J) blocks.
//
// What it restates (reference paths relative to /root/reference/src):
//   assemble.cu:321-357,1245-1291  J, J^-1 (batched LU there, cofactors here), detJ = |det J|
//   assemble.cu:1308-1328          shape gradients      sh[a][d] = dN_a/dx_d
//   assemble.cu:1586-1593          metric               G = Jinv * Jinv^T
//   assemble.cu:135-154,1601-1693  nodal gather + interpolation to the 4 quadrature points
//   assemble.cu:444-484,761-924    GetStabTau + AssembleWeakFormKernel<TENSOR=1>   (element residual)
//   assemble.cu:495-759            AssembleWeakFormLHSKernel                      (element Jacobian, u-p 4x4)
//   assemble.cu:279-319,1038-1214  face normal (Nanson) + FaceAssemblyKernel
//
// The reference evaluates the Jacobian with a 4-point quadrature loop per (a,b) pair (~190 flop per (a,b,q)).
// For linear tets grad N is constant and N_a(q) = SB + (SA-SB)*[a==q], so every q-sum collapses to a few
// per-element vectors (P, R, t) -- see JPrep/jac_block below: ~50 flop per (a,b) block instead of ~760.
// The re-association changes results at the 1e-16 level (parity bound: 1e-12).
//
// All functions are __host__ __device__ so that tests can probe the very same arithmetic on the CPU
// (tests/cpu_probe); the product only ever calls them from kernels.
#pragma once
#include <math.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif
#define DFB_HD __host__ __device__ __forceinline__

namespace dfb {
namespace em {

typedef double f64;

constexpr f64 RHOC = 0.5, DT = 5e-2;
constexpr f64 ALPHAM = (3.0 - RHOC) / (1.0 + RHOC);
constexpr f64 ALPHAF = 1.0 / (1.0 + RHOC);
constexpr f64 GAMMA = 0.5 + ALPHAM - ALPHAF;
constexpr f64 RHO = 1.0e3, CP = 1.0, KAPPA = 0.66, MU = 10.0 / 3.0;
constexpr f64 GW = 0.0416666666666667;
constexpr f64 SA = 0.5854101966249685, SB = 0.1381966011250105, SD = SA - SB;
constexpr f64 SN = SA + 3.0 * SB;  // sum_q N_b(q)
constexpr f64 FACT1 = ALPHAM, FACT2 = DT * ALPHAF * GAMMA;
constexpr f64 FB0 = 0.0, FB1 = 0.0, FB2 = -9.81 * 0.0;  // body force, assemble.cu:42
constexpr f64 GWB = 0.1666666666666667;                // face rule, assemble.cu:86
constexpr f64 T6 = 0.1666666666666667, T3 = 0.6666666666666667;

DFB_HD f64 shl(int a, int q) { return a == q ? SA : SB; }
DFB_HD f64 rsqrt_(f64 x) {
#ifdef __CUDA_ARCH__
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}

struct Geom {
  f64 sh[4][3];  // dN_a/dx_d
  f64 inv[3][3]; // Jinv(i,j) = d xi_i / d x_j
  f64 detJ;      // |det J|
};

// x[a][d] : coordinates of the 4 vertices
DFB_HD void geometry(const f64 x[4][3], Geom& g) {
  f64 c0[3], c1[3], c2[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    c0[d] = x[1][d] - x[0][d];
    c1[d] = x[2][d] - x[0][d];
    c2[d] = x[3][d] - x[0][d];
  }
  // rows of the inverse are the cross products of the columns of J divided by det
  f64 r0[3] = {c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]};
  f64 r1[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
  f64 r2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
  f64 det = c0[0] * r0[0] + c0[1] * r0[1] + c0[2] * r0[2];
  f64 idet = 1.0 / det;
  g.detJ = fabs(det);
#pragma unroll
  for (int d = 0; d < 3; d++) {
    g.inv[0][d] = r0[d] * idet;
    g.inv[1][d] = r1[d] * idet;
    g.inv[2][d] = r2[d] * idet;
    g.sh[1][d] = g.inv[0][d];
    g.sh[2][d] = g.inv[1][d];
    g.sh[3][d] = g.inv[2][d];
    g.sh[0][d] = -g.inv[0][d] - g.inv[1][d] - g.inv[2][d];
  }
}

// G(i,j) = sum_d Jinv(i,d) Jinv(j,d)   (assemble.cu:1586-1593)
DFB_HD void metric(const Geom& g, f64 G[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) G[i][j] = g.inv[i][0] * g.inv[j][0] + g.inv[i][1] * g.inv[j][1] + g.inv[i][2] * g.inv[j][2];
}

// ------------------------------------------------------------------------------------------
// residual.  val[comp][a]: nodal values; comp 0..2 = u (wgalpha), 3 = p (dwgalpha slot 3, defect D6),
// 4 = phi, 5 = T (wgalpha);  dval[comp][a]: nodal rates from dwgalpha.  eF[a][ii].
// ------------------------------------------------------------------------------------------
// Hoisted like the Jacobian: for linear tets grad N is constant and N_a(q) = SB + SD [a == q], so
//   sum_q N_a(q) X_q = SB sum_q X_q + SD X_a      and      sum_q (grad N_a . Y_q) = grad N_a . sum_q Y_q .
// The quadrature loop therefore only accumulates a handful of sums (T0, T1, V, B4, B5, ...) and the SD terms of row q; the
// 4 x 6 scatter over the shape functions happens once at the end instead of once per quadrature point (reference
// assemble.cu:761-924 evaluates it inside the loop).  The re-association changes results at the 1e-16 level.
DFB_HD void residual(const Geom& g, const f64 val[6][4], const f64 dval[6][4], f64 eF[4][6]) {
  f64 G[3][3];
  metric(g, G);
  f64 grad[6][3];
#pragma unroll
  for (int c = 0; c < 6; c++)
#pragma unroll
    for (int d = 0; d < 3; d++)
      grad[c][d] = g.sh[0][d] * val[c][0] + g.sh[1][d] * val[c][1] + g.sh[2][d] * val[c][2] + g.sh[3][d] * val[c][3];
  f64 gg = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) gg += G[i][j] * G[i][j];
  const f64 itr = 1.0 / (G[0][0] + G[1][1] + G[2][2]);
  const f64 nu = MU / RHO, al = KAPPA / (RHO * CP);
  const f64 divu = grad[0][0] + grad[1][1] + grad[2][2];
  const f64 t0 = 4.0 / (DT * DT);
  const f64 fb[3] = {FB0, FB1, FB2};
  // nodal sums for the interpolation: value(q) = SB * sum_a v_a + SD * v_q.  Needed: u, p of val; du, dphi, dT of dval.
  f64 sv[4], sd[6];
#pragma unroll
  for (int c = 0; c < 4; c++) sv[c] = SB * (val[c][0] + val[c][1] + val[c][2] + val[c][3]);
#pragma unroll
  for (int c = 0; c < 6; c++) sd[c] = SB * (dval[c][0] + dval[c][1] + dval[c][2] + dval[c][3]);
  f64 T0[3] = {0.0, 0.0, 0.0}, T1[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
  f64 V[3] = {0.0, 0.0, 0.0}, B4[3] = {0.0, 0.0, 0.0}, B5[3] = {0.0, 0.0, 0.0};
  f64 sbp = 0.0, sbtc = 0.0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const f64 u0 = sv[0] + SD * val[0][q], u1 = sv[1] + SD * val[1][q], u2 = sv[2] + SD * val[2][q], pq = sv[3] + SD * val[3][q];
    const f64 dq0 = sd[0] + SD * dval[0][q], dq1 = sd[1] + SD * dval[1][q], dq2 = sd[2] + SD * dval[2][q];
    const f64 dq4 = sd[4] + SD * dval[4][q], dq5 = sd[5] + SD * dval[5][q];
    const f64 uadv[3] = {u0, u1, u2};
    const f64 dqv[3] = {dq0, dq1, dq2};
    f64 rLi[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
      rLi[i] = RHO * (dqv[i] - fb[i]) + RHO * u0 * grad[i][0] + RHO * u1 * grad[i][1] + RHO * u2 * grad[i][2] + grad[3][i];
    // GetStabTau (assemble.cu:444-484)
    f64 t1 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) t1 += G[i][j] * uadv[i] * uadv[j];
    // divisions by constants / by the q-independent trace are multiplications by reciprocals, sqrt(s) = s * rsqrt(s):
    // 1-2 ulp away from the reference's expression (parity bound 1e-12), ~60 fewer FP64 instructions per quadrature point
    const f64 sC = t1 + 3.0 * nu * nu * gg;
    const f64 tauM = rsqrt_(t0 + sC) * (1.0 / RHO);
    const f64 tauC = sC * rsqrt_(sC) * itr;
    const f64 tauP = rsqrt_(t0 + t1);
    const f64 tauT = rsqrt_(t0 + t1 + 3.0 * al * al * gg) * (1.0 / (RHO * CP));
    const f64 pd = -pq + RHO * tauC * divu;
    const f64 bp = dq4 + u0 * grad[4][0] + u1 * grad[4][1] + u2 * grad[4][2];
    const f64 btc = RHO * CP * (dq5 + u0 * grad[5][0] + u1 * grad[5][1] + u2 * grad[5][2]);
    // fine-scale velocity u' = -tauM rLi:  ub = u + u',  T1 += rho tauM rLi (x) ub  (the viscous part of T1 does not depend on q
    // and is added once after the loop)
    f64 tr_[3], ub[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { tr_[i] = tauM * rLi[i]; ub[i] = uadv[i] - tr_[i]; }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const f64 tmp0 = RHO * (dqv[i] - fb[i]) + RHO * ub[0] * grad[i][0] + RHO * ub[1] * grad[i][1] + RHO * ub[2] * grad[i][2];
      T0[i] += tmp0;
      eF[q][i] = SD * tmp0;                       // the SD * X_a term of row a = q
      const f64 ai = RHO * tr_[i];
#pragma unroll
      for (int j = 0; j < 3; j++) T1[i][j] += ai * ub[j] + (i == j ? pd : 0.0);
      V[i] += tr_[i];
      B4[i] += bp * tauP * uadv[i];
      B5[i] += btc * (RHO * CP * tauT) * uadv[i];
    }
    sbp += bp;
    sbtc += btc;
    eF[q][4] = SD * bp;
    eF[q][5] = SD * btc;
  }
  const f64 wdet = GW * g.detJ;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) T1[i][j] += 4.0 * MU * (grad[i][j] + grad[j][i]);   // viscous stress at the 4 points
#pragma unroll
  for (int d = 0; d < 3; d++) B5[d] += 4.0 * KAPPA * grad[5][d];   // kappa grad T . grad N_a at the 4 points
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const f64 s0 = g.sh[a][0], s1 = g.sh[a][1], s2 = g.sh[a][2];
#pragma unroll
    for (int i = 0; i < 3; i++) eF[a][i] = (eF[a][i] + SB * T0[i] + s0 * T1[i][0] + s1 * T1[i][1] + s2 * T1[i][2]) * wdet;
    eF[a][3] = (SN * divu + s0 * V[0] + s1 * V[1] + s2 * V[2]) * wdet;
    eF[a][4] = (eF[a][4] + SB * sbp + s0 * B4[0] + s1 * B4[1] + s2 * B4[2]) * wdet;
    eF[a][5] = (eF[a][5] + SB * sbtc + s0 * B5[0] + s1 * B5[1] + s2 * B5[2]) * wdet;
  }
}

// The same residual with the nodal data read from four NODE RECORDS (shared memory in k_patchF) where it is used, instead of
// 48 values held in registers from the start.  Record layout (NREC doubles):
//   [0..2] x   [3..5] u (wgalpha)   [6..8] du (dwgalpha)   [9] p (dwgalpha slot 3, D6)   [10] phi  [11] dphi  [12] T  [13] dT
// Identical arithmetic, identical order of operations: bit-identical to residual().  The 24 results are written to eF as
// they are produced (the SD * X_q terms inside the quadrature loop, completed at the end), so they never occupy registers.
constexpr int NREC = 14;
DFB_HD void residual_rec(const Geom& g, const f64* n0, const f64* n1, const f64* n2, const f64* n3, f64* eF /* [4*6], may be shared memory */) {
  const f64* nr[4] = {n0, n1, n2, n3};
  // val[c][a]: c = 0..2 -> rec[3+c], 3 -> rec[9], 4 -> rec[10], 5 -> rec[12];  dval[c][a]: c = 0..2 -> rec[6+c], 4 -> rec[11], 5 -> rec[13]
  const int vo[6] = {3, 4, 5, 9, 10, 12};
  const int dvo[6] = {6, 7, 8, 9, 11, 13};
  f64 G[3][3];
  metric(g, G);
  f64 grad[6][3];
#pragma unroll
  for (int c = 0; c < 6; c++) {
    const f64 v0 = nr[0][vo[c]], v1 = nr[1][vo[c]], v2 = nr[2][vo[c]], v3 = nr[3][vo[c]];
#pragma unroll
    for (int d = 0; d < 3; d++) grad[c][d] = g.sh[0][d] * v0 + g.sh[1][d] * v1 + g.sh[2][d] * v2 + g.sh[3][d] * v3;
  }
  f64 gg = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) gg += G[i][j] * G[i][j];
  const f64 itr = 1.0 / (G[0][0] + G[1][1] + G[2][2]);
  const f64 nu = MU / RHO, al = KAPPA / (RHO * CP);
  const f64 divu = grad[0][0] + grad[1][1] + grad[2][2];
  const f64 t0 = 4.0 / (DT * DT);
  const f64 fb[3] = {FB0, FB1, FB2};
  f64 sv[4], sd[6];
#pragma unroll
  for (int c = 0; c < 4; c++) sv[c] = SB * (nr[0][vo[c]] + nr[1][vo[c]] + nr[2][vo[c]] + nr[3][vo[c]]);
#pragma unroll
  for (int c = 0; c < 6; c++) sd[c] = SB * (nr[0][dvo[c]] + nr[1][dvo[c]] + nr[2][dvo[c]] + nr[3][dvo[c]]);
  f64 T0[3] = {0.0, 0.0, 0.0}, T1[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
  f64 V[3] = {0.0, 0.0, 0.0}, B4[3] = {0.0, 0.0, 0.0}, B5[3] = {0.0, 0.0, 0.0};
  f64 sbp = 0.0, sbtc = 0.0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const f64* rq = nr[q];
    const f64 u0 = sv[0] + SD * rq[3], u1 = sv[1] + SD * rq[4], u2 = sv[2] + SD * rq[5], pq = sv[3] + SD * rq[9];
    const f64 dq0 = sd[0] + SD * rq[6], dq1 = sd[1] + SD * rq[7], dq2 = sd[2] + SD * rq[8];
    const f64 dq4 = sd[4] + SD * rq[11], dq5 = sd[5] + SD * rq[13];
    const f64 uadv[3] = {u0, u1, u2};
    const f64 dqv[3] = {dq0, dq1, dq2};
    f64 rLi[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
      rLi[i] = RHO * (dqv[i] - fb[i]) + RHO * u0 * grad[i][0] + RHO * u1 * grad[i][1] + RHO * u2 * grad[i][2] + grad[3][i];
    f64 t1 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) t1 += G[i][j] * uadv[i] * uadv[j];
    const f64 sC = t1 + 3.0 * nu * nu * gg;
    const f64 tauM = rsqrt_(t0 + sC) * (1.0 / RHO);
    const f64 tauC = sC * rsqrt_(sC) * itr;
    const f64 tauP = rsqrt_(t0 + t1);
    const f64 tauT = rsqrt_(t0 + t1 + 3.0 * al * al * gg) * (1.0 / (RHO * CP));
    const f64 pd = -pq + RHO * tauC * divu;
    const f64 bp = dq4 + u0 * grad[4][0] + u1 * grad[4][1] + u2 * grad[4][2];
    const f64 btc = RHO * CP * (dq5 + u0 * grad[5][0] + u1 * grad[5][1] + u2 * grad[5][2]);
    f64 tr_[3], ub[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { tr_[i] = tauM * rLi[i]; ub[i] = uadv[i] - tr_[i]; }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const f64 tmp0 = RHO * (dqv[i] - fb[i]) + RHO * ub[0] * grad[i][0] + RHO * ub[1] * grad[i][1] + RHO * ub[2] * grad[i][2];
      T0[i] += tmp0;
      eF[q * 6 + i] = SD * tmp0;
      const f64 ai = RHO * tr_[i];
#pragma unroll
      for (int j = 0; j < 3; j++) T1[i][j] += ai * ub[j] + (i == j ? pd : 0.0);
      V[i] += tr_[i];
      B4[i] += bp * tauP * uadv[i];
      B5[i] += btc * (RHO * CP * tauT) * uadv[i];
    }
    sbp += bp;
    sbtc += btc;
    eF[q * 6 + 4] = SD * bp;
    eF[q * 6 + 5] = SD * btc;
  }
  const f64 wdet = GW * g.detJ;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) T1[i][j] += 4.0 * MU * (grad[i][j] + grad[j][i]);
#pragma unroll
  for (int d = 0; d < 3; d++) B5[d] += 4.0 * KAPPA * grad[5][d];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const f64 s0 = g.sh[a][0], s1 = g.sh[a][1], s2 = g.sh[a][2];
#pragma unroll
    for (int i = 0; i < 3; i++) eF[a * 6 + i] = (eF[a * 6 + i] + SB * T0[i] + s0 * T1[i][0] + s1 * T1[i][1] + s2 * T1[i][2]) * wdet;
    eF[a * 6 + 3] = (SN * divu + s0 * V[0] + s1 * V[1] + s2 * V[2]) * wdet;
    eF[a * 6 + 4] = (eF[a * 6 + 4] + SB * sbp + s0 * B4[0] + s1 * B4[1] + s2 * B4[2]) * wdet;
    eF[a * 6 + 5] = (eF[a * 6 + 5] + SB * sbtc + s0 * B5[0] + s1 * B5[1] + s2 * B5[2]) * wdet;
  }
}

// ------------------------------------------------------------------------------------------
// Jacobian, hoisted form.
// ------------------------------------------------------------------------------------------
struct JPrep {
  f64 w;         // detJ * gw
  f64 c[4][4];   // c[q][a] = u(q) . grad N_a          (shconv of assemble.cu:574-583)
  f64 t[4][4];   // t[q][a] = tauM(q) * c[q][a]
  f64 tM[4];     // tauM(q)
  f64 P[4];      // P[a] = sum_q t[q][a]
  f64 R[4];      // R[b] = sum_q c[q][b]
  f64 sTM, sTC;  // sum_q tauM(q), sum_q tauC(q)
};

// u[a][d]: nodal advection velocity (wgalpha u-part)
DFB_HD void jac_prep(const Geom& g, const f64 u[4][3], JPrep& p) {
  f64 G[3][3];
  metric(g, G);
  f64 gg = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) gg += G[i][j] * G[i][j];
  const f64 itr = 1.0 / (G[0][0] + G[1][1] + G[2][2]);
  const f64 nu = MU / RHO;
  p.w = g.detJ * GW;
  p.sTM = 0.0;
  p.sTC = 0.0;
#pragma unroll
  for (int a = 0; a < 4; a++) { p.P[a] = 0.0; p.R[a] = 0.0; }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    f64 uq[3];
#pragma unroll
    for (int d = 0; d < 3; d++) uq[d] = shl(0, q) * u[0][d] + shl(1, q) * u[1][d] + shl(2, q) * u[2][d] + shl(3, q) * u[3][d];
#pragma unroll
    for (int a = 0; a < 4; a++) p.c[q][a] = g.sh[a][0] * uq[0] + g.sh[a][1] * uq[1] + g.sh[a][2] * uq[2];
    // tau of the LHS kernel: sum_{a=1..3} (u.gradN_a)^2 instead of u.G.u (defect D5), assemble.cu:592-602
    const f64 tmp = p.c[q][1] * p.c[q][1] + p.c[q][2] * p.c[q][2] + p.c[q][3] * p.c[q][3];
    const f64 sC = tmp + 3.0 * nu * nu * gg;
    const f64 tauM = rsqrt_(4.0 / (DT * DT) + sC) * (1.0 / RHO);
    const f64 tauC = sC * rsqrt_(sC) * itr;
    p.tM[q] = tauM;
    p.sTM += tauM;
    p.sTC += tauC;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      p.t[q][a] = tauM * p.c[q][a];
      p.P[a] += p.t[q][a];
      p.R[a] += p.c[q][a];
    }
  }
}

// branch-free 4-way select: keeps per-thread arrays in registers when the index is a run-time value
DFB_HD f64 sel4(int i, f64 v0, f64 v1, f64 v2, f64 v3) {
  f64 lo = (i & 1) ? v1 : v0, hi = (i & 1) ? v3 : v2;
  return (i & 2) ? hi : lo;
}

// everything of JPrep/Geom that is indexed by the ROW node a of a block (a is a run-time value per thread)
struct ARow {
  int a;
  f64 ga[3];    // grad N_a
  f64 cqa[4];   // c[q][a]
  f64 tqa[4];   // t[q][a]
  f64 cab[4];   // c[a][b]   (a used as quadrature index)
  f64 Pa;       // P[a]
};

DFB_HD void extract_row(const Geom& g, const JPrep& p, int a, ARow& r) {
  r.a = a;
#pragma unroll
  for (int d = 0; d < 3; d++) r.ga[d] = sel4(a, g.sh[0][d], g.sh[1][d], g.sh[2][d], g.sh[3][d]);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    r.cqa[q] = sel4(a, p.c[q][0], p.c[q][1], p.c[q][2], p.c[q][3]);
    r.tqa[q] = sel4(a, p.t[q][0], p.t[q][1], p.t[q][2], p.t[q][3]);
    r.cab[q] = sel4(a, p.c[0][q], p.c[1][q], p.c[2][q], p.c[3][q]);
  }
  r.Pa = sel4(a, p.P[0], p.P[1], p.P[2], p.P[3]);
}

// 4x4 (u,p) block of node pair (a,b): blk[ii*4+jj].  b must be a compile-time constant after unrolling.
DFB_HD void jac_block_row(const Geom& g, const JPrep& p, const ARow& r, int b, f64 blk[16]) {
  const f64* ga = r.ga;
  const f64* gb = g.sh[b];
  const f64 eK = ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2];
  const f64 mab = (r.a == b) ? (SA * SA + 3.0 * SB * SB) : (2.0 * SA * SB + 2.0 * SB * SB);
  const f64 stc = r.tqa[0] * p.c[0][b] + r.tqa[1] * p.c[1][b] + r.tqa[2] * p.c[2][b] + r.tqa[3] * p.c[3][b];
  const f64 T = p.w * (FACT1 * RHO * mab + FACT1 * RHO * RHO * (SB * r.Pa + SD * r.tqa[b]) +
                       FACT2 * RHO * (SB * p.R[b] + SD * r.cab[b]) + FACT2 * RHO * RHO * stc + 4.0 * FACT2 * MU * eK);
  const f64 k1 = 4.0 * p.w * FACT2 * MU;          // viscous transpose term
  const f64 k2 = p.w * FACT2 * RHO * p.sTC;       // grad-div (tauC) term
#pragma unroll
  for (int ii = 0; ii < 3; ii++)
#pragma unroll
    for (int jj = 0; jj < 3; jj++) blk[ii * 4 + jj] = k1 * ga[jj] * gb[ii] + k2 * ga[ii] * gb[jj] + (ii == jj ? T : 0.0);
  const f64 k3 = p.w * SN, k4 = RHO * p.w * r.Pa;
  const f64 k5 = p.w * RHO * (FACT1 * (SB * p.sTM + SD * p.tM[b]) + FACT2 * p.P[b]);
  const f64 k6 = FACT2 * p.w * SN;
#pragma unroll
  for (int ii = 0; ii < 3; ii++) {
    blk[ii * 4 + 3] = -k3 * ga[ii] + k4 * gb[ii];   // dRM/dP
    blk[3 * 4 + ii] = k5 * ga[ii] + k6 * gb[ii];    // dRC/dU
  }
  blk[15] = p.w * p.sTM * eK;                       // dRC/dP
}

DFB_HD void jac_block(const Geom& g, const JPrep& p, int a, int b, f64 blk[16]) {
  ARow r;
  extract_row(g, p, a, r);
  switch (b) {
    case 0: jac_block_row(g, p, r, 0, blk); break;
    case 1: jac_block_row(g, p, r, 1, blk); break;
    case 2: jac_block_row(g, p, r, 2, blk); break;
    default: jac_block_row(g, p, r, 3, blk); break;
  }
}

// ------------------------------------------------------------------------------------------
// Jacobian by node PAIRS (assemble.cu k_pairJ).  The blocks (a,b) and (b,a) of one element share their inputs and most of
// their arithmetic: the 3x3 parts are transposes of each other apart from the diagonal term T, and eK, stc, mab are
// symmetric.  One accumulator set of 24 doubles therefore carries BOTH nodal nonzeros (i,j) and (j,i):
//   acc[0..8]  S[ii][jj] = sum k1 ga[jj] gb[ii] + k2 ga[ii] gb[jj]      (A_ij 3x3 = S + T_ab I,  A_ji 3x3 = S^T + T_ba I)
//   acc[9] T_ab   acc[10] T_ba   acc[11..13] col_ab   acc[14..16] row_ab   acc[17..19] col_ba   acc[20..22] row_ba
//   acc[23] d = sum w sTM eK   (the (3,3) entry of both blocks)
// Element record (JREC doubles): corner x at [10x, 10x+10): g0 g1 g2 P c0 c1 c2 c3 R sTC  (c_q = u(q).grad N_x),
//                                tail at [40,46): w sTM tM0 tM1 tM2 tM3.
// ------------------------------------------------------------------------------------------
constexpr int JREC = 46;

DFB_HD void jrec_store(const Geom& g, const JPrep& p, f64* rec) {
#pragma unroll
  for (int a = 0; a < 4; a++) {
    rec[a * 10 + 0] = g.sh[a][0]; rec[a * 10 + 1] = g.sh[a][1]; rec[a * 10 + 2] = g.sh[a][2]; rec[a * 10 + 3] = p.P[a];
    rec[a * 10 + 4] = p.c[0][a]; rec[a * 10 + 5] = p.c[1][a]; rec[a * 10 + 6] = p.c[2][a]; rec[a * 10 + 7] = p.c[3][a];
    rec[a * 10 + 8] = p.R[a]; rec[a * 10 + 9] = p.sTC;
  }
  rec[40] = p.w; rec[41] = p.sTM; rec[42] = p.tM[0]; rec[43] = p.tM[1]; rec[44] = p.tM[2]; rec[45] = p.tM[3];
}

// A, B: the corner sub-records of local nodes a and b (a != b), T: the tail
DFB_HD void jrec_pair(const f64 A[10], const f64 B[10], const f64 T[6], int a, int b, f64 acc[24]) {
  const f64 w = T[0], sTM = T[1];
  const f64 ta0 = T[2] * A[4], ta1 = T[3] * A[5], ta2 = T[4] * A[6], ta3 = T[5] * A[7];
  const f64 stc = ta0 * B[4] + ta1 * B[5] + ta2 * B[6] + ta3 * B[7];
  const f64 eK = A[0] * B[0] + A[1] * B[1] + A[2] * B[2];
  const f64 tab = sel4(b, ta0, ta1, ta2, ta3);                 // tM[b] c[b][a]
  const f64 cab = sel4(a, B[4], B[5], B[6], B[7]);             // c[q=a][b]
  const f64 tMa = sel4(a, T[2], T[3], T[4], T[5]), tMb = sel4(b, T[2], T[3], T[4], T[5]);
  const f64 tba = tMa * cab;                                   // tM[a] c[a][b]
  const f64 cba = sel4(b, A[4], A[5], A[6], A[7]);             // c[q=b][a]
  const f64 common = FACT1 * RHO * (2.0 * SA * SB + 2.0 * SB * SB) + FACT2 * RHO * RHO * stc + 4.0 * FACT2 * MU * eK;
  acc[9] += w * (common + FACT1 * RHO * RHO * (SB * A[3] + SD * tab) + FACT2 * RHO * (SB * B[8] + SD * cab));
  acc[10] += w * (common + FACT1 * RHO * RHO * (SB * B[3] + SD * tba) + FACT2 * RHO * (SB * A[8] + SD * cba));
  const f64 k1 = 4.0 * w * FACT2 * MU, k2 = w * FACT2 * RHO * A[9];
  const f64 k1g[3] = {k1 * A[0], k1 * A[1], k1 * A[2]}, k2g[3] = {k2 * A[0], k2 * A[1], k2 * A[2]};
#pragma unroll
  for (int ii = 0; ii < 3; ii++)
#pragma unroll
    for (int jj = 0; jj < 3; jj++) acc[ii * 3 + jj] += k1g[jj] * B[ii] + k2g[ii] * B[jj];
  const f64 k3 = w * SN, k6 = FACT2 * w * SN;
  const f64 k4a = RHO * w * A[3], k4b = RHO * w * B[3];
  const f64 k5b = w * RHO * (FACT1 * (SB * sTM + SD * tMb) + FACT2 * B[3]);
  const f64 k5a = w * RHO * (FACT1 * (SB * sTM + SD * tMa) + FACT2 * A[3]);
#pragma unroll
  for (int ii = 0; ii < 3; ii++) {
    acc[11 + ii] += -k3 * A[ii] + k4a * B[ii];   // block (a,b) column 3
    acc[14 + ii] += k5b * A[ii] + k6 * B[ii];    // block (a,b) row 3
    acc[17 + ii] += -k3 * B[ii] + k4b * A[ii];   // block (b,a) column 3
    acc[20 + ii] += k5a * B[ii] + k6 * A[ii];    // block (b,a) row 3
  }
  acc[23] += w * sTM * eK;
}

// the two 4x4 blocks of a pair accumulator: ab[ii*4+jj] = A_ij, ba = A_ji
DFB_HD void jrec_pair_blocks(const f64 acc[24], f64 ab[16], f64 ba[16]) {
#pragma unroll
  for (int ii = 0; ii < 3; ii++) {
#pragma unroll
    for (int jj = 0; jj < 3; jj++) {
      ab[ii * 4 + jj] = acc[ii * 3 + jj] + (ii == jj ? acc[9] : 0.0);
      ba[ii * 4 + jj] = acc[jj * 3 + ii] + (ii == jj ? acc[10] : 0.0);
    }
    ab[ii * 4 + 3] = acc[11 + ii]; ab[12 + ii] = acc[14 + ii];
    ba[ii * 4 + 3] = acc[17 + ii]; ba[12 + ii] = acc[20 + ii];
  }
  ab[15] = acc[23];
  ba[15] = acc[23];
}

// diagonal contribution (a,a) of one element from the corner sub-record and the tail, added to acc[ii*4+jj]
DFB_HD void jrec_diag(const f64 A[10], const f64 T[6], int a, f64 acc[16]) {
  const f64 w = T[0], sTM = T[1];
  const f64 stc = (T[2] * A[4]) * A[4] + (T[3] * A[5]) * A[5] + (T[4] * A[6]) * A[6] + (T[5] * A[7]) * A[7];
  const f64 eK = A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
  const f64 caa = sel4(a, A[4], A[5], A[6], A[7]);
  const f64 tMa = sel4(a, T[2], T[3], T[4], T[5]);
  const f64 Tt = w * (FACT1 * RHO * (SA * SA + 3.0 * SB * SB) + FACT1 * RHO * RHO * (SB * A[3] + SD * (tMa * caa)) +
                      FACT2 * RHO * (SB * A[8] + SD * caa) + FACT2 * RHO * RHO * stc + 4.0 * FACT2 * MU * eK);
  const f64 k12 = 4.0 * w * FACT2 * MU + w * FACT2 * RHO * A[9];
#pragma unroll
  for (int ii = 0; ii < 3; ii++)
#pragma unroll
    for (int jj = 0; jj < 3; jj++) acc[ii * 4 + jj] += (k12 * A[ii]) * A[jj] + (ii == jj ? Tt : 0.0);
  const f64 kc = RHO * w * A[3] - w * SN;
  const f64 kr = w * RHO * (FACT1 * (SB * sTM + SD * tMa) + FACT2 * A[3]) + FACT2 * w * SN;
#pragma unroll
  for (int ii = 0; ii < 3; ii++) {
    acc[ii * 4 + 3] += kc * A[ii];
    acc[12 + ii] += kr * A[ii];
  }
  acc[15] += w * sTM * eK;
}

// ------------------------------------------------------------------------------------------
// boundary face (weak BC).  iorn = local index of the vertex opposite the face.
// val[comp][a]: comp 0..2 = u (wgalpha), 3 = p (dwgalpha slot 3).
// ------------------------------------------------------------------------------------------
DFB_HD f64 shlub(int iorn, int q, int a) {
  // c_shlub[iorn*12 + q*4 + a], assemble.cu:87-102: 0 on the opposite vertex, 2/3 on the "peak" face vertex of
  // quadrature point q, 1/6 on the other two.  Peak vertex table [iorn][q] = {3,2,1},{3,2,0},{0,1,3},{1,2,0},
  // packed 2 bits per entry.
  if (a == iorn) return 0.0;
  const unsigned PK = 0x2742dbu;
  int pk = (int)((PK >> (2 * (iorn * 3 + q))) & 3u);
  return a == pk ? T3 : T6;
}

struct FacePrep {
  f64 nv[3];
  f64 tau_b;
};

DFB_HD void face_prep(const Geom& g, int iorn, FacePrep& f) {
  // c_nv2 (assemble.cu:114-118) and Nanson: nv = detJ * Jinv^T n_ref
  const f64 nref[3] = {iorn == 0 ? 1.0 : (iorn == 1 ? -1.0 : 0.0), iorn == 0 ? 1.0 : (iorn == 2 ? -1.0 : 0.0),
                       iorn == 0 ? 1.0 : (iorn == 3 ? -1.0 : 0.0)};
#pragma unroll
  for (int n = 0; n < 3; n++) f.nv[n] = (g.inv[0][n] * nref[0] + g.inv[1][n] * nref[1] + g.inv[2][n] * nref[2]) * g.detJ;
  f64 h = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    f64 v = g.inv[i][0] * f.nv[0] + g.inv[i][1] * f.nv[1] + g.inv[i][2] * f.nv[2];
    h += v * v;
  }
  f.tau_b = 4.0 * MU * sqrt(h);
}

DFB_HD void face_residual(const Geom& g, const FacePrep& f, int iorn, const f64 val[4][4], f64 eF[4][6]) {
  f64 grad[4][3];
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int d = 0; d < 3; d++)
      grad[c][d] = g.sh[0][d] * val[c][0] + g.sh[1][d] * val[c][1] + g.sh[2][d] * val[c][2] + g.sh[3][d] * val[c][3];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int i = 0; i < 6; i++) eF[a][i] = 0.0;
  const f64* nv = f.nv;
  for (int q = 0; q < 3; q++) {
    f64 N[4], vq[4];
#pragma unroll
    for (int a = 0; a < 4; a++) N[a] = shlub(iorn, q, a);
#pragma unroll
    for (int c = 0; c < 4; c++) vq[c] = N[0] * val[c][0] + N[1] * val[c][1] + N[2] * val[c][2] + N[3] * val[c][3];
    const f64 unor = vq[0] * nv[0] + vq[1] * nv[1] + vq[2] * nv[2];
    const f64 uneg = (unor - fabs(unor)) * 0.5;
    f64 tmp0[3], tmp1[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      tmp0[i] = nv[i] * vq[3] - MU * (nv[0] * grad[i][0] + nv[1] * grad[i][1] + nv[2] * grad[i][2]) -
                MU * (nv[0] * grad[0][i] + nv[1] * grad[1][i] + nv[2] * grad[2][i]) - RHO * uneg * vq[i] + f.tau_b * vq[i];
#pragma unroll
      for (int j = 0; j < 3; j++) tmp1[i][j] = -MU * (nv[i] * vq[j] + nv[j] * vq[i]);
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        f64 bm = N[a] * tmp0[i] + g.sh[a][0] * tmp1[i][0] + g.sh[a][1] * tmp1[i][1] + g.sh[a][2] * tmp1[i][2];
        eF[a][i] += bm * GWB;
      }
      eF[a][3] -= N[a] * unor * GWB;
    }
  }
}

// 4x4 block of pair (a,b) of the face Jacobian (assemble.cu:1127-1192); u[a][d] nodal velocity
DFB_HD void face_block(const Geom& g, const FacePrep& f, int iorn, const f64 u[4][3], int a, int b, f64 blk[16]) {
  const f64* nv = f.nv;
  const f64 sna = g.sh[a][0] * nv[0] + g.sh[a][1] * nv[1] + g.sh[a][2] * nv[2];
  const f64 snb = g.sh[b][0] * nv[0] + g.sh[b][1] * nv[1] + g.sh[b][2] * nv[2];
#pragma unroll
  for (int i = 0; i < 16; i++) blk[i] = 0.0;
  for (int q = 0; q < 3; q++) {
    f64 N[4];
#pragma unroll
    for (int k = 0; k < 4; k++) N[k] = shlub(iorn, q, k);
    f64 uq[3];
#pragma unroll
    for (int d = 0; d < 3; d++) uq[d] = N[0] * u[0][d] + N[1] * u[1][d] + N[2] * u[2][d] + N[3] * u[3][d];
    const f64 unor = uq[0] * nv[0] + uq[1] * nv[1] + uq[2] * nv[2];
    const f64 uneg = (unor - fabs(unor)) * 0.5;
    const f64 Na = N[a], Nb = N[b];
    f64 t0 = -MU * (snb * Na + sna * Nb) - RHO * Na * Nb * uneg + f.tau_b * Na * Nb;
    const f64 d = FACT2 * t0 * GWB;
#pragma unroll
    for (int ii = 0; ii < 3; ii++)
#pragma unroll
      for (int jj = 0; jj < 3; jj++) {
        f64 t = -MU * Na * g.sh[b][ii] * nv[jj] - MU * Nb * g.sh[a][jj] * nv[ii];
        blk[ii * 4 + jj] += FACT2 * t * GWB + (ii == jj ? d : 0.0);
      }
    const f64 nn = Na * Nb;
#pragma unroll
    for (int ii = 0; ii < 3; ii++) {
      blk[3 * 4 + ii] -= FACT2 * nn * nv[ii] * GWB;
      blk[ii * 4 + 3] += nn * nv[ii] * GWB;
    }
  }
}

}  // namespace em
}  // namespace dfb
