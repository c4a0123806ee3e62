// rt1q.cuh -- sm_100a kernels for `-ov 1 -op 1 -os 1` (RT1 / L2-Q1 velocity-pressure, saturation in DG-Q1:
// BASELINE config C5).  With S in DG-Q1 the mobility varies inside a cell, so
//   * the RT1 mass matrix loses its line structure: M_d|cell = (geo_d / K_d) Phi^T diag(w_q / lambda(S(x_q))) Phi
//     is a dense 12x12 block per direction and the quadrature rule (4^3 Gauss points in 3-D, 3^2 in 2-D:
//     order OrderW + 2 el.order [MFEM-recall]) is part of the operator's definition;
//   * the saturation operator gets a volume term and face integrals with upwinding per face quadrature point
//     (SURVEY A.5): Kt = Convection(v,1)^T - DGTrace(v,1,1/2) [REF src/saturation_solver.cpp:32-35,53-55].
// B, the right-hand side, S = B diag(M)^-1 B^T and both Schur preconditioners keep the structure of rt1.cuh:
// only diag(M) (hence T = 1/diag(M)) changes.
//
// Parity: these kernels evaluate the operators matrix-free in sum-factorised form, the assembled-CSR checker builds
// element matrices entry by entry; those two agree to round-off (tests: 1e-12 relative per operator application).
// The checker's canonical arithmetic for this path is a matrix-free mirror that performs the operations below in
// the same order, and the tests compare against it bit for bit -- so the order of every sum in this file is part
// of a contract: change it together with the mirror.
// FMA contraction is off for the file (arithmetic contract of kernels.cuh); the dense element products use
// explicit __fma_rn (mirrored by fma() on the CPU).
#pragma once
#include "rt1.cuh"

namespace impes {

struct QTab {
  double gw[4];      // mass rule on [0,1]: 4 points (3-D) / 3 points (2-D)
  double C[4][3];    // closed basis (Gauss-Lobatto nodes 0, 1/2, 1) at the mass-rule points
  double O[4][2];    // open / L2 / DG basis (the 2 Gauss nodes) at the mass-rule points
  double fw[3];      // face rule of the DG trace integrator: 3 points (3-D) / 2 points (2-D)
  double OF[3][2];   // open basis at the face-rule points
  double LC[2][3];   // closed basis at the two DG nodes (volume term: collocated 2-point rule, exact here)
  double E0[2], E1[2];  // open basis at xi = 0 and xi = 1 (element traces)
  double dl[2];      // derivative of the open basis (constant): -sqrt 3, +sqrt 3
  double sn[2];      // the DG / L2 nodes 1/2 -+ 1/(2 sqrt 3)
  double wv;         // weight of a DG node in the collocated rule: (1/2)^dim
  double OO[4][4][4];  // 3-D: O[qe][te] * O[qf][tf] at [qe][qf][te + 2 tf] (host-computed: the same IEEE products)
  double OOF[3][3][4]; // 3-D: OF[qe][te] * OF[qf][tf] at [qe][qf][te + 2 tf]
  // RT0 velocity with S in DG-Q1 (-ov 0 -op 0 -os 1): mass rule of order OrderW + 2 = 3 points (3-D) / 2 points (2-D)
  double x0[3], gw0[3], O0[3][2];   // points on [0,1], weights, DG basis at the points
};

// 3-D products of the transverse open bases (QTab::OO / QTab::OOF) in constant memory: the rolled quadrature loops
// index them with a run-time point index (a dynamically indexed kernel parameter would be copied to local memory)
__constant__ double c_OO3[4][4][4];
__constant__ double c_OOF3[3][3][4];
__constant__ double c_WF3[3][3];   // fw[qe] * fw[qf] of the 3-D face rule
__constant__ double c_OO3SQ[4][4][4];  // 3-D: (O[qe][te] O[qe][te]) * (O[qf][tf] O[qf][tf]) at [qe][qf][te + 2 tf] (k1q_md_a)

template <bool HY>
struct QD {
  static constexpr int NQ = HY ? 4 : 3;     // mass-rule points per direction
  static constexpr int NQY = HY ? 4 : 1;
  static constexpr int NQP = NQ * NQY * NQ; // per cell
  static constexpr int NSL = HY ? 8 : 4;    // DG-Q1 / L2-Q1 dofs per cell
  static constexpr int NL = HY ? 4 : 2;     // lines per direction and cell
  static constexpr int NDIR = HY ? 3 : 2;
  static constexpr int NUP = NDIR * NL;     // upper-face-node contributions per cell
  static constexpr int NF = HY ? 3 : 2;     // face-rule points per transverse direction
};

// open-basis value / weight along internal direction `dir` (2-D: the absent y direction has one node, basis 1)
template <bool HY>
__device__ __forceinline__ double ob(const QTab& tb, int dir, int q, int t) { return (!HY && dir == 1) ? 1.0 : tb.O[q][t]; }

// ---------------------------------------------------------------------------------------
// K4 (os = 1): w_q / lambda(S(x_q)) at the mass-rule points of every cell (ghost planes included)
// CWVectorCoefficient::Eval with a DG-Q1 GridFunctionCoefficient [REF src/param.hpp:246-261]
// ---------------------------------------------------------------------------------------
// One thread per (cell, point): coalesced stores of the cell-major point array (the one-thread-per-cell form wrote 64
// doubles per lane with a 512-byte stride between lanes).  Same sums in the same order as before.
template <bool HY>
__global__ void __launch_bounds__(BLK) k1q_rlam(Lay L, Phys ph, QTab tb, const double* __restrict__ S,
                                                double* __restrict__ rlam) {
  using Q = QD<HY>;
  __shared__ double s_O[4][2], s_gw[4];   // run-time point indices: shared-memory copies of the tables
  if (threadIdx.x < 8) s_O[threadIdx.x >> 1][threadIdx.x & 1] = (threadIdx.x & 1) ? tb.O[threadIdx.x >> 1][1] : tb.O[threadIdx.x >> 1][0];
  if (threadIdx.x >= 32 && threadIdx.x < 36) {
    const int q = threadIdx.x - 32;
    s_gw[q] = q == 0 ? tb.gw[0] : (q == 1 ? tb.gw[1] : (q == 2 ? tb.gw[2] : tb.gw[3]));
  }
  __syncthreads();
  const long e = (long)blockIdx.x * BLK + threadIdx.x;
  if (e >= (long)L.ncv * Q::NQP) return;
  const int c = (int)(e / Q::NQP), q = (int)(e - (long)c * Q::NQP);
  const int k = c / L.plane - 1 + L.zlo;
  if (k < 0 || k >= L.NZ) { rlam[e] = 0.0; return; }
  const int q0 = q % Q::NQ, q1 = (q / Q::NQ) % Q::NQY, q2 = q / (Q::NQ * Q::NQY);
  double sq = 0.0;
#pragma unroll
  for (int m = 0; m < Q::NSL; ++m) {
    const int m0 = m & 1, m1 = HY ? (m >> 1) & 1 : 0, m2 = m >> (HY ? 2 : 1);
    sq += S[(long)c * Q::NSL + m] * (s_O[q0][m0] * (HY ? s_O[q1][m1] : 1.0) * s_O[q2][m2]);
  }
  const double w = s_gw[q0] * (HY ? s_gw[q1] : 1.0) * s_gw[q2];
  rlam[e] = w / lambda_of(ph, sq);
}

// transverse directions (e < f) of direction D and the stride of a mass-rule point index along each direction
template <bool HY, int D>
struct QDir {
  static constexpr int E = D == 0 ? 1 : 0, F = D == 2 ? 1 : 2;
  __host__ __device__ static constexpr int nq(int dir) { return dir == 1 ? QD<HY>::NQY : QD<HY>::NQ; }
  __host__ __device__ static constexpr int sq(int dir) { return dir == 0 ? 1 : (dir == 1 ? QD<HY>::NQ : QD<HY>::NQ * QD<HY>::NQY); }
  __host__ __device__ static constexpr int no(int dir) { return (dir == 1 && !HY) ? 1 : 2; }  // open nodes along a transverse direction
};

template <bool HY, int D>
__device__ __forceinline__ Line1 line_of(const Lay1& L, const Cell1& q, int te, int tf) {
  if constexpr (D == 0) return line_x<HY>(L, q, te, tf);
  else if constexpr (D == 1) return line_y(L, q, te, tf);
  else return line_z<HY>(L, q, te, tf);
}

// sum over the cell's mass-rule points of c_q phi_(i,te,tf)(q)^2
template <bool HY, int D>
__device__ __forceinline__ double q_diag(const QTab& tb, const double* __restrict__ rl, int i, int te, int tf) {
  using R = QDir<HY, D>;
  double s = 0.0;
  for (int qf = 0; qf < R::nq(R::F); ++qf)
    for (int qe = 0; qe < R::nq(R::E); ++qe) {
      const double oe = ob<HY>(tb, R::E, qe, te), of = ob<HY>(tb, R::F, qf, tf);
      const double oo = (oe * oe) * (of * of);
      for (int qd = 0; qd < QD<HY>::NQ; ++qd) {
        const double cb = tb.C[qd][i];
        s += rl[qd * R::sq(D) + qe * R::sq(R::E) + qf * R::sq(R::F)] * (cb * cb) * oo;
      }
    }
  return s;
}

// diag(M) and T = 1/diag(M) per velocity node of direction D (1 / 0 on eliminated boundary nodes)
template <bool HY, int D>
__device__ __forceinline__ void q_md_dir(const Lay1& L, const Cell1& q, const QTab& tb, const double* __restrict__ rlam,
                                         const double* __restrict__ K, double geo, double* __restrict__ Md,
                                         double* __restrict__ T) {
  using R = QDir<HY, D>;
  using Q = QD<HY>;
  const double* rl = rlam + (long)q.c * Q::NQP;
  const double sc = geo / K[q.c];
  for (int tf = 0; tf < R::no(R::F); ++tf)
    for (int te = 0; te < R::no(R::E); ++te) {
      const Line1 l = line_of<HY, D>(L, q, te, tf);
      if (l.cd > 0) {
        const int cl = q.c - l.cstride;
        const double md = (geo / K[cl]) * q_diag<HY, D>(tb, rlam + (long)cl * Q::NQP, 2, te, tf) + sc * q_diag<HY, D>(tb, rl, 0, te, tf);
        Md[l.base] = md; T[l.base] = 1.0 / md;
      } else { Md[l.base] = 1.0; T[l.base] = 0.0; }
      const double md1 = sc * q_diag<HY, D>(tb, rl, 1, te, tf);
      Md[l.base + l.sd] = md1; T[l.base + l.sd] = 1.0 / md1;
      if (l.cd + 1 == l.nd) { Md[l.base + 2 * l.sd] = 1.0; T[l.base + 2 * l.sd] = 0.0; }
      else if (D == 2 && q.k == L.zhi - 1) {  // ghost-above node plane of a slab: needed by the Schur kernels
        const int cu = q.c + l.cstride;
        const double md = sc * q_diag<HY, D>(tb, rl, 2, te, tf) + (geo / K[cu]) * q_diag<HY, D>(tb, rlam + (long)cu * Q::NQP, 0, te, tf);
        Md[l.base + 2 * l.sd] = md; T[l.base + 2 * l.sd] = 1.0 / md;
      }
    }
}
template <bool HY>
__global__ void __launch_bounds__(BLK) k1q_md(Lay1 L, QTab tb, const double* __restrict__ rlam,
                                              const double* __restrict__ Kx, const double* __restrict__ Ky,
                                              const double* __restrict__ Kz, double gx, double gy, double gz,
                                              double* __restrict__ Md, double* __restrict__ T) {
  const Cell1 q = owned_cell1(L);
  if (!q.ok) return;
  q_md_dir<HY, 0>(L, q, tb, rlam, Kx, gx, Md, T);
  if constexpr (HY) q_md_dir<HY, 1>(L, q, tb, rlam, Ky, gy, Md, T);
  q_md_dir<HY, 2>(L, q, tb, rlam, Kz, gz, Md, T);
}

// diagonal of S = B diag(M)^-1 B^T per L2 dof from T (p-part of Md); same terms and order as line_coeff
template <bool HY>
__global__ void __launch_bounds__(BLK) k1_schur_diag(Lay1 L, Rt1Const C, const double* __restrict__ T,
                                                     double* __restrict__ Md) {
  const Dof1 d = owned_dof1<HY>(L);
  if (!d.q.ok) return;
  const Cell1& q = d.q;
  const int l = d.l;
  const int a0 = l & 1, a1 = HY ? (l >> 1) & 1 : 0, a2 = l >> (HY ? 2 : 1);
  double D = 0.0;
  {
    const Line1 ln = line_x<HY>(L, q, a1, a2);
    if (ln.cd > 0) D += BV(a0, 0) * (BV(a0, 0) * T[ln.base]);
    D += BV(a0, 1) * (BV(a0, 1) * T[ln.base + ln.sd]);
    if (ln.cd + 1 < ln.nd) D += BV(a0, 2) * (BV(a0, 2) * T[ln.base + 2 * ln.sd]);
  }
  if constexpr (HY) {
    const Line1 ln = line_y(L, q, a0, a2);
    if (ln.cd > 0) D += BV(a1, 0) * (BV(a1, 0) * T[ln.base]);
    D += BV(a1, 1) * (BV(a1, 1) * T[ln.base + ln.sd]);
    if (ln.cd + 1 < ln.nd) D += BV(a1, 2) * (BV(a1, 2) * T[ln.base + 2 * ln.sd]);
  }
  {
    const Line1 ln = line_z<HY>(L, q, a0, a1);
    if (ln.cd > 0) D += BV(a2, 0) * (BV(a2, 0) * T[ln.base]);
    D += BV(a2, 1) * (BV(a2, 1) * T[ln.base + ln.sd]);
    if (ln.cd + 1 < ln.nd) D += BV(a2, 2) * (BV(a2, 2) * T[ln.base + 2 * ln.sd]);
  }
  Md[L.off_p + q.pb + l] = D;
}

// ---------------------------------------------------------------------------------------
// K1 (os = 1), pass A.  One thread per (owned cell, direction): a CTA covers K1Q_TC consecutive owned cells, its
// warps are specialised by direction (x | y | z), so a thread applies ONE dense 12x12 element block
//   y_loc = (geo_d / K_d) Phi^T diag(c_q) Phi x_loc          (c_q = w_q / lambda_q from k1q_rlam)
// in sum-factorised form (per transverse point pair: collapse the open directions, run the 4-point line along d,
// expand back), adds the line-structured B^T p terms and writes
//   interior-node rows: complete;  lower-face-node rows: own element + B^T (the lower cell's element part is added
//   by pass B);  upper-face-node contributions: to yup[cell][dir][line] for the upper neighbour's pass B.
// The c_q of the CTA's cells (one contiguous range of rlam) and their pressure dofs are staged through shared memory
// with coalesced loads (odd row stride: conflict-free reads); the three direction threads of a cell meet in shared
// memory for the pressure rows  yp = (yp_x + yp_y) + yp_z  and the cell's share of <x, y>
//   acc = ((acc_x + acc_y) + acc_z) + sum_l pc_l yp_l,
// which goes to accA[cell]; pass B reduces it together with its own share (fixed trees, no atomics on doubles).
// Eliminated boundary nodes enter with x = 0 and keep identity rows [REF src/pressure_solver.cpp:56-61].
// (Round 1's version ran one thread per cell through all three directions: 128 registers + 676 B of spills,
// scattered 8-byte loads of the 64 c_q -- 9 % of the HBM roofline.)
// ---------------------------------------------------------------------------------------
#ifndef IMPES_K1Q_REGS
#define IMPES_K1Q_REGS 128
#endif
#ifndef IMPES_K1Q_UNROLL
#define IMPES_K1Q_UNROLL 1   // point pairs per trip of elem_apply's quadrature loop
#endif
constexpr int K1Q_TC = 32;  // cells per CTA: one warp per direction.  Register cap 128 (15 warps per SM): at 96 (18 warps)
                            // the compiler reused one register pair for the four c_q loads of a quadrature line and
                            // serialised LDS -> DMUL four times per iteration (ncu: short-scoreboard 1.7 per issue)

template <bool HY>
struct K1QS {
  static constexpr int ND = HY ? 3 : 2;
  static constexpr int RS = QD<HY>::NQP | 1;        // row stride of the staged c_q (odd)
  static constexpr int PS = QD<HY>::NSL + 1;        // row stride of the staged pressure dofs / per-direction partials
  static constexpr int THREADS = K1Q_TC * ND;
  static constexpr int SMEM_DOUBLES = K1Q_TC * RS + K1Q_TC * PS + ND * K1Q_TC * PS;
};

// rl: the cell's c_q in shared memory.  3-D: the 16 transverse point pairs run as a ROLLED loop (a fully unrolled
// body is 1600 instructions per direction, executed once per thread: ncu showed the warps starving on instruction
// fetch, stall_no_instruction 4.6 per issue); 2-D (9 points) stays unrolled.
template <bool HY, int D>
__device__ __forceinline__ void elem_apply(const QTab& tb, const double* rl, const double (&xl)[12], double (&yl)[12]) {
  using R = QDir<HY, D>;
  constexpr int NOE = R::no(R::E), NOF = R::no(R::F);
#pragma unroll
  for (int n = 0; n < 12; ++n) yl[n] = 0.0;
  if constexpr (HY) {
#if IMPES_K1Q_UNROLL > 1
#pragma unroll 1
    for (int pp = 0; pp < 16; pp += IMPES_K1Q_UNROLL)
#pragma unroll
    for (int p = pp; p < pp + IMPES_K1Q_UNROLL; ++p) {
#else
#pragma unroll 1
    for (int p = 0; p < 16; ++p) {
#endif
      const int qe = p & 3, qf = p >> 2;
      const double* oo = c_OO3[qe][qf];
      const double o0 = oo[0], o1 = oo[1], o2 = oo[2], o3 = oo[3];
      double v[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        v[i] = __fma_rn(o3, xl[i + 9], __fma_rn(o2, xl[i + 6], __fma_rn(o1, xl[i + 3], __fma_rn(o0, xl[i], 0.0))));
      double r[3] = {0.0, 0.0, 0.0};
      const double* rq = rl + qe * R::sq(R::E) + qf * R::sq(R::F);
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        const double u = __fma_rn(tb.C[qd][2], v[2], __fma_rn(tb.C[qd][1], v[1], tb.C[qd][0] * v[0]));
        const double cu = rq[qd * R::sq(D)] * u;
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = __fma_rn(tb.C[qd][i], cu, r[i]);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        yl[i] = __fma_rn(o0, r[i], yl[i]);
        yl[i + 3] = __fma_rn(o1, r[i], yl[i + 3]);
        yl[i + 6] = __fma_rn(o2, r[i], yl[i + 6]);
        yl[i + 9] = __fma_rn(o3, r[i], yl[i + 9]);
      }
    }
  } else {
#pragma unroll
    for (int qf = 0; qf < R::nq(R::F); ++qf)
#pragma unroll
      for (int qe = 0; qe < R::nq(R::E); ++qe) {
        double oo[NOE * NOF], v[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
          for (int te = 0; te < NOE; ++te) {
            const double o2 = ob<HY>(tb, R::E, qe, te) * ob<HY>(tb, R::F, qf, tf);
            oo[te + NOE * tf] = o2;
#pragma unroll
            for (int i = 0; i < 3; ++i) v[i] = __fma_rn(o2, xl[i + 3 * (te + NOE * tf)], v[i]);
          }
        double r[3] = {0.0, 0.0, 0.0};
        const double* rq = rl + qe * R::sq(R::E) + qf * R::sq(R::F);
#pragma unroll
        for (int qd = 0; qd < QD<HY>::NQ; ++qd) {
          const double u = __fma_rn(tb.C[qd][2], v[2], __fma_rn(tb.C[qd][1], v[1], tb.C[qd][0] * v[0]));
          const double cu = rq[qd * R::sq(D)] * u;
#pragma unroll
          for (int i = 0; i < 3; ++i) r[i] = __fma_rn(tb.C[qd][i], cu, r[i]);
        }
#pragma unroll
        for (int t = 0; t < NOE * NOF; ++t)
#pragma unroll
          for (int i = 0; i < 3; ++i) yl[i + 3 * t] = __fma_rn(oo[t], r[i], yl[i + 3 * t]);
      }
  }
}

// CTA-wide barrier that warps reach from different code paths (the three direction bodies below): a named
// barrier is identified by its id, not by the program counter (PTX barrier.sync id, count).
template <int NTHREADS>
__device__ __forceinline__ void named_barrier_1() {
  asm volatile("barrier.sync 1, %0;" ::"n"(NTHREADS) : "memory");
}

// one (cell, direction) of pass A.  rl / pc: the cell's c_q and pressure dofs in shared memory (valid after the
// barrier); ypd: this direction's partial pressure rows [NSL] followed by its share of the dot.  ok = false (ragged
// last CTA): the lane shadows another cell and stores nothing.
template <bool HY, int D, bool DOT>
__device__ __forceinline__ void q_apply_dir(const Lay1& L, const Cell1& q, bool ok, const Rt1Const& C, const QTab& tb,
                                            const double* rl, const double* __restrict__ K, double geo,
                                            const double* __restrict__ x, double mu, double* __restrict__ y,
                                            double* __restrict__ yup, const double* pc, double* ypd) {
  using R = QDir<HY, D>;
  using Q = QD<HY>;
  constexpr int NOE = R::no(R::E), NOF = R::no(R::F), NT = NOE * NOF;
  constexpr int DI = HY ? D : (D == 0 ? 0 : 1);  // index of the direction among the present ones
  constexpr int PS1 = (D == 0) ? 1 : (D == 1 ? 2 : (HY ? 4 : 2));  // stride between the two pressure dofs of a line
  // mu: lazy MINRES scaling of the input vector (MinresState::mu; 1 outside MINRES): every entry is multiplied once,
  // right after its load, so all arithmetic below sees the doubles an explicitly scaled vector would hold
  double xl[12], ttlo[NT];
  int cd = 0, nd = 1;
  // ---- every global load of the thread, then the barrier that publishes the staged data
  const double Kc = K[q.c];
#pragma unroll
  for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
    for (int te = 0; te < NOE; ++te) {
      const Line1 l = line_of<HY, D>(L, q, te, tf);
      const int t = te + NOE * tf;
      cd = l.cd; nd = l.nd;
      xl[3 * t] = x[l.base] * mu;
      xl[1 + 3 * t] = x[l.base + l.sd] * mu;
      xl[2 + 3 * t] = x[l.base + 2 * l.sd] * mu;
    }
  const bool lo = cd > 0, hi = cd + 1 < nd;
#pragma unroll
  for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
    for (int te = 0; te < NOE; ++te) {  // lower neighbour's pressure dofs (clamped to the own cell where there is none)
      const Line1 l = line_of<HY, D>(L, q, te, tf);
      const double* pL = x + (q.pb - (lo ? l.pdn : 0) + l.lp);
      double tt = 0.0;
      tt += C.Bv[0][2] * (pL[0] * mu); tt += C.Bv[1][2] * (pL[l.ps] * mu);
      ttlo[te + NOE * tf] = tt;
    }
#pragma unroll
  for (int n = 3 * NT; n < 12; ++n) xl[n] = 0.0;
  cp_async_wait<0>();
  named_barrier_1<K1QS<HY>::THREADS>();
  const double sc = geo / Kc;
#pragma unroll
  for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
    for (int te = 0; te < NOE; ++te) {
      const int t = te + NOE * tf;
      const int lp = (D == 0) ? 2 * te + 2 * (HY ? 2 : 1) * tf : (D == 1 ? te + 4 * tf : te + 2 * tf);
      // B^T terms of the lower-face rows, completed before the element product
      ttlo[t] += C.Bv[0][0] * (pc[lp] * mu); ttlo[t] += C.Bv[1][0] * (pc[lp + PS1] * mu);
      // eliminated boundary nodes enter the element product (and the pressure rows) with x = 0
      xl[3 * t] = lo ? xl[3 * t] : 0.0;
      xl[2 + 3 * t] = hi ? xl[2 + 3 * t] : 0.0;
    }
  double yl[12];
  elem_apply<HY, D>(tb, rl, xl, yl);
  double acc = 0.0;
#pragma unroll
  for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
    for (int te = 0; te < NOE; ++te) {
      const Line1 l = line_of<HY, D>(L, q, te, tf);
      const int t = te + NOE * tf;
      const int lp = (D == 0) ? 2 * te + 2 * (HY ? 2 : 1) * tf : (D == 1 ? te + 4 * tf : te + 2 * tf);
      const double pc0 = pc[lp] * mu, pc1 = pc[lp + PS1] * mu;
      const double x0 = xl[3 * t], x1 = xl[1 + 3 * t], x2 = xl[2 + 3 * t];
      {  // interior node
        double tt = 0.0;
        tt += C.Bv[0][1] * pc0; tt += C.Bv[1][1] * pc1;
        const double y1 = sc * yl[1 + 3 * t] + tt;
        if (ok) y[l.base + l.sd] = y1;
        if (DOT) acc += x1 * y1;
      }
      if (lo) {  // lower face node: own element part + B^T; completed by pass B
        if (ok) y[l.base] = sc * yl[3 * t] + ttlo[t];
      } else {   // eliminated boundary node: identity row (its raw value is re-read: boundary cells only)
        const double xr = x[l.base] * mu;
        if (ok) y[l.base] = xr;
        if (DOT) acc += xr * xr;
      }
      if (hi) {
        if (ok) yup[(long)q.c * Q::NUP + DI * Q::NL + t] = sc * yl[2 + 3 * t];
      } else {
        const double xr = x[l.base + 2 * l.sd] * mu;
        if (ok) y[l.base + 2 * l.sd] = xr;
        if (DOT) acc += xr * xr;
      }
      // this direction's part of the two pressure rows of the line
      double a0 = 0.0, a1 = 0.0;
      a0 += C.Bv[0][0] * x0; a0 += C.Bv[0][1] * x1; a0 += C.Bv[0][2] * x2;
      a1 += C.Bv[1][0] * x0; a1 += C.Bv[1][1] * x1; a1 += C.Bv[1][2] * x2;
      ypd[lp] = a0; ypd[lp + PS1] = a1;
    }
  ypd[Q::NSL] = acc;
}

template <bool HY, bool DOT>
__global__ void __launch_bounds__(K1QS<HY>::THREADS) __maxnreg__(HY ? IMPES_K1Q_REGS : 80)
k1q_apply_a(Lay1 L, Rt1Const C, QTab tb, const double* __restrict__ rlam, const double* __restrict__ Kx,
            const double* __restrict__ Ky, const double* __restrict__ Kz, double gx, double gy, double gz,
            const double* __restrict__ x, double* __restrict__ y, double* __restrict__ yup,
            double* __restrict__ accA, const MinresState* st) {
  using Q = QD<HY>;
  using S = K1QS<HY>;
  if (st && st->done) return;
  extern __shared__ double k1q_sm[];
  double* s_rl = k1q_sm;                       // [TC][RS]
  double* s_pc = s_rl + K1Q_TC * S::RS;        // [TC][PS]
  double* s_yp = s_pc + K1Q_TC * S::PS;        // [ND][TC][PS]
  const int tid = threadIdx.x;
  const int ci = tid % K1Q_TC, dir = tid / K1Q_TC;   // dir is warp-uniform (K1Q_TC is a multiple of 32)
  const unsigned t0 = blockIdx.x * (unsigned)K1Q_TC;  // first owned cell of the CTA
  const unsigned nc = min((unsigned)K1Q_TC, (unsigned)L.ncl - t0);
  {  // stage c_q and the pressure dofs of the CTA's cells: both are contiguous ranges (cell-major arrays).
     // cp.async: every copy of the thread is in flight at once and lands in shared memory without a register
     // round trip (the LDG -> STS loop it replaces paid one L2 latency per iteration: ncu long-scoreboard 4.5/issue)
    const double* g_rl = rlam + ((long)t0 + L.plane) * Q::NQP;
    constexpr int NRL = (K1Q_TC * Q::NQP + S::THREADS - 1) / S::THREADS;
#pragma unroll
    for (int it = 0; it < NRL; ++it) {
      const unsigned e = tid + it * S::THREADS;
      if (e < nc * Q::NQP) {
        const unsigned c = e / Q::NQP, qq = e - c * Q::NQP;
        cp_async8(s_rl + c * S::RS + qq, g_rl + e);
      }
    }
    const double* g_pc = x + L.off_p + ((long)t0 + L.plane) * Q::NSL;
    constexpr int NPC = (K1Q_TC * Q::NSL + S::THREADS - 1) / S::THREADS;
#pragma unroll
    for (int it = 0; it < NPC; ++it) {
      const unsigned e = tid + it * S::THREADS;
      if (e < nc * Q::NSL) {
        const unsigned c = e / Q::NSL, l = e - c * Q::NSL;
        cp_async8(s_pc + c * S::PS + l, g_pc + e);
      }
    }
    cp_async_commit();
  }
  Cell1 q;
  {
    const unsigned t = t0 + ci;
    q.ok = (unsigned)ci < nc;
    const unsigned tt = q.ok ? t : t0;   // idle lanes (ragged last CTA) shadow the CTA's first cell and store nothing
    const unsigned r = fastdiv(tt, L.fdx);
    q.i = (int)(tt - r * (unsigned)L.NX);
    q.kl = (int)fastdiv(r, L.fdy);
    q.j = (int)(r - (unsigned)q.kl * (unsigned)L.NY);
    q.k = q.kl + L.zlo;
    q.c = q.i + L.NX * (q.j + L.NY * (q.kl + 1));
    q.pb = L.off_p + q.c * L.npl;
  }
  const int cs = q.ok ? ci : 0;
  const double* rl = s_rl + cs * S::RS;
  const double* pc = s_pc + cs * S::PS;
  double* ypd = s_yp + (dir * K1Q_TC + ci) * S::PS;
  // dir is warp-uniform: no divergence between the three bodies; each contains the named barrier once
  const double mu = st ? st->mu : 1.0;
  if (dir == 0) q_apply_dir<HY, 0, DOT>(L, q, q.ok, C, tb, rl, Kx, gx, x, mu, y, yup, pc, ypd);
  else if (HY && dir == 1) q_apply_dir<HY, 1, DOT>(L, q, q.ok, C, tb, rl, Ky, gy, x, mu, y, yup, pc, ypd);
  else q_apply_dir<HY, 2, DOT>(L, q, q.ok, C, tb, rl, Kz, gz, x, mu, y, yup, pc, ypd);
  __syncthreads();
  if (dir == 0 && q.ok) {  // pressure rows and the cell's share of <x, y>
    const double* a = s_yp + ci * S::PS;
    const double* b = s_yp + (K1Q_TC + ci) * S::PS;
    const double* c2 = s_yp + (2 * K1Q_TC + ci) * S::PS;
    double acc = HY ? (a[Q::NSL] + b[Q::NSL]) + c2[Q::NSL] : a[Q::NSL] + b[Q::NSL];
    double ap = 0.0;
#pragma unroll
    for (int l = 0; l < Q::NSL; ++l) {
      const double yp = HY ? (a[l] + b[l]) + c2[l] : a[l] + b[l];
      y[q.pb + l] = yp;
      if (DOT) ap += (pc[l] * mu) * yp;
    }
    if (DOT) accA[t0 + ci] = acc + ap;
  }
}

// pass B: lower-face-node rows += the lower cell's element contribution; <x, y> = tree(accA) + tree(own share)
template <bool HY, int D, bool DOT>
__device__ __forceinline__ void q_face_dir(const Lay1& L, const Cell1& q, const double* __restrict__ x, double mu,
                                           double* __restrict__ y, const double* __restrict__ yup, double& acc) {
  using R = QDir<HY, D>;
  using Q = QD<HY>;
  constexpr int NOE = R::no(R::E), NOF = R::no(R::F);
  constexpr int DI = HY ? D : (D == 0 ? 0 : 1);
#pragma unroll
  for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
    for (int te = 0; te < NOE; ++te) {
      const Line1 l = line_of<HY, D>(L, q, te, tf);
      if (l.cd > 0) {
        const double yv = y[l.base] + yup[(long)(q.c - l.cstride) * Q::NUP + DI * Q::NL + te + NOE * tf];
        y[l.base] = yv;
        if (DOT) acc += (x[l.base] * mu) * yv;
      }
    }
}
// DOT: the nblk 256-cell blocks are walked by a resident grid (block b -> partial[b], partial[nblk + b]) so that the
// fence + ticket of the final reduction is paid once per CTA (ncu, one CTA per block: barrier 15.7 stalls per issue).
template <bool HY, bool DOT>
__global__ void __launch_bounds__(BLK) k1q_apply_b(Lay1 L, const double* __restrict__ x, double* __restrict__ y,
                                                   const double* __restrict__ yup, const double* __restrict__ accA,
                                                   double* partial, double* out, unsigned* ticket,
                                                   const MinresState* st, unsigned nblk) {
  if (st && st->done) return;
  __shared__ double sm[BLK / 32];
  for (unsigned blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const Cell1 q = owned_cell1_at(L, blk);
    double acc = 0.0, accq = 0.0;
    if (q.ok) {
      const double mu = st ? st->mu : 1.0;
      q_face_dir<HY, 0, DOT>(L, q, x, mu, y, yup, acc);
      if constexpr (HY) q_face_dir<HY, 1, DOT>(L, q, x, mu, y, yup, acc);
      q_face_dir<HY, 2, DOT>(L, q, x, mu, y, yup, acc);
      if (DOT) accq = accA[blk * (unsigned)BLK + threadIdx.x];
    }
    if (DOT) {
      const double sa = block_tree_256(accq, sm);
      __syncthreads();
      const double sb = block_tree_256(acc, sm);
      if (threadIdx.x == 0) {
        partial[blk] = sa;
        partial[nblk + blk] = sb;
      }
      __syncthreads();   // sm is reused by the next block
    }
  }
  if (DOT) {
    __shared__ int is_last;
    if (threadIdx.x == 0) {
      __threadfence();
      is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      const double a = ordered_partial_sum(partial, nblk), b = ordered_partial_sum(partial + nblk, nblk);
      __syncthreads();
      const double ta = block_tree_256(a, sm);
      __syncthreads();
      const double tb2 = block_tree_256(b, sm);
      if (threadIdx.x == 0) {
        *out = ta + tb2;
        *ticket = 0u;
      }
    }
  }
}

// ---- the same diag(M) / T, two passes in the shape of K1 (k1q_apply_a / _b below).  k1q_md above walks 36 + 12 node
// sums of 64 terms per thread straight from global memory with a 512-byte stride between lanes: 21.6 ms at 192^3,
// ten K1 applications.  Here pass A runs one thread per (owned cell, direction), the c_q of the CTA's cells staged by
// coalesced cp.async, ONE walk over the 64 points feeding the 12 node sums of the direction (each sum adds the same
// terms  (c_q (C C)) ((Oe Oe)(Of Of))  in the same order qf, qe, qd as q_diag: same bits); it completes interior and
// boundary nodes, leaves its own part on lower-face nodes and hands the upper-face part to the upper neighbour through
// up[cell][dir][line] (the K1 scratch).  Pass B: md = lower cell's part + own part (q_md_dir's order), T = 1 / md.
// Lower faces of a slab's bottom plane and the ghost node plane above its top plane involve ghost cells: k1q_md_slab
// (the direct form, two cell planes only).
template <bool HY, int D>
__device__ __forceinline__ void q_md_a_dir(const Lay1& L, const Cell1& q, const QTab& tb, const double* rl,
                                           const double* __restrict__ K, double geo, double* __restrict__ Md,
                                           double* __restrict__ T, double* __restrict__ up) {
  using R = QDir<HY, D>;
  using Q = QD<HY>;
  constexpr int NOE = R::no(R::E), NOF = R::no(R::F), NT = NOE * NOF;
  constexpr int DI = HY ? D : (D == 0 ? 0 : 1);
  double s[3][NT];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int t = 0; t < NT; ++t) s[i][t] = 0.0;
  auto pair = [&](const double* rq, const double (&oo)[NT]) {
#pragma unroll
    for (int qd = 0; qd < Q::NQ; ++qd) {
      const double c = rq[qd * R::sq(D)];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double ci = c * (tb.C[qd][i] * tb.C[qd][i]);
#pragma unroll
        for (int t = 0; t < NT; ++t) s[i][t] += ci * oo[t];
      }
    }
  };
  if constexpr (HY) {
#pragma unroll 1
    for (int p = 0; p < 16; ++p) {   // qf outer, qe inner
      const int qe = p & 3, qf = p >> 2;
      const double* o = c_OO3SQ[qe][qf];
      const double oo[NT] = {o[0], o[1], o[2], o[3]};
      pair(rl + qe * R::sq(R::E) + qf * R::sq(R::F), oo);
    }
  } else {
#pragma unroll
    for (int qf = 0; qf < R::nq(R::F); ++qf)
#pragma unroll
      for (int qe = 0; qe < R::nq(R::E); ++qe) {
        double oo[NT];
#pragma unroll
        for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
          for (int te = 0; te < NOE; ++te) {
            const double oe = ob<HY>(tb, R::E, qe, te), of = ob<HY>(tb, R::F, qf, tf);
            oo[te + NOE * tf] = (oe * oe) * (of * of);
          }
        pair(rl + qe * R::sq(R::E) + qf * R::sq(R::F), oo);
      }
  }
  if (!q.ok) return;
  const double sc = geo / K[q.c];
#pragma unroll
  for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
    for (int te = 0; te < NOE; ++te) {
      const Line1 l = line_of<HY, D>(L, q, te, tf);
      const int t = te + NOE * tf;
      if (l.cd > 0) Md[l.base] = sc * s[0][t];   // own part; pass B / k1q_md_slab add the lower cell's
      else { Md[l.base] = 1.0; T[l.base] = 0.0; }
      const double md1 = sc * s[1][t];
      Md[l.base + l.sd] = md1; T[l.base + l.sd] = 1.0 / md1;
      if (l.cd + 1 == l.nd) { Md[l.base + 2 * l.sd] = 1.0; T[l.base + 2 * l.sd] = 0.0; }
      else up[(long)q.c * Q::NUP + DI * Q::NL + t] = sc * s[2][t];
    }
}
template <bool HY>
__global__ void __launch_bounds__(K1QS<HY>::THREADS) __maxnreg__(HY ? 96 : 80)
k1q_md_a(Lay1 L, QTab tb, const double* __restrict__ rlam, const double* __restrict__ Kx, const double* __restrict__ Ky,
         const double* __restrict__ Kz, double gx, double gy, double gz, double* __restrict__ Md, double* __restrict__ T,
         double* __restrict__ up) {
  using Q = QD<HY>;
  using S = K1QS<HY>;
  extern __shared__ double k1q_sm[];
  double* s_rl = k1q_sm;   // [TC][RS]
  const int tid = threadIdx.x;
  const int ci = tid % K1Q_TC, dir = tid / K1Q_TC;
  const unsigned t0 = blockIdx.x * (unsigned)K1Q_TC;
  const unsigned nc = min((unsigned)K1Q_TC, (unsigned)L.ncl - t0);
  const double* g_rl = rlam + ((long)t0 + L.plane) * Q::NQP;
  constexpr int NRL = (K1Q_TC * Q::NQP + S::THREADS - 1) / S::THREADS;
#pragma unroll
  for (int it = 0; it < NRL; ++it) {
    const unsigned e = tid + it * S::THREADS;
    if (e < nc * Q::NQP) {
      const unsigned c = e / Q::NQP, qq = e - c * Q::NQP;
      cp_async8(s_rl + c * S::RS + qq, g_rl + e);
    }
  }
  cp_async_commit();
  Cell1 q;
  {
    q.ok = (unsigned)ci < nc;
    const unsigned tt = q.ok ? t0 + ci : t0;
    const unsigned r = fastdiv(tt, L.fdx);
    q.i = (int)(tt - r * (unsigned)L.NX);
    q.kl = (int)fastdiv(r, L.fdy);
    q.j = (int)(r - (unsigned)q.kl * (unsigned)L.NY);
    q.k = q.kl + L.zlo;
    q.c = q.i + L.NX * (q.j + L.NY * (q.kl + 1));
    q.pb = L.off_p + q.c * L.npl;
  }
  const double* rl = s_rl + (q.ok ? ci : 0) * S::RS;
  cp_async_wait<0>();
  __syncthreads();
  if (dir == 0) q_md_a_dir<HY, 0>(L, q, tb, rl, Kx, gx, Md, T, up);
  else if (HY && dir == 1) { if constexpr (HY) q_md_a_dir<HY, 1>(L, q, tb, rl, Ky, gy, Md, T, up); }
  else q_md_a_dir<HY, 2>(L, q, tb, rl, Kz, gz, Md, T, up);
}
// pass B: lower-face nodes whose lower cell belongs to the slab
template <bool HY, int D>
__device__ __forceinline__ void q_md_b_dir(const Lay1& L, const Cell1& q, double* __restrict__ Md, double* __restrict__ T,
                                           const double* __restrict__ up) {
  using R = QDir<HY, D>;
  using Q = QD<HY>;
  constexpr int NOE = R::no(R::E), NOF = R::no(R::F);
  constexpr int DI = HY ? D : (D == 0 ? 0 : 1);
  if (D == 2 && q.kl == 0) return;   // the lower cell is a ghost (or absent): k1q_md_slab
#pragma unroll
  for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
    for (int te = 0; te < NOE; ++te) {
      const Line1 l = line_of<HY, D>(L, q, te, tf);
      if (l.cd > 0) {
        const double md = up[(long)(q.c - l.cstride) * Q::NUP + DI * Q::NL + te + NOE * tf] + Md[l.base];
        Md[l.base] = md; T[l.base] = 1.0 / md;
      }
    }
}
template <bool HY>
__global__ void __launch_bounds__(BLK) k1q_md_b(Lay1 L, double* __restrict__ Md, double* __restrict__ T,
                                                const double* __restrict__ up) {
  const Cell1 q = owned_cell1(L);
  if (!q.ok) return;
  q_md_b_dir<HY, 0>(L, q, Md, T, up);
  if constexpr (HY) q_md_b_dir<HY, 1>(L, q, Md, T, up);
  q_md_b_dir<HY, 2>(L, q, Md, T, up);
}
// z-face nodes shared with the neighbouring slabs (multi-GPU only): block 0.. of the grid = bottom owned plane (lower
// face nodes), the second half = top owned plane (the ghost node plane above).  Direct form of q_md_dir.
template <bool HY>
__global__ void __launch_bounds__(BLK) k1q_md_slab(Lay1 L, QTab tb, const double* __restrict__ rlam,
                                                   const double* __restrict__ Kz, double gz, double* __restrict__ Md,
                                                   double* __restrict__ T) {
  using R = QDir<HY, 2>;
  using Q = QD<HY>;
  const int t = blockIdx.x * BLK + threadIdx.x;
  if (t >= 2 * L.plane) return;
  const int side = t / L.plane, ip = t - side * L.plane;
  Cell1 q;
  q.i = ip % L.NX; q.j = ip / L.NX; q.kl = side ? L.nzl - 1 : 0; q.k = q.kl + L.zlo;
  q.c = q.i + L.NX * (q.j + L.NY * (q.kl + 1));
  q.pb = L.off_p + q.c * L.npl; q.ok = true;
  if (side == 0 && (q.k == 0 || (L.nzl == 1 && false))) return;
  if (side == 1 && q.k + 1 >= L.NZ) return;
  const double* rl = rlam + (long)q.c * Q::NQP;
  const double sc = gz / Kz[q.c];
  for (int tf = 0; tf < R::no(R::F); ++tf)
    for (int te = 0; te < R::no(R::E); ++te) {
      const Line1 l = line_of<HY, 2>(L, q, te, tf);
      if (side == 0) {
        const int cl = q.c - l.cstride;
        const double md = (gz / Kz[cl]) * q_diag<HY, 2>(tb, rlam + (long)cl * Q::NQP, 2, te, tf) + sc * q_diag<HY, 2>(tb, rl, 0, te, tf);
        Md[l.base] = md; T[l.base] = 1.0 / md;
      } else {
        const int cu = q.c + l.cstride;
        const double md = sc * q_diag<HY, 2>(tb, rl, 2, te, tf) + (gz / Kz[cu]) * q_diag<HY, 2>(tb, rlam + (long)cu * Q::NQP, 0, te, tf);
        Md[l.base + 2 * l.sd] = md; T[l.base + 2 * l.sd] = 1.0 / md;
      }
    }
}

// ---------------------------------------------------------------------------------------
// K6 (os = 1): one explicit stage of the DG-Q1 saturation equation, fused:
//   rate = M_phi^-1 (Kt f(S) + b),   Kt = Convection(v,1)^T - DGTrace(v,1,1/2)   (see the file header)
//   RATE ? Sout = rate : Sout = Sin + dt rate,  Fout = f(Sout)            [REF src/saturation_solver.cpp:74-92]
// One thread per owned cell.  Volume term with the 2-point Gauss rule collocated at the DG nodes (exact for this
// integrand; MFEM's rule has more points, same value); face terms with the DG trace integrator's own rule
// because (v.n)+- is only piecewise polynomial; upwinding per face point with a = un/2, b = |un|/2 like
// DGTraceIntegrator [MFEM-recall].  M_phi and the Gauss-node Lagrange basis are diagonal.
// F = f(S) (or S in single-phase mode) travels with S, ghost planes included.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double sat_source_at(const SatParam& sp, double px, double py, double pz) {
  // SaturationSourceCoefficient::eval2D/eval3D [REF src/param.hpp:382-406]; internal coordinates (2-D: y -> z)
  if (sp.well_2d) return (fabs(px - sp.inj_c0) <= sp.inj_r && fabs(pz - sp.inj_c1) <= sp.inj_r) ? sp.src : 0.;
  return (fabs(px - sp.inj_c0) <= sp.inj_r && fabs(pz - sp.inj_c2) <= sp.inj_r && py >= sp.inj_c1 &&
          py <= sp.inj_c1 + sp.inj_h) ? sp.src : 0.;
}

// The velocity is frozen during the sub-steps of a step [REF src/saturation_solver.cpp:74-92], so everything the
// stage kernel needs from it is formed ONCE per step by k1q_sat_pre, in dof-major ("SoA") arrays over the cvec cell
// index c (ghost planes included), so that the 32 cells of a warp read 32 consecutive doubles:
//   vh[(dir NVH + 2 t + n) ncv + c]   v.e_dir at DG node n of line t (volume term)
//   fn[(dir NFP + p) ncv + c]         w_p (v.n)(x_p) at the face-rule points of the cell's LOWER face in direction dir
// (the upper face of a cell is the lower face of its upper neighbour -- the same RT1 face dofs, hence the same bits;
// the top owned plane of a slab also fills the z-faces of the ghost plane above).  f(S) travels in the same layout:
//   F[m ncv + c],  m = local DG dof.
// The stage kernel then costs ~360 instead of ~520 fp64 instructions per (cell, direction) and every global access
// of a warp is one contiguous 256-byte run (ncu on the cell-major form: l1tex data-pipe wavefronts 66-74 %, the top
// unit, from the 64-byte-strided gathers of the neighbours' F and the stride-2 walks over the RT1 node grids).
// Upwinding: DGTraceIntegrator's (a + b, b - a) with a = un/2, b = |un|/2 are exactly (un+, un-), so
//   w (a + b) g1 - w (b - a) g2  ==  c * (c >= 0 ? g1 : g2),   c = w un     (same bits for finite g)
#ifndef IMPES_K1QS_UNROLL3
#define IMPES_K1QS_UNROLL3 1   // 1: face loop as 3 rolled rows of 3 unrolled points (-2 % in the step); 0: 9 rolled points; 2: fully unrolled (+6 %)
#endif
template <bool HY>
struct SatQ {
  static constexpr int ND = HY ? 3 : 2;
  static constexpr int NT = HY ? 4 : 2;    // lines per (cell, direction)
  static constexpr int NFP = HY ? 9 : 2;   // face-rule points per face
  static constexpr int NVH = 2 * NT;       // vhat values per (cell, direction)
  static constexpr int THREADS = K1Q_TC * ND;
};

template <bool HY, int D>
__device__ __forceinline__ void q_sat_pre_dir(const Lay1& L, const Cell1& q, const QTab& tb, const double* __restrict__ x,
                                              long ncv, double* __restrict__ vh, double* __restrict__ fn) {
  using R = QDir<HY, D>;
  using SQ = SatQ<HY>;
  constexpr int NOE = R::no(R::E), NOF = R::no(R::F), NT = NOE * NOF;
  constexpr int NFE = (R::E == 1 && !HY) ? 1 : QD<HY>::NF, NFF = (R::F == 1 && !HY) ? 1 : QD<HY>::NF;
  constexpr int DI = HY ? D : (D == 0 ? 0 : 1);
  static_assert(NT == SQ::NT && NFE * NFF == SQ::NFP, "record sizes");
  double ulo[NT], uhi[NT];
  double* vo = vh + (long)DI * SQ::NVH * ncv + q.c;
#pragma unroll
  for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
    for (int te = 0; te < NOE; ++te) {
      const Line1 l = line_of<HY, D>(L, q, te, tf);
      const int t = te + NOE * tf;
      const double u0 = x[l.base], u1 = x[l.base + l.sd], u2 = x[l.base + 2 * l.sd];
      ulo[t] = u0; uhi[t] = u2;
      vo[(2 * t) * ncv] = tb.LC[0][0] * u0 + tb.LC[0][1] * u1 + tb.LC[0][2] * u2;
      vo[(2 * t + 1) * ncv] = tb.LC[1][0] * u0 + tb.LC[1][1] * u1 + tb.LC[1][2] * u2;
    }
  const bool top = D == 2 && q.kl == L.nzl - 1;   // slab top: the ghost plane's lower z-faces are this plane's upper ones
  double* fo = fn + (long)DI * SQ::NFP * ncv + q.c;
#pragma unroll
  for (int qf = 0; qf < NFF; ++qf)
#pragma unroll
    for (int qe = 0; qe < NFE; ++qe) {
      double unl = 0.0, unh = 0.0;
#pragma unroll
      for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
        for (int te = 0; te < NOE; ++te) {
          const int t = te + NOE * tf;
          // the same IEEE products as c_OOF3 (3-D) / the unrolled 2-D form of the stage kernel
          const double oo = HY ? tb.OOF[qe][qf][t] : ((NFE == 1) ? 1.0 : tb.OF[qe][te]) * ((NFF == 1) ? 1.0 : tb.OF[qf][tf]);
          unl = __fma_rn(ulo[t], oo, unl); unh = __fma_rn(uhi[t], oo, unh);
        }
      const double w = ((NFE == 1) ? 1.0 : tb.fw[qe]) * ((NFF == 1) ? 1.0 : tb.fw[qf]);
      fo[(qe + NFE * qf) * ncv] = w * unl;
      if (top) fo[(qe + NFE * qf) * ncv + L.plane] = w * unh;
    }
}
template <bool HY>
__global__ void __launch_bounds__(SatQ<HY>::THREADS) k1q_sat_pre(Lay1 L, QTab tb, const double* __restrict__ x, long ncv,
                                                                 double* __restrict__ vh, double* __restrict__ fn) {
  const int ci = threadIdx.x % K1Q_TC, dir = threadIdx.x / K1Q_TC;
  const unsigned t = blockIdx.x * (unsigned)K1Q_TC + ci;
  if (t >= (unsigned)L.ncl) return;
  Cell1 q;
  const unsigned r = fastdiv(t, L.fdx);
  q.i = (int)(t - r * (unsigned)L.NX);
  q.kl = (int)fastdiv(r, L.fdy);
  q.j = (int)(r - (unsigned)q.kl * (unsigned)L.NY);
  q.k = q.kl + L.zlo;
  q.c = q.i + L.NX * (q.j + L.NY * (q.kl + 1));
  q.pb = L.off_p + q.c * L.npl;
  q.ok = true;
  if (dir == 0) q_sat_pre_dir<HY, 0>(L, q, tb, x, ncv, vh, fn);
  else if (HY && dir == 1) { if constexpr (HY) q_sat_pre_dir<HY, 1>(L, q, tb, x, ncv, vh, fn); }
  else q_sat_pre_dir<HY, 2>(L, q, tb, x, ncv, vh, fn);
}

// One (cell, direction) of a stage: rd[m] = this direction's part of (Kt F)_m for the cell's NSL dofs.
// F, vh, fn: the dof-major arrays described above.  ok = false: the lane shadows another cell (ragged last CTA).
template <bool HY, int D>
__device__ __forceinline__ void q_sat_dir(const Lay1& L, const Cell1& q, const QTab& tb, long ncv,
                                          const double* __restrict__ F, const double* __restrict__ vh,
                                          const double* __restrict__ fn, double* s_fn, double* rd) {
  using R = QDir<HY, D>;
  using Q = QD<HY>;
  using SQ = SatQ<HY>;
  constexpr int NOE = R::no(R::E), NOF = R::no(R::F), NT = NOE * NOF;
  constexpr int NFE = (R::E == 1 && !HY) ? 1 : Q::NF, NFF = (R::F == 1 && !HY) ? 1 : Q::NF;
  constexpr int PS1 = (D == 0) ? 1 : (D == 1 ? 2 : (HY ? 4 : 2));
  constexpr int DI = HY ? D : (D == 0 ? 0 : 1);
  // position of the cell along the direction and the cell stride to its neighbours
  const int cd = D == 0 ? q.i : (D == 1 ? q.j : q.k), nd = D == 0 ? L.NX : (D == 1 ? L.NY : L.NZ);
  const long cst = D == 0 ? 1 : (D == 1 ? L.NX : L.plane);
  const bool lo = cd > 0, hi = cd + 1 < nd;
  // ---- every global load of the prologue first (each one a contiguous run over the warp's cells)
  const double* Fo = F + q.c;
  const double* Fl = Fo - (lo ? cst : 0);   // clamped to the own cell where there is no neighbour (result unused)
  const double* Fh = Fo + (hi ? cst : 0);
  const double* vo = vh + (long)DI * SQ::NVH * ncv + q.c;
  const double* fl_p = fn + (long)DI * SQ::NFP * ncv + q.c;
  const double* fh_p = fl_p + cst;          // beyond the top boundary: a valid address whose value is not used
  if constexpr (HY) {
#pragma unroll
    for (int p = 0; p < 9; ++p) {
      cp_async8(s_fn + p * K1Q_TC, fl_p + p * ncv);
      cp_async8(s_fn + (9 + p) * K1Q_TC, fh_p + p * ncv);
    }
    cp_async_commit();
  }
  double g0[NT], g1[NT], gnl0[NT], gnl1[NT], gnh0[NT], gnh1[NT], v0[NT], v1[NT];
#pragma unroll
  for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
    for (int te = 0; te < NOE; ++te) {
      const int t = te + NOE * tf;
      const int lp = (D == 0) ? 2 * te + 2 * (HY ? 2 : 1) * tf : (D == 1 ? te + 4 * tf : te + 2 * tf);
      g0[t] = Fo[lp * ncv]; g1[t] = Fo[(lp + PS1) * ncv];
      gnl0[t] = Fl[lp * ncv]; gnl1[t] = Fl[(lp + PS1) * ncv];
      gnh0[t] = Fh[lp * ncv]; gnh1[t] = Fh[(lp + PS1) * ncv];
      v0[t] = vo[(2 * t) * ncv]; v1[t] = vo[(2 * t + 1) * ncv];
    }
  double go0[NT], go1[NT], gn_lo[NT], gn_hi[NT];
#pragma unroll
  for (int m = 0; m < Q::NSL; ++m) rd[m] = 0.0;
#pragma unroll
  for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
    for (int te = 0; te < NOE; ++te) {
      const int t = te + NOE * tf;
      const int lp = (D == 0) ? 2 * te + 2 * (HY ? 2 : 1) * tf : (D == 1 ? te + 4 * tf : te + 2 * tf);
      // volume term: sum over the two DG nodes of the line of  w g vhat dpsi_i/dxi_d
      const double vol = tb.wv * (g0[t] * v0[t] + g1[t] * v1[t]);
      rd[lp] += tb.dl[0] * vol;
      rd[lp + PS1] += tb.dl[1] * vol;
      // traces of g on the two faces: own side, and the neighbour's side
      go0[t] = tb.E0[0] * g0[t] + tb.E0[1] * g1[t];
      go1[t] = tb.E1[0] * g0[t] + tb.E1[1] * g1[t];
      gn_lo[t] = tb.E1[0] * gnl0[t] + tb.E1[1] * gnl1[t];   // lower cell, its xi_d = 1 face
      gn_hi[t] = tb.E0[0] * gnh0[t] + tb.E0[1] * gnh1[t];   // upper cell, its xi_d = 0 face
    }
  double cl[NT], ch[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) { cl[t] = 0.0; ch[t] = 0.0; }
  // face quadrature: traces at the point (fma chains over the face's NT nodes), upwind flux, back to the nodes
  auto face_point = [&](const double (&oo)[NT], double cfl, double cfh) {
    double g1l = 0.0, g2l = 0.0, g1h = 0.0, g2h = 0.0;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      g1l = __fma_rn(gn_lo[t], oo[t], g1l); g2l = __fma_rn(go0[t], oo[t], g2l);   // lower face: element 1 = lower cell, 2 = this cell
      g1h = __fma_rn(go1[t], oo[t], g1h); g2h = __fma_rn(gn_hi[t], oo[t], g2h);   // upper face: element 1 = this cell
    }
    const double fl = cfl * (cfl >= 0.0 ? g1l : g2l);
    const double fh = cfh * (cfh >= 0.0 ? g1h : g2h);
#pragma unroll
    for (int t = 0; t < NT; ++t) { cl[t] = __fma_rn(fl, oo[t], cl[t]); ch[t] = __fma_rn(fh, oo[t], ch[t]); }
  };
  if constexpr (HY) {
    // 9 points as a rolled loop (instruction-fetch friendly, like elem_apply).  The 18 face coefficients were sent on
    // their way to the thread's own shared-memory slots (cp.async, s_fn[point][lane]) before the prologue's loads:
    // their latency hides behind the prologue instead of stalling the loop (ncu: long-scoreboard at the first use of
    // a register-prefetched row was a quarter of all warp-stall samples).
    cp_async_wait<0>();
#if IMPES_K1QS_UNROLL3
#if IMPES_K1QS_UNROLL3 == 2
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int qf = 0; qf < 3; ++qf) {   // rows of three points: the loop overhead (uniform address arithmetic) is paid per row
#pragma unroll
      for (int qe = 0; qe < 3; ++qe) {
        const double* o = c_OOF3[qe][qf];
        const double oo[NT] = {o[0], o[1], o[2], o[3]};
        face_point(oo, s_fn[(qe + 3 * qf) * K1Q_TC], s_fn[(9 + qe + 3 * qf) * K1Q_TC]);
      }
    }
#else
#pragma unroll 1
    for (int p = 0; p < 9; ++p) {
      const int qf = p / 3, qe = p - 3 * qf;
      const double* o = c_OOF3[qe][qf];
      const double oo[NT] = {o[0], o[1], o[2], o[3]};
      face_point(oo, s_fn[p * K1Q_TC], s_fn[(9 + p) * K1Q_TC]);
    }
#endif
  } else {
#pragma unroll
    for (int qf = 0; qf < NFF; ++qf)
#pragma unroll
      for (int qe = 0; qe < NFE; ++qe) {
        double oo[NT];
#pragma unroll
        for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
          for (int te = 0; te < NOE; ++te) {
            const double oe = (NFE == 1) ? 1.0 : tb.OF[qe][te], of = (NFF == 1) ? 1.0 : tb.OF[qf][tf];
            oo[te + NOE * tf] = oe * of;
          }
        face_point(oo, fl_p[(qe + NFE * qf) * ncv], fh_p[(qe + NFE * qf) * ncv]);
      }
  }
#pragma unroll
  for (int tf = 0; tf < NOF; ++tf)
#pragma unroll
    for (int te = 0; te < NOE; ++te) {
      const int t = te + NOE * tf;
      const int lp = (D == 0) ? 2 * te + 2 * (HY ? 2 : 1) * tf : (D == 1 ? te + 4 * tf : te + 2 * tf);
      if (lo) {  // rows of element 2 of the lower face: + flux * psi_i(xi_d = 0)
        rd[lp] += tb.E0[0] * cl[t];
        rd[lp + PS1] += tb.E0[1] * cl[t];
      }
      if (hi) {  // rows of element 1 of the upper face: - flux * psi_i(xi_d = 1)
        rd[lp] -= tb.E1[0] * ch[t];
        rd[lp + PS1] -= tb.E1[1] * ch[t];
      }
    }
}

// One thread per (owned cell, direction) like K1 pass A: each thread forms its direction's part of Kt F (volume term +
// both faces of the cell along that direction) from coalesced loads, the three parts meet in shared memory as
// r = (r_x + r_y) + r_z, and the update runs one thread per DG dof (S is cell-major like the host's vector: coalesced).
// (Round 1's version: one thread per cell through all directions, 254 -> 128 registers with spills, 17 % of the HBM
// roofline.)
#ifndef IMPES_K1QS_REGS
#define IMPES_K1QS_REGS 128
#endif

template <bool HY, bool RATE>
__global__ void __launch_bounds__(K1Q_TC * (HY ? 3 : 2)) __maxnreg__(HY ? IMPES_K1QS_REGS : 80)
k1q_sat(Lay1 L, QTab tb, Phys ph, SatParam sp, long ncv, const double* __restrict__ vh, const double* __restrict__ fn,
        const double* __restrict__ phi, const double* __restrict__ Sin, const double* __restrict__ Fin,
        double* __restrict__ Sout, double* __restrict__ Fout) {
  using Q = QD<HY>;
  using SQ = SatQ<HY>;
  constexpr int ND = SQ::ND, THREADS = SQ::THREADS, PS = Q::NSL + 1;
  __shared__ double s_r[ND * K1Q_TC * PS];         // per-direction parts of Kt F
  __shared__ double s_f[HY ? ND * 18 * K1Q_TC : 1];   // 3-D: face coefficients, [direction][lower 9 | upper 9][lane]
  __shared__ int s_src[K1Q_TC];                       // cell may hold a DG node inside the injection box
  const int tid = threadIdx.x;
  const int ci = tid % K1Q_TC, dir = tid / K1Q_TC;
  const unsigned t0 = blockIdx.x * (unsigned)K1Q_TC;
  const unsigned nc = min((unsigned)K1Q_TC, (unsigned)L.ncl - t0);
  const long c0 = (long)t0 + L.plane;               // cvec index of the CTA's first cell (owned cells are contiguous)
  const long sb0 = c0 * Q::NSL;                     // its first DG dof in the cell-major S
  constexpr int NIT = (K1Q_TC * Q::NSL + THREADS - 1) / THREADS;
  // operands of the update pass (one DG dof per thread and pass): in flight during the flux computation
  double phi_r[NIT], sin_r[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const unsigned e = tid + it * THREADS;
    const bool in = e < nc * Q::NSL;
    phi_r[it] = phi[c0 + (in ? e / Q::NSL : 0)];
    sin_r[it] = RATE ? 0.0 : Sin[sb0 + (in ? e : 0)];
  }
  Cell1 q;
  {
    q.ok = (unsigned)ci < nc;
    const unsigned tt = q.ok ? t0 + ci : t0;
    const unsigned r = fastdiv(tt, L.fdx);
    q.i = (int)(tt - r * (unsigned)L.NX);
    q.kl = (int)fastdiv(r, L.fdy);
    q.j = (int)(r - (unsigned)q.kl * (unsigned)L.NY);
    q.k = q.kl + L.zlo;
    q.c = q.i + L.NX * (q.j + L.NY * (q.kl + 1));
    q.pb = L.off_p + q.c * L.npl;
  }
  if (dir == 0)
    s_src[ci] = q.i >= sp.ni0 && q.i <= sp.ni1 && q.j >= sp.nj0 && q.j <= sp.nj1 && q.k >= sp.nk0 && q.k <= sp.nk1;
  double* rd = s_r + (dir * K1Q_TC + ci) * PS;
  double* sfn = s_f + (HY ? dir * 18 * K1Q_TC + ci : 0);
  if (dir == 0) q_sat_dir<HY, 0>(L, q, tb, ncv, Fin, vh, fn, sfn, rd);
  else if (HY && dir == 1) { if constexpr (HY) q_sat_dir<HY, 1>(L, q, tb, ncv, Fin, vh, fn, sfn, rd); }
  else q_sat_dir<HY, 2>(L, q, tb, ncv, Fin, vh, fn, sfn, rd);
  __syncthreads();
  const double wm = sp.vol * tb.wv;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {   // one DG dof per thread and pass
    const unsigned e = tid + it * THREADS;
    if (e >= nc * Q::NSL) break;
    const unsigned c = e / Q::NSL, m = e - c * Q::NSL;
    const double ra = s_r[c * PS + m], rb = s_r[(K1Q_TC + c) * PS + m];
    const double r = HY ? (ra + rb) + s_r[(2 * K1Q_TC + c) * PS + m] : ra + rb;
    double z = r;
    if (s_src[c]) {   // rare: only cells at the injection well evaluate the source at their nodes
      const unsigned tt = t0 + c;
      const unsigned rr = fastdiv(tt, L.fdx);
      const int i = (int)(tt - rr * (unsigned)L.NX);
      const int kl = (int)fastdiv(rr, L.fdy);
      const int j = (int)(rr - (unsigned)kl * (unsigned)L.NY);
      const int k = kl + L.zlo;
      const int m0 = m & 1, m1 = HY ? (m >> 1) & 1 : 0, m2 = m >> (HY ? 2 : 1);
      // run-time node index: selects instead of indexing the kernel parameter (which would be copied to local memory)
      const double px = (i + (m0 ? tb.sn[1] : tb.sn[0])) * sp.hx, py = (j + (m1 ? tb.sn[1] : tb.sn[0])) * sp.hy,
                   pz = (k + (m2 ? tb.sn[1] : tb.sn[0])) * sp.hz;
      z = r + sat_source_at(sp, px, py, pz) * wm;
    }
    const double mphi = phi_r[it] * wm;
    double yv = div_rcp(z, mphi);   // same bits as `/` (kernels.cuh), without the call into the division subroutine
    double so;
    if (RATE) so = yv;
    else {
      yv *= sp.dt;
      so = sin_r[it] + yv;
    }
    Sout[sb0 + e] = so;
    if (!RATE) {
      double fo = so;
      if (sp.two_phase) {
        double mw, mo;
        mobilities(ph, so, mw, mo);
        fo = div_rcp(mw, mw + mo);   // FS [REF src/saturation_solver.cpp:8-17]
      }
      Fout[(long)m * ncv + c0 + c] = fo;   // dof-major
    }
  }
}

// ---------------------------------------------------------------------------------------
// -ov 0 -op 0 -os 1: RT0 / P0 velocity-pressure with the saturation in DG-Q1.  The RT0 mass block of a direction is
// the 2x2 matrix  (geo_d / K_d) sum_q (w_q / lambda(S(x_q))) phi_i(xi_d) phi_j(xi_d),  phi = (1 - xi, xi), with the
// rule of VectorFEMassIntegrator for RT0 (order OrderW + 2 el.order [MFEM-recall]: 3^3 / 2^2 points).  K4 stores its
// three entries per cell and direction, a[(3 dir + {00, 01, 11}) ncv + c]; K1 and the face coefficients are the RT0
// kernels with those entries in place of m/3, m/6 (kernels.cuh, VAR); K6 is k1q_sat fed with records formed from the
// RT0 face fluxes (an RT0 field is the RT1 field whose four face nodes carry the face value and whose interior nodes
// interpolate linearly).
// ---------------------------------------------------------------------------------------
template <bool HY>
__global__ void __launch_bounds__(BLK) k0q_coeff(Lay L, Phys ph, QTab tb, const double* __restrict__ S,
                                                 const double* __restrict__ Kx, const double* __restrict__ Ky,
                                                 const double* __restrict__ Kz, double gx, double gy, double gz,
                                                 double* __restrict__ a) {
  constexpr int NQ = HY ? 3 : 2, NQY = HY ? NQ : 1, NSL = HY ? 8 : 4;
  const long c = (long)blockIdx.x * BLK + threadIdx.x;
  if (c >= L.ncv) return;
  const long n = L.ncv;
  const int k = (int)(c / L.plane) - 1 + L.zlo;
  if (k < 0 || k >= L.NZ) {
    for (int e = 0; e < 9; ++e) a[e * n + c] = 0.0;
    return;
  }
  double s[NSL];
#pragma unroll
  for (int m = 0; m < NSL; ++m) s[m] = S[c * NSL + m];
  double ax[3] = {0.0, 0.0, 0.0}, ay[3] = {0.0, 0.0, 0.0}, az[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int q2 = 0; q2 < NQ; ++q2)
#pragma unroll
    for (int q1 = 0; q1 < NQY; ++q1)
#pragma unroll
      for (int q0 = 0; q0 < NQ; ++q0) {
        double sq = 0.0;
#pragma unroll
        for (int m = 0; m < NSL; ++m) {
          const int m0 = m & 1, m1 = HY ? (m >> 1) & 1 : 0, m2 = m >> (HY ? 2 : 1);
          sq += s[m] * (tb.O0[q0][m0] * (HY ? tb.O0[q1][m1] : 1.0) * tb.O0[q2][m2]);
        }
        const double w = tb.gw0[q0] * (HY ? tb.gw0[q1] : 1.0) * tb.gw0[q2];
        const double cq = w / lambda_of(ph, sq);
        const double x0 = tb.x0[q0], x1 = HY ? tb.x0[q1] : 0.0, x2 = tb.x0[q2];
        ax[0] += cq * ((1.0 - x0) * (1.0 - x0)); ax[1] += cq * ((1.0 - x0) * x0); ax[2] += cq * (x0 * x0);
        if (HY) { ay[0] += cq * ((1.0 - x1) * (1.0 - x1)); ay[1] += cq * ((1.0 - x1) * x1); ay[2] += cq * (x1 * x1); }
        az[0] += cq * ((1.0 - x2) * (1.0 - x2)); az[1] += cq * ((1.0 - x2) * x2); az[2] += cq * (x2 * x2);
      }
  const double fx = gx / Kx[c], fz = gz / Kz[c];
#pragma unroll
  for (int e = 0; e < 3; ++e) { a[e * n + c] = fx * ax[e]; a[(6 + e) * n + c] = fz * az[e]; }
  if (HY) {
    const double fy = gy / Ky[c];
#pragma unroll
    for (int e = 0; e < 3; ++e) a[(3 + e) * n + c] = fy * ay[e];
  }
}

// K6 records (see k1q_sat_pre) from the RT0 face fluxes u (dvec layout of kernels.cuh): along a direction the normal
// component is (1 - xi) u_lo + xi u_hi for all NT lines; at a face point it is the face value (the transverse open
// basis sums to one).
template <bool HY>
__global__ void __launch_bounds__(BLK) k0q_sat_pre(Lay L, QTab tb, const double* __restrict__ u, long ncv,
                                                   double* __restrict__ vh, double* __restrict__ fn) {
  using SQ = SatQ<HY>;
  const CellIdx ci = owned_cell(L);
  if (!ci.ok) return;
  const int i = ci.i, j = ci.j, kl = ci.kl, kk = kl + 1;
  const long c = ix_c(L, i, j, kk);
  const int fx = ix_ux(L, i, j, kl), fz = ix_uz(L, i, j, kk);
  const double lo[3] = {u[fx], HY ? u[ix_uy(L, i, j, kl)] : 0.0, u[fz]};
  const double hi[3] = {u[fx + 1], HY ? u[ix_uy(L, i, j, kl) + L.NX] : 0.0, u[fz + L.plane]};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (!HY && d == 1) continue;
    const int di = HY ? d : (d == 0 ? 0 : 1);
    const double v0 = (1.0 - tb.sn[0]) * lo[d] + tb.sn[0] * hi[d], v1 = (1.0 - tb.sn[1]) * lo[d] + tb.sn[1] * hi[d];
#pragma unroll
    for (int t = 0; t < SQ::NT; ++t) {
      vh[((long)di * SQ::NVH + 2 * t) * ncv + c] = v0;
      vh[((long)di * SQ::NVH + 2 * t + 1) * ncv + c] = v1;
    }
    const bool top = d == 2 && kl == L.nzl - 1;
#pragma unroll
    for (int p = 0; p < SQ::NFP; ++p) {
      const int qe = HY ? p % 3 : p, qf = HY ? p / 3 : 0;
      const double w = HY ? tb.fw[qe] * tb.fw[qf] : tb.fw[qe];
      fn[((long)di * SQ::NFP + p) * ncv + c] = w * lo[d];
      if (top) fn[((long)di * SQ::NFP + p) * ncv + c + L.plane] = w * hi[d];
    }
  }
}

// f(S) over a flat dof range (ghost planes included)
// S: cell-major (nsl dofs per cell); F: dof-major, F[m ncv + c] (the layout of the DG-Q1 stage kernel)
__global__ void __launch_bounds__(BLK) k_frac_flat(long ncv, int nsl, Phys ph, int two_phase, const double* __restrict__ S,
                                                   double* __restrict__ F) {
  const long n = ncv * nsl;
  for (long i = (long)blockIdx.x * BLK + threadIdx.x; i < n; i += (long)gridDim.x * BLK) {
    const double s = S[i];
    const long c = i / nsl;
    const int m = (int)(i - c * nsl);
    F[(long)m * ncv + c] = two_phase ? frac_flow_of(ph, s) : s;
  }
}
// guards over a flat owned dof range
__global__ void __launch_bounds__(BLK) k_guards_flat(long n, const double* __restrict__ S, unsigned long long* n_bad,
                                                     unsigned long long* max_bits) {
  unsigned bad = 0;
  unsigned long long key = 0ull;
  for (long i = (long)blockIdx.x * BLK + threadIdx.x; i < n; i += (long)gridDim.x * BLK) {
    const double s = S[i];
    if (!isfinite(s)) { ++bad; continue; }
    const unsigned long long b = guard_key(s);
    key = b > key ? b : key;
  }
  guard_commit(bad, key, n_bad, max_bits);
}

}  // namespace impes
