// rt1.cuh -- sm_100a kernels of the IMPES pressure path for RT1 / L2-Q1 (the reference's DEFAULT orders
// `-ov 1 -op 1` [REF src/param.cpp:43]) on a Cartesian grid, saturation in DG_0 (`-os 0`).
//
// Discrete structure: with MFEM's nodal bases (closed = Gauss-Lobatto {0,1/2,1} along the
// component's own direction, open = the 2 Gauss points transversally; L2-Q1 nodes = the same Gauss points)
// and a cell-constant coefficient,
//     M_d|cell = m_d * Mc (x) Mo (x) Mo,     B|cell,d = -G (x) Mo (x) Mo,     Mo = diag(1/2, 1/2),
// i.e. every operator acts on LINES of nodes along d with fixed transverse open-node indices: a 3-node
// quadratic mass block per cell and line, and two pressure dofs per cell and line.  Nothing is assembled.
//
// Layout ("dvec1", per rank; slab along internal z like layout.h; 2-D (nx, ny) is stored as (nx, 1, ny) with
// one open node of weight 1 along the absent direction):
//   ux : (2NX+1) x (oy NY) x (2 nzl)          node (2i+a, oy j+t1, 2 kl+t2)
//   uy : (2NX) x (2NY+1) x (2 nzl)            (absent when !has_y)
//   uz : (2NX) x (oy NY) x (2 nzl + 3)        node plane gz - 2 zlo + 2  (two ghost node planes below, one above)
//   p  : (nzl + 2) cell planes x npl          cell-major, local index a0 + 2 (a1 + oy a2)   (= MFEM's L2 order)
// which is the native host layout of include/impes_b200.h restricted to the slab, so transfers are
// contiguous copies.  Arithmetic contract: as kernels.cuh (-fmad=false, CSR column order of the oracle).
#pragma once
#include "kernels.cuh"
#include "mg.cuh"

namespace impes {

struct Lay1 {
  FastDiv fdx, fdy;  // division by NX, NY (layout.h)
  int NX, NY, NZ, has_y, oy, npl;
  int zlo, zhi, nzl, rank, nranks;
  int plane, ncl, nblk;
  int gx0, gx1, gy0, gy1, gz0, gz1;  // node-grid extents (x, y) of the three components
  int off_x, off_y, off_z, off_p, ntot;
};

// Reference-element constants, already multiplied by the transverse open-node weights
// wt = 1/4 (3-D) or 1/2 (2-D):  CW = Mc * wt,  Bv = -(G * wt).
struct Rt1Const {
  double CW[3][3];
  double Bv[2][3];
  double wflux;  // weight of one face node in the face flux: 1/4 (3-D), 1/2 (2-D)
};

// Bv[a][i] with a possibly run-time a: a select between two constant-bank operands (a dynamically indexed
// kernel parameter would be copied to local memory)
#define BV(a, i) ((a) ? C.Bv[1][i] : C.Bv[0][i])

struct Cell1 {
  int i, j, kl, k, c, pb;  // c: cvec index (ghost plane below included), pb: index of the cell's first p dof
  bool ok;
};
__device__ __forceinline__ Cell1 owned_cell1(const Lay1& L) {
  Cell1 q;
  const unsigned t = blockIdx.x * (unsigned)BLK + threadIdx.x;
  q.ok = t < (unsigned)L.ncl;
  const unsigned tt = q.ok ? t : 0u;
  const unsigned r = fastdiv(tt, L.fdx);
  q.i = (int)(tt - r * (unsigned)L.NX);
  q.kl = (int)fastdiv(r, L.fdy);
  q.j = (int)(r - (unsigned)q.kl * (unsigned)L.NY);
  q.k = q.kl + L.zlo;
  q.c = q.i + L.NX * (q.j + L.NY * (q.kl + 1));
  q.pb = L.off_p + q.c * L.npl;
  return q;
}

// Line geometry of one (cell, direction, transverse pair): index of closed node 0, stride along the line,
// position of the cell along the line, local index of the line's first pressure dof and the stride between
// its two pressure dofs, and the cell-index stride to the lower neighbour along the line.
struct Line1 {
  int base, sd, cd, nd, lp, ps, cstride, pdn;
};
template <bool HY>
__device__ __forceinline__ Line1 line_x(const Lay1& L, const Cell1& q, int t1, int t2) {
  constexpr int oy = HY ? 2 : 1;
  Line1 l;
  l.base = L.off_x + 2 * q.i + L.gx0 * ((oy * q.j + t1) + L.gx1 * (2 * q.kl + t2));
  l.sd = 1; l.cd = q.i; l.nd = L.NX; l.lp = 2 * t1 + 2 * oy * t2; l.ps = 1; l.cstride = 1; l.pdn = L.npl;
  return l;
}
__device__ __forceinline__ Line1 line_y(const Lay1& L, const Cell1& q, int t0, int t2) {
  Line1 l;
  l.base = L.off_y + (2 * q.i + t0) + L.gy0 * (2 * q.j + L.gy1 * (2 * q.kl + t2));
  l.sd = L.gy0; l.cd = q.j; l.nd = L.NY; l.lp = t0 + 4 * t2; l.ps = 2; l.cstride = L.NX; l.pdn = L.NX * L.npl;
  return l;
}
template <bool HY>
__device__ __forceinline__ Line1 line_z(const Lay1& L, const Cell1& q, int t0, int t1) {
  constexpr int oy = HY ? 2 : 1;
  Line1 l;
  l.base = L.off_z + (2 * q.i + t0) + L.gz0 * ((oy * q.j + t1) + L.gz1 * (2 * q.kl + 2));
  l.sd = L.gz0 * L.gz1; l.cd = q.k; l.nd = L.NZ; l.lp = t0 + 2 * t1; l.ps = 2 * oy; l.cstride = L.plane;
  l.pdn = L.plane * L.npl;
  return l;
}

// ---------------------------------------------------------------------------------------
// K4 (RT1): diag(M) and T = 1/diag(M) per velocity node (1 / 0 on eliminated boundary nodes
// [REF src/pressure_solver.cpp:56-60]) and the diagonal of S = B diag(M)^-1 B^T [REF :76-81] per pressure dof
// (stored in the p-part of Md).  m_d: k_cell_coeff's per-cell coefficient (cvec layout).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void line_coeff(const Rt1Const& C, const Line1& l, const double* __restrict__ m, int c,
                                           double* __restrict__ Md, double* __restrict__ T, bool own_top,
                                           double& D0, double& D1) {
  const double mR = m[c];
  double t0 = 0.0, t2 = 0.0;
  if (l.cd > 0) {
    const double md = m[c - l.cstride] * C.CW[2][2] + mR * C.CW[0][0];
    t0 = 1.0 / md;
    Md[l.base] = md; T[l.base] = t0;
  } else { Md[l.base] = 1.0; T[l.base] = 0.0; }
  const double md1 = mR * C.CW[1][1];
  const double t1 = 1.0 / md1;
  Md[l.base + l.sd] = md1; T[l.base + l.sd] = t1;
  if (l.cd + 1 < l.nd) {
    const double md = mR * C.CW[2][2] + m[c + l.cstride] * C.CW[0][0];
    t2 = 1.0 / md;
    if (own_top) { Md[l.base + 2 * l.sd] = md; T[l.base + 2 * l.sd] = t2; }  // ghost-above node plane of a slab
  } else { Md[l.base + 2 * l.sd] = 1.0; T[l.base + 2 * l.sd] = 0.0; }
  // S_pp = sum over the line's nodes, ascending:  b (b T)   (boundary nodes: eliminated columns, skipped)
  if (l.cd > 0) { D0 += C.Bv[0][0] * (C.Bv[0][0] * t0); D1 += C.Bv[1][0] * (C.Bv[1][0] * t0); }
  D0 += C.Bv[0][1] * (C.Bv[0][1] * t1); D1 += C.Bv[1][1] * (C.Bv[1][1] * t1);
  if (l.cd + 1 < l.nd) { D0 += C.Bv[0][2] * (C.Bv[0][2] * t2); D1 += C.Bv[1][2] * (C.Bv[1][2] * t2); }
}

template <bool HY>
__global__ void __launch_bounds__(BLK) k1_node_coeff(Lay1 L, Rt1Const C, const double* __restrict__ mx,
                                                     const double* __restrict__ my, const double* __restrict__ mz,
                                                     double* __restrict__ Md, double* __restrict__ T) {
  constexpr int oy = HY ? 2 : 1;
  constexpr int NPL = 4 * oy;
  const Cell1 q = owned_cell1(L);
  if (!q.ok) return;
  double D[NPL];
#pragma unroll
  for (int l = 0; l < NPL; ++l) D[l] = 0.0;
#pragma unroll
  for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
    for (int t1 = 0; t1 < oy; ++t1) {
      const Line1 l = line_x<HY>(L, q, t1, t2);
      const int lp = 2 * t1 + 2 * oy * t2;
      line_coeff(C, l, mx, q.c, Md, T, false, D[lp], D[lp + 1]);
    }
  if constexpr (HY) {
#pragma unroll
    for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
      for (int t0 = 0; t0 < 2; ++t0) {
        const Line1 l = line_y(L, q, t0, t2);
        const int lp = t0 + 4 * t2;
        line_coeff(C, l, my, q.c, Md, T, false, D[lp], D[lp + 2]);
      }
  }
#pragma unroll
  for (int t1 = 0; t1 < oy; ++t1)
#pragma unroll
    for (int t0 = 0; t0 < 2; ++t0) {
      const Line1 l = line_z<HY>(L, q, t0, t1);
      const int lp = t0 + 2 * t1;
      line_coeff(C, l, mz, q.c, Md, T, q.k == L.zhi - 1, D[lp], D[lp + 2 * oy]);
    }
#pragma unroll
  for (int l = 0; l < NPL; ++l) Md[q.pb + l] = D[l];
}

// ---------------------------------------------------------------------------------------
// K1 (RT1): y = [M(S) B^T; B 0] x, matrix-free, boundary rows/cols eliminated [REF src/pressure_solver.cpp:47-74];
// optional fused <x, y>.  One thread per L2 dof (2^dim consecutive lanes per cell, see owned_dof1): the thread
// with local index (a0, a1, a2) computes its pressure row and, on each of its three lines, the row of closed
// node a_d (0 = the cell's lower face node, 1 = its interior node; node 2 only on the upper boundary) -- 8 x 3
// = the 24 velocity rows a cell owns.  Gather form, no atomics; sums in the oracle's CSR column order (M row,
// then the B^T row added to it).  The other pressure dof of a line comes from the neighbouring lane.
// ---------------------------------------------------------------------------------------
struct Dof1 {
  Cell1 q;
  int l;
};
template <bool HY>
__device__ __forceinline__ Dof1 owned_dof1(const Lay1& L) {
  constexpr int NPL = HY ? 8 : 4;
  Dof1 d;
  const unsigned t = blockIdx.x * (unsigned)BLK + threadIdx.x;
  const unsigned cell = t / NPL;
  d.l = (int)(t % NPL);
  Cell1& q = d.q;
  q.ok = cell < (unsigned)L.ncl;
  const unsigned tt = q.ok ? cell : 0u;
  const unsigned r = fastdiv(tt, L.fdx);
  q.i = (int)(tt - r * (unsigned)L.NX);
  q.kl = (int)fastdiv(r, L.fdy);
  q.j = (int)(r - (unsigned)q.kl * (unsigned)L.NY);
  q.k = q.kl + L.zlo;
  q.c = q.i + L.NX * (q.j + L.NY * (q.kl + 1));
  q.pb = q.c * NPL;  // p-part relative (callers add L.off_p for dvec1 addressing)
  return d;
}

// Branch-free: every load is unconditional (addresses clamped into the cell's own line where the neighbour does
// not exist) so that the compiler issues all of a thread's loads up front -- one memory round trip instead of
// one per branch region (SASS of the branchy form: 6 dependent LDG groups) -- and the skipped CSR entries
// become `+ 0.0`, which leaves every partial sum bit-identical.
struct LineData {
  double mR, mL, x0, x1, x2, xm1, xm2, pL0, pL1;
};
__device__ __forceinline__ LineData line_load(const Line1& l, const double* __restrict__ m, int c,
                                              const double* __restrict__ x, int pb) {
  const bool lo = l.cd > 0;
  LineData v;
  v.mR = m[c];
  v.mL = m[lo ? c - l.cstride : c];
  v.x0 = x[l.base]; v.x1 = x[l.base + l.sd]; v.x2 = x[l.base + 2 * l.sd];
  v.xm1 = x[lo ? l.base - l.sd : l.base];
  v.xm2 = x[lo ? l.base - 2 * l.sd : l.base];
  const double* pL = x + (pb - (lo ? l.pdn : 0) + l.lp);
  v.pL0 = pL[0]; v.pL1 = pL[l.ps];
  return v;
}
template <bool DOT>
__device__ __forceinline__ void line_rows_dof(const Rt1Const& C, const Line1& l, const LineData& v, int a,
                                              double* __restrict__ y, double p_lo, double p_hi, bool ok, double& yp,
                                              double& acc) {
  const bool lo = l.cd > 0, lo2 = l.cd > 1, hi = l.cd + 1 < l.nd;
  const double mR = v.mR, mL = v.mL, x0 = v.x0, x1 = v.x1, x2 = v.x2;
  const bool fa = a == 0;  // face row (closed node 0) or interior row (closed node 1)
  // M row, ascending columns: x(-2), x(-1) [lower cell, face rows only], x0, x1, x2
  const double c3 = fa ? (mL * C.CW[2][2] + mR * C.CW[0][0]) : (mR * C.CW[1][0]);
  const double c4 = mR * (fa ? C.CW[0][1] : C.CW[1][1]);
  const double c5 = mR * (fa ? C.CW[0][2] : C.CW[1][2]);
  double s = 0.0;
  s += (fa && lo2) ? (mL * C.CW[2][0]) * v.xm2 : 0.0;
  s += fa ? (mL * C.CW[2][1]) * v.xm1 : 0.0;
  s += lo ? c3 * x0 : 0.0;
  s += c4 * x1;
  s += hi ? c5 * x2 : 0.0;
  // B^T row: lower cell's two dofs (face rows only), then this cell's two
  double t = 0.0;
  t += fa ? C.Bv[0][2] * v.pL0 : 0.0;
  t += fa ? C.Bv[1][2] * v.pL1 : 0.0;
  t += (fa ? C.Bv[0][0] : C.Bv[0][1]) * p_lo;
  t += (fa ? C.Bv[1][0] : C.Bv[1][1]) * p_hi;
  const double xa = fa ? x0 : x1;
  const double ya = (fa && !lo) ? x0 : s + t;  // eliminated boundary row: identity
  if (ok) y[l.base + (fa ? 0 : l.sd)] = ya;
  if (DOT) acc += xa * ya;
  if (!fa && !hi) {  // upper boundary node: identity row
    if (ok) y[l.base + 2 * l.sd] = x2;
    if (DOT) acc += x2 * x2;
  }
  yp += lo ? BV(a, 0) * x0 : 0.0;
  yp += BV(a, 1) * x1;
  yp += hi ? BV(a, 2) * x2 : 0.0;
}

// Branch-free: every load is unconditional (addresses clamped into the cell's own line where the neighbour does
// not exist) and all of a thread's loads are issued before any arithmetic -- one memory round trip instead of
// one per line / branch region (SASS of the branchy form: 6 dependent LDG groups).  Skipped CSR entries become
// `+ 0.0`, which leaves every partial sum bit-identical.
#ifndef IMPES_K1_MINB
#define IMPES_K1_MINB 4
#endif
template <bool HY, bool DOT>
__global__ void __launch_bounds__(BLK, IMPES_K1_MINB) k1_darcy_apply(Lay1 L, Rt1Const C, const double* __restrict__ mx,
                                                      const double* __restrict__ my, const double* __restrict__ mz,
                                                      const double* __restrict__ x, double* __restrict__ y,
                                                      double* partial, double* out, unsigned* ticket,
                                                      const MinresState* st) {
  if (st && st->done) return;
  const Dof1 d = owned_dof1<HY>(L);
  const Cell1& q = d.q;
  const int l = d.l;
  const int a0 = l & 1, a1 = HY ? (l >> 1) & 1 : 0, a2 = l >> (HY ? 2 : 1);
  const int pb = L.off_p + q.pb;
  const bool ok = q.ok;
  const Line1 lx = line_x<HY>(L, q, a1, a2);
  const Line1 lz = line_z<HY>(L, q, a0, a1);
  Line1 ly = lx;
  if constexpr (HY) ly = line_y(L, q, a0, a2);
  // ---- load phase
  const double pc = x[pb + l];
  const LineData vx = line_load(lx, mx, q.c, x, pb);
  LineData vy = vx;
  if constexpr (HY) vy = line_load(ly, my, q.c, x, pb);
  const LineData vz = line_load(lz, mz, q.c, x, pb);
  // ---- compute phase
  const double px = __shfl_xor_sync(0xffffffffu, pc, 1);
  const double py = HY ? __shfl_xor_sync(0xffffffffu, pc, 2) : 0.0;
  const double pz = __shfl_xor_sync(0xffffffffu, pc, HY ? 4 : 2);
  double acc = 0.0, yp = 0.0;
  line_rows_dof<DOT>(C, lx, vx, a0, y, a0 ? px : pc, a0 ? pc : px, ok, yp, acc);
  if constexpr (HY) line_rows_dof<DOT>(C, ly, vy, a1, y, a1 ? py : pc, a1 ? pc : py, ok, yp, acc);
  line_rows_dof<DOT>(C, lz, vz, a2, y, a2 ? pz : pc, a2 ? pc : pz, ok, yp, acc);
  if (ok) y[pb + l] = yp;
  if (DOT) {
    acc += pc * yp;
    finish_dot(ok ? acc : 0.0, partial, out, ticket);
  }
}

// Canonical dot <a, b> over the dofs each thread of k1_darcy_apply owns, in the order it visits them.
__device__ __forceinline__ void line_dot_dof(const Line1& l, int a, const double* __restrict__ u,
                                             const double* __restrict__ v, double& acc) {
  if (a == 0) acc += u[l.base] * v[l.base];
  else {
    acc += u[l.base + l.sd] * v[l.base + l.sd];
    if (l.cd + 1 == l.nd) acc += u[l.base + 2 * l.sd] * v[l.base + 2 * l.sd];
  }
}
template <bool HY>
__global__ void __launch_bounds__(BLK) k1_dot_owned(Lay1 L, const double* __restrict__ a, const double* __restrict__ b,
                                                    double* partial, double* out, unsigned* ticket,
                                                    const MinresState* st) {
  if (st && st->done) return;
  const Dof1 d = owned_dof1<HY>(L);
  const Cell1& q = d.q;
  const int l = d.l;
  const int a0 = l & 1, a1 = HY ? (l >> 1) & 1 : 0, a2 = l >> (HY ? 2 : 1);
  double acc = 0.0;
  if (q.ok) {
    line_dot_dof(line_x<HY>(L, q, a1, a2), a0, a, b, acc);
    if constexpr (HY) line_dot_dof(line_y(L, q, a0, a2), a1, a, b, acc);
    line_dot_dof(line_z<HY>(L, q, a0, a1), a2, a, b, acc);
    const int p = L.off_p + q.pb + l;
    acc += a[p] * b[p];
  }
  finish_dot(acc, partial, out, ticket);
}

// ---------------------------------------------------------------------------------------
// K3 (RT1), parity mode: one symmetric Gauss-Seidel sweep on S = B diag(M)^-1 B^T in the natural order of the
// L2 dofs (cell-major, 2^dim dofs per cell) -- GSSmoother(S) [REF src/pressure_solver.cpp:84].  S couples a dof
// to its line partners: the other dof of its cell on each line (3 shared nodes) and the two dofs of each
// line-neighbour cell (1 shared face node), so lexicographic GS is a wavefront over CELLS (hyperplanes
// i+j+k) with the cell's dofs swept sequentially by one thread.  Entries of S are formed on the fly from
// T = 1/diag(M) in the oracle's SpGEMM order:  S[p,p'] = sum over shared nodes f ascending of  b_pf (b_p'f T_f).
// Row sums run over ascending columns like SparseMatrix::Gauss_Seidel_forw/back [MFEM-recall].
// ---------------------------------------------------------------------------------------
// in-cell partner entry for the row with line index a (partner 1-a)
__device__ __forceinline__ double s_in(const Rt1Const& C, int a, double T0, double T1, double T2) {
  double v = BV(a, 0) * (BV(1 - a, 0) * T0);
  v += BV(a, 1) * (BV(1 - a, 1) * T1);
  v += BV(a, 2) * (BV(1 - a, 2) * T2);
  return v;
}

// contributions of the lower / upper line-neighbour cell to row (line index a)
__device__ __forceinline__ void sgs_lower(const Rt1Const& C, const Line1& l, int a, const double* __restrict__ T,
                                          const double* yp, int pb, double& sum) {
  if (l.cd > 0) {
    const double T0 = T[l.base];
    const double* yl = yp + (pb - l.pdn + l.lp);
    sum += yl[0] * (BV(a, 0) * (C.Bv[0][2] * T0));
    sum += yl[l.ps] * (BV(a, 0) * (C.Bv[1][2] * T0));
  }
}
__device__ __forceinline__ void sgs_upper(const Rt1Const& C, const Line1& l, int a, const double* __restrict__ T,
                                          const double* yp, int pb, double& sum) {
  if (l.cd + 1 < l.nd) {
    const double T2 = T[l.base + 2 * l.sd];
    const double* yu = yp + (pb + l.pdn + l.lp);
    sum += yu[0] * (BV(a, 2) * (C.Bv[0][0] * T2));
    sum += yu[l.ps] * (BV(a, 2) * (C.Bv[1][0] * T2));
  }
}
__device__ __forceinline__ double sgs_in(const Rt1Const& C, const Line1& l, int a, const double* __restrict__ T) {
  return s_in(C, a, l.cd > 0 ? T[l.base] : 0.0, T[l.base + l.sd], l.cd + 1 < l.nd ? T[l.base + 2 * l.sd] : 0.0);
}

// Off-diagonal part of row l of S for one cell: sum over ascending columns of y[col] * S[row, col].
// FULL = false: only columns below the row (forward sweep from a zero guess: the others multiply zeros).
// yl: the cell's own dofs (registers); y: p-addressed array for the neighbour cells' dofs (q.pb-relative).
template <bool HY, bool FULL>
__device__ __forceinline__ double q1_row_sum(const Lay1& L, const Rt1Const& C, const Cell1& q,
                                             const double* __restrict__ T, const double* y,
                                             const double (&yl)[HY ? 8 : 4], int l) {
  constexpr int oy = HY ? 2 : 1;
  constexpr int NPL = 4 * oy;
  const int a0 = l & 1, a1 = HY ? (l >> 1) & 1 : 0, a2 = l >> (HY ? 2 : 1);
  const Line1 lx = line_x<HY>(L, q, a1, a2);
  const Line1 lz = line_z<HY>(L, q, a0, a1);
  double sum = 0.0;
  // lower cells, ascending cell index: z-, y-, x-
  sgs_lower(C, lz, a2, T, y, q.pb, sum);
  if constexpr (HY) sgs_lower(C, line_y(L, q, a0, a2), a1, T, y, q.pb, sum);
  sgs_lower(C, lx, a0, T, y, q.pb, sum);
  // in-cell partners with a smaller local index, ascending: flip a2, a1, a0 from 1 to 0
  if (a2 == 1) sum += yl[l >= 2 * oy ? l - 2 * oy : 0] * sgs_in(C, lz, 1, T);
  if constexpr (HY) {
    if (a1 == 1) sum += yl[l >= 2 ? l - 2 : 0] * sgs_in(C, line_y(L, q, a0, a2), 1, T);
  }
  if (a0 == 1) sum += yl[l >= 1 ? l - 1 : 0] * sgs_in(C, lx, 1, T);
  if (FULL) {
    // in-cell partners with a larger local index, ascending: flip a0, a1, a2 from 0 to 1
    if (a0 == 0) sum += yl[(l + 1) % NPL] * sgs_in(C, lx, 0, T);
    if constexpr (HY) {
      if (a1 == 0) sum += yl[(l + 2) % NPL] * sgs_in(C, line_y(L, q, a0, a2), 0, T);
    }
    if (a2 == 0) sum += yl[(l + 2 * oy) % NPL] * sgs_in(C, lz, 0, T);
    // upper cells: x+, y+, z+
    sgs_upper(C, lx, a0, T, y, q.pb, sum);
    if constexpr (HY) sgs_upper(C, line_y(L, q, a0, a2), a1, T, y, q.pb, sum);
    sgs_upper(C, lz, a2, T, y, q.pb, sum);
  }
  return sum;
}

// one cell of a sweep.  BWD = false: forward sweep from a zero guess; BWD = true: backward sweep over all columns.
template <bool HY, bool BWD>
__device__ __forceinline__ void sgs_cell(const Lay1& L, const Rt1Const& C, const Cell1& q, const double* __restrict__ T,
                                         const double* __restrict__ D, const double* __restrict__ r, double* y) {
  constexpr int NPL = HY ? 8 : 4;
  double yl[NPL];
#pragma unroll
  for (int l = 0; l < NPL; ++l) yl[l] = BWD ? y[q.pb + l] : 0.0;
#pragma unroll
  for (int ll = 0; ll < NPL; ++ll) {
    const int l = BWD ? NPL - 1 - ll : ll;
    const double sum = q1_row_sum<HY, BWD>(L, C, q, T, y, yl, l);
    yl[l] = (r[q.pb + l] - sum) / D[q.pb + l];
  }
#pragma unroll
  for (int l = 0; l < NPL; ++l) y[q.pb + l] = yl[l];
}

template <int MODE, bool HY>
__global__ void __launch_bounds__(1024) k1_sgs_wavefront(Lay1 L, Rt1Const C, const double* __restrict__ T,
                                                         const double* __restrict__ Md, const double* __restrict__ r,
                                                         double* y, int do_fwd, int do_bwd, const MinresState* st) {
  if (st && st->done) return;
  cg::grid_group grid = cg::this_grid();
  const int nstage = L.NX + L.NY + L.nzl - 2;
  const long npair = (long)L.NY * L.nzl;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long)gridDim.x * blockDim.x;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 0 ? !do_fwd : !do_bwd) continue;
    for (int s = 0; s < nstage; ++s) {
      const int d = pass == 0 ? s : nstage - 1 - s;
      for (long jk = tid; jk < npair; jk += nth) {
        Cell1 q;
        q.j = (int)(jk % L.NY);
        q.kl = (int)(jk / L.NY);
        q.i = d - q.j - q.kl;
        if (q.i < 0 || q.i >= L.NX) continue;
        q.k = q.kl + L.zlo;
        q.c = q.i + L.NX * (q.j + L.NY * (q.kl + 1));
        q.pb = L.off_p + q.c * L.npl;
        q.ok = true;
        if (pass == 0) sgs_cell<HY, false>(L, C, q, T, Md, r, y);
        else sgs_cell<HY, true>(L, C, q, T, Md, r, y);
      }
      if (!(pass == 1 && s == nstage - 1)) wave_sync<MODE>(grid);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Face fluxes of the RT1 velocity in the RT0 dvec layout (layout.h) for the DG_0 saturation step:
// int_F v.n = sum of the face's RT dofs x Gauss weights [REF src/saturation_solver.cpp:32-35],
// summed second transverse direction outer, first inner.  The K6 kernels then run unchanged.
// ---------------------------------------------------------------------------------------
template <bool HY>
__global__ void __launch_bounds__(BLK) k1_face_flux(Lay1 L1, Lay L0, Rt1Const C, const double* __restrict__ x,
                                                    double* __restrict__ flux) {
  constexpr int oy = HY ? 2 : 1;
  const Cell1 q = owned_cell1(L1);
  if (!q.ok) return;
  {
    double F0 = 0.0, F2 = 0.0;
#pragma unroll
    for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
      for (int t1 = 0; t1 < oy; ++t1) {
        const Line1 l = line_x<HY>(L1, q, t1, t2);
        F0 += C.wflux * x[l.base];
        F2 += C.wflux * x[l.base + 2];
      }
    const int f = ix_ux(L0, q.i, q.j, q.kl);
    flux[f] = F0;
    if (q.i + 1 == L1.NX) flux[f + 1] = F2;
  }
  if constexpr (HY) {
    double F0 = 0.0, F2 = 0.0;
#pragma unroll
    for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
      for (int t0 = 0; t0 < 2; ++t0) {
        const Line1 l = line_y(L1, q, t0, t2);
        F0 += C.wflux * x[l.base];
        F2 += C.wflux * x[l.base + 2 * l.sd];
      }
    const int f = ix_uy(L0, q.i, q.j, q.kl);
    flux[f] = F0;
    if (q.j + 1 == L1.NY) flux[f + L0.NX] = F2;
  }
  {
    double F0 = 0.0, F2 = 0.0;
#pragma unroll
    for (int t1 = 0; t1 < oy; ++t1)
#pragma unroll
      for (int t0 = 0; t0 < 2; ++t0) {
        const Line1 l = line_z<HY>(L1, q, t0, t1);
        F0 += C.wflux * x[l.base];
        F2 += C.wflux * x[l.base + 2 * l.sd];
      }
    const int f = ix_uz(L0, q.i, q.j, q.kl + 1);
    flux[f] = F0;
    if (q.k + 1 == L1.zhi) flux[f + L0.plane] = F2;  // upper face plane of the slab (ghost above / top boundary)
  }
}

// ---------------------------------------------------------------------------------------
// K3 (RT1), throughput mode: the finest level of the multigrid V-cycle (mg.cuh) acts on the L2-Q1 dofs with the
// operator S itself (rows formed on the fly as above); it is aggregated cell-wise (the 2^dim dofs of a cell ->
// one coarse unknown) onto the cell-level two-point hierarchy of mg.cuh.  All arrays here are p-parts
// (index cell*npl + local, ghost plane below included).
// ---------------------------------------------------------------------------------------
// One thread per L2 dof (2^dim consecutive lanes per cell): every array access is coalesced and the cell's
// other dofs come through warp shuffles.  Same terms in the same order as q1_row_sum<HY, true>, so the result
// is bit-identical to the one-thread-per-cell form (ncu: that form was L1/LSU bound, 95 % l1tex throughput at
// 18 % DRAM, because each lane walked its own 64-byte block with 8-byte loads).
// Branch-free neighbour terms for the per-dof kernels: T is stored as 0 on eliminated boundary nodes, so a
// missing neighbour contributes y * (b (b' 0)) = +-0 with y loaded from a clamped (own-cell) address.
// e: cfac * (coarse correction of the neighbour cell) when the prolongation is fused into the sweep (PRO), else unused
template <bool PRO>
__device__ __forceinline__ void nb_lower(const Rt1Const& C, const Line1& l, int a, double T0, const double* y, int pb,
                                         double e, double& sum) {
  const double* yl = y + (pb - (l.cd > 0 ? l.pdn : 0) + l.lp);
  sum += (PRO ? yl[0] + e : yl[0]) * (BV(a, 0) * (C.Bv[0][2] * T0));
  sum += (PRO ? yl[l.ps] + e : yl[l.ps]) * (BV(a, 0) * (C.Bv[1][2] * T0));
}
template <bool PRO>
__device__ __forceinline__ void nb_upper(const Rt1Const& C, const Line1& l, int a, double T2, const double* y, int pb,
                                         double e, double& sum) {
  const double* yu = y + (pb + (l.cd + 1 < l.nd ? l.pdn : 0) + l.lp);
  sum += (PRO ? yu[0] + e : yu[0]) * (BV(a, 2) * (C.Bv[0][0] * T2));
  sum += (PRO ? yu[l.ps] + e : yu[l.ps]) * (BV(a, 2) * (C.Bv[1][0] * T2));
}
// off-diagonal part of row l (all columns); ys = y of this dof; all 32 lanes must call.
// PRO: y is the iterate BEFORE the prolongation y += cfac * ec(cell) (k1_mg_prolong); the correction is added on the
// fly to every neighbour value (same expression, same bits), ec = cell-level iterate with one ghost plane below.
template <bool HY, bool PRO = false>
__device__ __forceinline__ double q1_row_sum_dof(const Lay1& L, const Rt1Const& C, const Dof1& d,
                                                 const double* __restrict__ T, const double* __restrict__ y, double ys,
                                                 const double* __restrict__ ec = nullptr, double cfac = 0.0) {
  const Cell1& q = d.q;
  const int l = d.l;
  const int a0 = l & 1, a1 = HY ? (l >> 1) & 1 : 0, a2 = l >> (HY ? 2 : 1);
  const double yx = __shfl_xor_sync(0xffffffffu, ys, 1);
  const double yy = HY ? __shfl_xor_sync(0xffffffffu, ys, 2) : 0.0;
  const double yz = __shfl_xor_sync(0xffffffffu, ys, HY ? 4 : 2);
  const Line1 lx = line_x<HY>(L, q, a1, a2);
  const Line1 lz = line_z<HY>(L, q, a0, a1);
  Line1 ly = lx;
  if constexpr (HY) ly = line_y(L, q, a0, a2);
  const double Tx0 = T[lx.base], Tx1 = T[lx.base + lx.sd], Tx2 = T[lx.base + 2 * lx.sd];
  const double Ty0 = HY ? T[ly.base] : 0.0, Ty1 = HY ? T[ly.base + ly.sd] : 0.0, Ty2 = HY ? T[ly.base + 2 * ly.sd] : 0.0;
  const double Tz0 = T[lz.base], Tz1 = T[lz.base + lz.sd], Tz2 = T[lz.base + 2 * lz.sd];
  const double Sx = s_in(C, a0, Tx0, Tx1, Tx2), Sz = s_in(C, a2, Tz0, Tz1, Tz2);
  const double Sy = HY ? s_in(C, a1, Ty0, Ty1, Ty2) : 0.0;
  double ezl = 0.0, eyl = 0.0, exl = 0.0, exu = 0.0, eyu = 0.0, ezu = 0.0;
  if constexpr (PRO) {  // coarse corrections of the six line-neighbour cells (clamped like the y loads)
    const int cc = q.i + L.NX * (q.j + L.NY * q.kl);   // cell index in the level-0 arrays (ghost plane at -plane)
    ezl = cfac * ec[cc - (lz.cd > 0 ? L.plane : 0)]; ezu = cfac * ec[cc + (lz.cd + 1 < lz.nd ? L.plane : 0)];
    if constexpr (HY) { eyl = cfac * ec[cc - (ly.cd > 0 ? L.NX : 0)]; eyu = cfac * ec[cc + (ly.cd + 1 < ly.nd ? L.NX : 0)]; }
    exl = cfac * ec[cc - (lx.cd > 0 ? 1 : 0)]; exu = cfac * ec[cc + (lx.cd + 1 < lx.nd ? 1 : 0)];
  }
  double sum = 0.0;
  // lower cells, ascending cell index: z-, y-, x-
  nb_lower<PRO>(C, lz, a2, Tz0, y, q.pb, ezl, sum);
  if constexpr (HY) nb_lower<PRO>(C, ly, a1, Ty0, y, q.pb, eyl, sum);
  nb_lower<PRO>(C, lx, a0, Tx0, y, q.pb, exl, sum);
  // in-cell partners: smaller local index (flip a2, a1, a0 from 1 to 0), then larger (flip a0, a1, a2 from 0 to 1)
  sum += a2 ? yz * Sz : 0.0;
  if constexpr (HY) sum += a1 ? yy * Sy : 0.0;
  sum += a0 ? yx * Sx : 0.0;
  sum += a0 ? 0.0 : yx * Sx;
  if constexpr (HY) sum += a1 ? 0.0 : yy * Sy;
  sum += a2 ? 0.0 : yz * Sz;
  // upper cells: x+, y+, z+
  nb_upper<PRO>(C, lx, a0, Tx2, y, q.pb, exu, sum);
  if constexpr (HY) nb_upper<PRO>(C, ly, a1, Ty2, y, q.pb, eyu, sum);
  nb_upper<PRO>(C, lz, a2, Tz2, y, q.pb, ezu, sum);
  return sum;
}

// PRO: the first post-smoothing sweep with the prolongation fused in: yin is the iterate before
// y += cfac * ec(cell) (what k1_mg_prolong would have written, one whole pass over the level saved); a missing
// neighbour contributes (y + e) * (b (b' 0)) = +-0 as before.
// DOT: the last sweep of a V-cycle inside MINRES also finishes beta^2 = <v0, P v0>: its right-hand side rp IS the
// pressure part of v0 and its result the pressure part of P v0, so the products are formed here (canonical tree over
// the CTA's 256 dofs -> partial_p[cta]); the last CTA adds tree(partial_u) (velocity part, from k_minres_b_dot) and
// tree(partial_p) and writes the scalar.
template <bool HY, bool PRO = false, bool DOT = false>
__global__ void __launch_bounds__(BLK) k1_mg_jacobi(Lay1 L, Rt1Const C, const double* __restrict__ T,
                                                    const double* __restrict__ Dp, const double* __restrict__ rp,
                                                    const double* __restrict__ yin, double* __restrict__ yout,
                                                    double omega, int first, const MinresState* st,
                                                    const double* __restrict__ ec = nullptr, double cfac = 0.0,
                                                    double* partial_p = nullptr, const double* partial_u = nullptr,
                                                    unsigned n_u_chunks = 0, double* out = nullptr,
                                                    unsigned* ticket = nullptr) {
  if (st && st->done) return;
  const Dof1 d = owned_dof1<HY>(L);
  const int p = d.q.pb + d.l;
  if (first) {
    if (d.q.ok) yout[p] = omega * (rp[p] / Dp[p]);
    return;
  }
  double ys = yin[p];
  if constexpr (PRO) ys = ys + cfac * ec[d.q.i + L.NX * (d.q.j + L.NY * d.q.kl)];
  const double sum = q1_row_sum_dof<HY, PRO>(L, C, d, T, yin, ys, ec, cfac);
  const double dd = Dp[p];
  const double sy = dd * ys + sum;
  const double rv = rp[p];
  const double yn = ys + omega * ((rv - sy) / dd);
  if (d.q.ok) yout[p] = yn;
  if constexpr (DOT) {
    __shared__ double sm[BLK / 32];
    __shared__ int is_last;
    const double s = block_tree_256(d.q.ok ? rv * yn : 0.0, sm);
    if (threadIdx.x == 0) {
      partial_p[blockIdx.x] = s;
      __threadfence();
      is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      const double a = ordered_partial_sum(partial_u, n_u_chunks), b = ordered_partial_sum(partial_p, gridDim.x);
      __syncthreads();
      const double ta = block_tree_256(a, sm);
      __syncthreads();
      const double tb = block_tree_256(b, sm);
      if (threadIdx.x == 0) {
        *out = ta + tb;
        *ticket = 0u;
      }
    }
  }
}

// cell rhs = sum over the cell's dofs (local order) of the residual r - S y
template <bool HY>
__global__ void __launch_bounds__(BLK) k1_mg_restrict(Lay1 L, Rt1Const C, const double* __restrict__ T,
                                                      const double* __restrict__ Dp, const double* __restrict__ rp,
                                                      const double* __restrict__ y, double* __restrict__ rc,
                                                      const MinresState* st) {
  constexpr int NPL = HY ? 8 : 4;
  if (st && st->done) return;
  const Dof1 d = owned_dof1<HY>(L);
  const int p = d.q.pb + d.l;
  const double ys = y[p];
  const double sum = q1_row_sum_dof<HY>(L, C, d, T, y, ys);
  const double res = rp[p] - (Dp[p] * ys + sum);
  // ordered sum over the cell's lanes: ((0 + res_0) + res_1) + ...
  const int lane0 = (threadIdx.x & 31) & ~(NPL - 1);
  double acc = 0.0;
#pragma unroll
  for (int l = 0; l < NPL; ++l) acc += __shfl_sync(0xffffffffu, res, lane0 + l);
  if (d.q.ok && d.l == 0) rc[d.q.i + L.NX * (d.q.j + L.NY * d.q.kl)] = acc;
}

// cell-level transmissibilities below the Q1 level: Galerkin with cell-wise constant prolongation.  The two L2
// dofs of a cell on a line sum to -/+ wt on the line's face nodes and to 0 on its interior node, so the
// aggregated operator is two-point with T_face = sum over the face's RT1 nodes of wt^2 T_node.
template <bool HY>
__global__ void __launch_bounds__(BLK) k1_mg_copy_fine(Lay1 L, Rt1Const C, MgLevel v, const double* __restrict__ T) {
  constexpr int oy = HY ? 2 : 1;
  const Cell1 q = owned_cell1(L);
  if (!q.ok) return;
  const double w2 = C.wflux * C.wflux;
  const int i = q.i, j = q.j, k = q.kl, c = i + v.nx * (j + v.ny * k), pl = v.nx * v.ny;
  {
    double t = 0.0;
#pragma unroll
    for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
      for (int t1 = 0; t1 < oy; ++t1) t += w2 * T[line_x<HY>(L, q, t1, t2).base];
    v.Tx[i + (v.nx + 1) * (j + v.ny * k)] = t;
    if (i + 1 == v.nx) v.Tx[i + 1 + (v.nx + 1) * (j + v.ny * k)] = 0.0;
  }
  if constexpr (HY) {
    double t = 0.0;
#pragma unroll
    for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
      for (int t0 = 0; t0 < 2; ++t0) t += w2 * T[line_y(L, q, t0, t2).base];
    v.Ty[i + v.nx * (j + (v.ny + 1) * k)] = t;
    if (j + 1 == v.ny) v.Ty[i + v.nx * (j + 1 + (v.ny + 1) * k)] = 0.0;
  }
  {
    double t = 0.0;
#pragma unroll
    for (int t1 = 0; t1 < oy; ++t1)
#pragma unroll
      for (int t0 = 0; t0 < 2; ++t0) t += w2 * T[line_z<HY>(L, q, t0, t1).base];
    v.Tz[c] = t;
    if (k + 1 == v.nz) {  // slab top: the ghost-above face nodes (0 on the global boundary)
      double tt = 0.0;
      if (k + v.zoff + 1 < v.nzg) {
#pragma unroll
        for (int t1 = 0; t1 < oy; ++t1)
#pragma unroll
          for (int t0 = 0; t0 < 2; ++t0) {
            const Line1 l = line_z<HY>(L, q, t0, t1);
            tt += w2 * T[l.base + 2 * l.sd];
          }
      }
      v.Tz[c + pl] = tt;
    }
  }
}

}  // namespace impes
