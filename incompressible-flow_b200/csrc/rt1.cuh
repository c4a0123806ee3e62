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
__device__ __forceinline__ Cell1 owned_cell1_at(const Lay1& L, unsigned blk) {
  Cell1 q;
  const unsigned t = blk * (unsigned)BLK + threadIdx.x;
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
__device__ __forceinline__ Cell1 owned_cell1(const Lay1& L) { return owned_cell1_at(L, blockIdx.x); }

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
// Branch-free neighbour terms: T is stored as 0 on eliminated boundary nodes, so a missing neighbour contributes
// y * (b (b' 0)) = +-0 with y loaded from a clamped (own-cell) address.
//
// One thread per PAIR of L2 dofs (a0 = 0, 1 with the same (a1, a2): consecutive in memory): the two rows share the
// cell's index arithmetic, the x-line and every 16-byte load (their y / T operands are neighbours in memory), and are
// each other's x-partner, so the instruction count per row drops by ~40 % against the one-thread-per-dof form of round
// 1 (456 instructions per row, issue bound at 45 % of the HBM roofline).  Every row still adds the same terms in the
// same order (the ascending-column order of the assembled S): z-, y-, x- neighbour cells, in-cell partners with a
// smaller, then larger index, x+, y+, z+ neighbour cells.
constexpr int MGQ_T = 128;  // threads per CTA = 256 dofs (the canonical reduction block)
#ifndef IMPES_MGQ_MINB
#define IMPES_MGQ_MINB 4    // resident CTAs per SM of the plain sweep: the pair form needs ~108 registers; capped at
                            // 64 (8 CTAs) it spilled 112 B per thread and the spill traffic made l1tex the top unit
#endif
#ifndef IMPES_MGQ_MINB2
#define IMPES_MGQ_MINB2 4   // ... of the sweeps with fused prolongation / dot and of the restriction
#endif

struct double2v { double v0, v1; };
__device__ __forceinline__ double2v ld2(const double* p) {
  const double2 t = *reinterpret_cast<const double2*>(p);
  return {t.x, t.y};
}

struct Pair1 {
  Cell1 q;       // q.pb: index of the cell's first dof RELATIVE to the p-part
  int a1, a2, l0;  // transverse indices of the pair and local index of its first dof
};
template <bool HY>
__device__ __forceinline__ Pair1 owned_pair1(const Lay1& L, unsigned blk) {
  constexpr int NG = HY ? 4 : 2, NPL = HY ? 8 : 4;
  Pair1 d;
  const unsigned t = blk * (unsigned)MGQ_T + threadIdx.x;
  const unsigned cell = t / NG, g = t % NG;
  d.a1 = HY ? (int)(g & 1) : 0;
  d.a2 = HY ? (int)(g >> 1) : (int)g;
  d.l0 = HY ? 2 * d.a1 + 4 * d.a2 : 2 * d.a2;
  Cell1& q = d.q;
  q.ok = cell < (unsigned)L.ncl;
  const unsigned tt = q.ok ? cell : 0u;
  const unsigned r = fastdiv(tt, L.fdx);
  q.i = (int)(tt - r * (unsigned)L.NX);
  q.kl = (int)fastdiv(r, L.fdy);
  q.j = (int)(r - (unsigned)q.kl * (unsigned)L.NY);
  q.k = q.kl + L.zlo;
  q.c = q.i + L.NX * (q.j + L.NY * (q.kl + 1));
  q.pb = q.c * NPL;
  return d;
}

// two-dof contribution of one neighbour cell on a line: row a adds  yA * (ba (b0 T)) + yB * (ba (b1 T))
__device__ __forceinline__ void nb_term(double yA, double yB, double ba, double b0, double b1, double T, double& sum) {
  sum += yA * (ba * (b0 * T));
  sum += yB * (ba * (b1 * T));
}

// off-diagonal parts of the pair's two rows; ys0 / ys1 = y of the two dofs (after the fused prolongation if PRO);
// all 32 lanes must call.  PRO: see k1_mg_jacobi.
template <bool HY, bool PRO>
__device__ __forceinline__ void q1_rows_pair(const Lay1& L, const Rt1Const& C, const Pair1& d, const double* __restrict__ T,
                                             const double* __restrict__ y, double ys0, double ys1,
                                             const double* __restrict__ ec, double cfac, double& sum0, double& sum1) {
  constexpr int NPL = HY ? 8 : 4;
  const Cell1& q = d.q;
  const int a1 = d.a1, a2 = d.a2;
  const bool xlo = q.i > 0, xhi = q.i + 1 < L.NX, ylo = HY && q.j > 0, yhi = HY && q.j + 1 < L.NY;
  const bool zlo = q.k > 0, zhi = q.k + 1 < L.NZ;
  // ---- loads: T on the x-line (shared by both rows) and on the two y- / z-lines (one 16-byte load per node)
  const Line1 lx = line_x<HY>(L, q, a1, a2);
  const Line1 lz = line_z<HY>(L, q, 0, a1);
  const double Tx0 = T[lx.base], Tx1 = T[lx.base + 1], Tx2 = T[lx.base + 2];
  const double2v Tz0 = ld2(T + lz.base), Tz1 = ld2(T + lz.base + lz.sd), Tz2 = ld2(T + lz.base + 2 * lz.sd);
  double2v Ty0 = {0.0, 0.0}, Ty1 = {0.0, 0.0}, Ty2 = {0.0, 0.0};
  int pdn_y = 0;
  if constexpr (HY) {
    const Line1 ly = line_y(L, q, 0, a2);
    Ty0 = ld2(T + ly.base); Ty1 = ld2(T + ly.base + ly.sd); Ty2 = ld2(T + ly.base + 2 * ly.sd);
    pdn_y = ly.pdn;
  }
  // neighbour cells' dofs on the pair's lines (clamped to the own cell where there is no neighbour)
  const int pb = q.pb;
  const double2v xl = ld2(y + (pb - (xlo ? NPL : 0) + d.l0)), xu = ld2(y + (pb + (xhi ? NPL : 0) + d.l0));
  const int oz = HY ? 2 * a1 : 0, sz = HY ? 4 : 2;   // z-line dofs: local (a0 + oz) and (a0 + oz + sz)
  const double2v zlA = ld2(y + (pb - (zlo ? lz.pdn : 0) + oz)), zlB = ld2(y + (pb - (zlo ? lz.pdn : 0) + oz + sz));
  const double2v zuA = ld2(y + (pb + (zhi ? lz.pdn : 0) + oz)), zuB = ld2(y + (pb + (zhi ? lz.pdn : 0) + oz + sz));
  double2v ylA = {0.0, 0.0}, ylB = {0.0, 0.0}, yuA = {0.0, 0.0}, yuB = {0.0, 0.0};
  if constexpr (HY) {   // y-line dofs: local (a0 + 4 a2) and (a0 + 4 a2 + 2)
    ylA = ld2(y + (pb - (ylo ? pdn_y : 0) + 4 * a2)); ylB = ld2(y + (pb - (ylo ? pdn_y : 0) + 4 * a2 + 2));
    yuA = ld2(y + (pb + (yhi ? pdn_y : 0) + 4 * a2)); yuB = ld2(y + (pb + (yhi ? pdn_y : 0) + 4 * a2 + 2));
  }
  double ezl = 0.0, eyl = 0.0, exl = 0.0, exu = 0.0, eyu = 0.0, ezu = 0.0;
  if constexpr (PRO) {  // coarse corrections of the six line-neighbour cells (clamped like the y loads)
    const int cc = q.i + L.NX * (q.j + L.NY * q.kl);   // cell index in the level-0 arrays (ghost plane at -plane)
    ezl = cfac * ec[cc - (zlo ? L.plane : 0)]; ezu = cfac * ec[cc + (zhi ? L.plane : 0)];
    if constexpr (HY) { eyl = cfac * ec[cc - (ylo ? L.NX : 0)]; eyu = cfac * ec[cc + (yhi ? L.NX : 0)]; }
    exl = cfac * ec[cc - (xlo ? 1 : 0)]; exu = cfac * ec[cc + (xhi ? 1 : 0)];
  }
  // in-cell partners of the other lanes of the cell (same a0): a1 flipped = lane ^ 1, a2 flipped = lane ^ (HY ? 2 : 1)
  const double yy0 = HY ? __shfl_xor_sync(0xffffffffu, ys0, 1) : 0.0, yy1 = HY ? __shfl_xor_sync(0xffffffffu, ys1, 1) : 0.0;
  const double yz0 = __shfl_xor_sync(0xffffffffu, ys0, HY ? 2 : 1), yz1 = __shfl_xor_sync(0xffffffffu, ys1, HY ? 2 : 1);
  // ---- the two rows.  PRO adds the neighbour cell's correction to its dofs on the fly (same expression as the prolongation)
  const double b10 = BV(a1, 0), b12 = BV(a1, 2), b20 = BV(a2, 0), b22 = BV(a2, 2);
#define IMPES_Q1_ROW(A0, YS_OTHER, TZ0, TZ1, TZ2, TY0, TY1, TY2, ZLA, ZLB, YLA, YLB, XUA_, YUA, YUB, ZUA, ZUB, YY, YZ, SUM) \
  {                                                                                                                        \
    double sum = 0.0;                                                                                                      \
    /* lower cells, ascending cell index: z-, y-, x- */                                                                    \
    nb_term(PRO ? ZLA + ezl : ZLA, PRO ? ZLB + ezl : ZLB, b20, C.Bv[0][2], C.Bv[1][2], TZ0, sum);                          \
    if constexpr (HY) nb_term(PRO ? YLA + eyl : YLA, PRO ? YLB + eyl : YLB, b10, C.Bv[0][2], C.Bv[1][2], TY0, sum);        \
    nb_term(PRO ? xl.v0 + exl : xl.v0, PRO ? xl.v1 + exl : xl.v1, C.Bv[A0][0], C.Bv[0][2], C.Bv[1][2], Tx0, sum);          \
    /* in-cell partners: smaller local index first (a2, a1, a0 flipped from 1 to 0), then larger (a0, a1, a2 from 0 to 1) */ \
    const double Sz = s_in(C, a2, TZ0, TZ1, TZ2), Sx = s_in(C, A0, Tx0, Tx1, Tx2);                                         \
    const double Sy = HY ? s_in(C, a1, TY0, TY1, TY2) : 0.0;                                                               \
    if (a2) sum += YZ * Sz;                                                                                                \
    if constexpr (HY) { if (a1) sum += YY * Sy; }                                                                          \
    if (A0) sum += YS_OTHER * Sx;                                                                                          \
    if (!A0) sum += YS_OTHER * Sx;                                                                                         \
    if constexpr (HY) { if (!a1) sum += YY * Sy; }                                                                         \
    if (!a2) sum += YZ * Sz;                                                                                               \
    /* upper cells: x+, y+, z+ */                                                                                          \
    nb_term(PRO ? xu.v0 + exu : xu.v0, PRO ? xu.v1 + exu : xu.v1, C.Bv[A0][2], C.Bv[0][0], C.Bv[1][0], Tx2, sum);          \
    if constexpr (HY) nb_term(PRO ? YUA + eyu : YUA, PRO ? YUB + eyu : YUB, b12, C.Bv[0][0], C.Bv[1][0], TY2, sum);        \
    nb_term(PRO ? ZUA + ezu : ZUA, PRO ? ZUB + ezu : ZUB, b22, C.Bv[0][0], C.Bv[1][0], TZ2, sum);                          \
    SUM = sum;                                                                                                             \
  }
  IMPES_Q1_ROW(0, ys1, Tz0.v0, Tz1.v0, Tz2.v0, Ty0.v0, Ty1.v0, Ty2.v0, zlA.v0, zlB.v0, ylA.v0, ylB.v0, 0, yuA.v0, yuB.v0, zuA.v0, zuB.v0, yy0, yz0, sum0)
  IMPES_Q1_ROW(1, ys0, Tz0.v1, Tz1.v1, Tz2.v1, Ty0.v1, Ty1.v1, Ty2.v1, zlA.v1, zlB.v1, ylA.v1, ylB.v1, 0, yuA.v1, yuB.v1, zuA.v1, zuB.v1, yy1, yz1, sum1)
#undef IMPES_Q1_ROW
}

// One damped-Jacobi sweep on the Q1 level: first = 1: y_out = omega (r / D) (zero initial guess), else
// y_out = y_in + omega ((r - S y_in) / D).
// PRO: the first post-smoothing sweep with the prolongation fused in: yin is the iterate before
// y += cfac * ec(cell) (one whole pass over the level saved); a missing neighbour contributes
// (y + e) * (b (b' 0)) = +-0 as before.
// DOT: the last sweep of a V-cycle inside MINRES also finishes beta^2 = <v0, P v0>: its right-hand side rp IS the
// pressure part of v0 and its result the pressure part of P v0, so the products are formed here (canonical tree over
// the CTA's 256 dofs -> partial_p[cta]); the last CTA adds tree(partial_u) (velocity part, from k_minres_b_dot) and
// tree(partial_p) and writes the scalar.
template <bool HY, bool PRO = false, bool DOT = false>
__global__ void __launch_bounds__(MGQ_T, (PRO || DOT) ? IMPES_MGQ_MINB2 : IMPES_MGQ_MINB) k1_mg_jacobi(Lay1 L, Rt1Const C, const double* __restrict__ T,
                                                      const double* __restrict__ Dp, const double* __restrict__ rp,
                                                      const double* __restrict__ yin, double* __restrict__ yout,
                                                      double omega, int first, const MinresState* st, unsigned nblk,
                                                      const double* __restrict__ ec = nullptr, double cfac = 0.0,
                                                      double* partial_p = nullptr, const double* partial_u = nullptr,
                                                      unsigned n_u_chunks = 0, double* out = nullptr,
                                                      unsigned* ticket = nullptr) {
  if (st && st->done) return;
  // DOT: nblk 256-dof blocks are walked by a grid of resident CTAs (block b -> partial_p[b]) so that the fence +
  // ticket of the final reduction is paid once per CTA, not once per block (ncu, one CTA per block: barrier 2.9 and
  // membar 0.5 stalls per issue, 365 us against 215 us for the plain sweep at 128^3); otherwise one block per CTA.
  for (unsigned blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const Pair1 d = owned_pair1<HY>(L, blk);
    const int p = d.q.pb + d.l0;
    const double2v rv = ld2(rp + p), dd = ld2(Dp + p);
    if (first) {   // kept for 2-D / callers without the streaming kernel k1_mg_first
      if (d.q.ok) {
        double2 o;
        o.x = omega * div_rcp(rv.v0, dd.v0);   // same bits as `/`, half the instructions
        o.y = omega * div_rcp(rv.v1, dd.v1);
        *reinterpret_cast<double2*>(yout + p) = o;
      }
      continue;
    }
    double2v ys = ld2(yin + p);
    if constexpr (PRO) {
      const double e = cfac * ec[d.q.i + L.NX * (d.q.j + L.NY * d.q.kl)];
      ys.v0 = ys.v0 + e; ys.v1 = ys.v1 + e;
    }
    double sum0, sum1;
    q1_rows_pair<HY, PRO>(L, C, d, T, yin, ys.v0, ys.v1, ec, cfac, sum0, sum1);
    const double sy0 = dd.v0 * ys.v0 + sum0, sy1 = dd.v1 * ys.v1 + sum1;
    const double yn0 = ys.v0 + omega * div_rcp(rv.v0 - sy0, dd.v0), yn1 = ys.v1 + omega * div_rcp(rv.v1 - sy1, dd.v1);
    if (d.q.ok) {
      double2 o;
      o.x = yn0; o.y = yn1;
      *reinterpret_cast<double2*>(yout + p) = o;
    }
    if constexpr (DOT) {
      __shared__ double smb[8];
      const double s = block_tree_256_pairs(d.q.ok ? rv.v0 * yn0 : 0.0, d.q.ok ? rv.v1 * yn1 : 0.0, smb);
      if (threadIdx.x == 0) partial_p[blk] = s;
      __syncthreads();   // smb is reused by the next block
    }
  }
  if constexpr (DOT) {
    __shared__ double sm[8];
    __shared__ int is_last;
    if (threadIdx.x == 0) {
      __threadfence();
      is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      // thread t of the canonical 256-thread reduction block = value slots t (x0) and t + 128 (x1) here:
      // ordered sums of entries t, t + 256, ... like ordered_partial_sum, two slots per thread
      const double a0 = ordered_partial_sum_at(partial_u, n_u_chunks, threadIdx.x), a1 = ordered_partial_sum_at(partial_u, n_u_chunks, threadIdx.x + 128);
      const double b0 = ordered_partial_sum_at(partial_p, nblk, threadIdx.x), b1 = ordered_partial_sum_at(partial_p, nblk, threadIdx.x + 128);
      __syncthreads();
      const double ta = block_tree_256_split(a0, a1, sm);
      __syncthreads();
      const double tb = block_tree_256_split(b0, b1, sm);
      if (threadIdx.x == 0) {
        *out = ta + tb;
        *ticket = 0u;
      }
    }
  }
}

// first sweep from a zero initial guess: y = omega (r / D) over the owned L2 dofs, a contiguous range of the p-part
// (the generic sweep kernel spends its few resident warps on index arithmetic: 0.51 of the HBM roofline for 3 streams)
__global__ void __launch_bounds__(BLK) k1_mg_first(long n2, const double2* __restrict__ D, const double2* __restrict__ r,
                                                   double2* __restrict__ y, double omega, const MinresState* st) {
  if (st && st->done) return;
  for (long i = (long)blockIdx.x * BLK + threadIdx.x; i < n2; i += (long)gridDim.x * BLK) {
    const double2 rv = r[i], dd = D[i];
    double2 o;
    o.x = omega * div_rcp(rv.x, dd.x);
    o.y = omega * div_rcp(rv.y, dd.y);
    y[i] = o;
  }
}

// cell rhs = sum over the cell's dofs (local order) of the residual r - S y
template <bool HY>
__global__ void __launch_bounds__(MGQ_T, IMPES_MGQ_MINB2) k1_mg_restrict(Lay1 L, Rt1Const C, const double* __restrict__ T,
                                                        const double* __restrict__ Dp, const double* __restrict__ rp,
                                                        const double* __restrict__ y, double* __restrict__ rc,
                                                        const MinresState* st) {
  constexpr int NG = HY ? 4 : 2;
  if (st && st->done) return;
  const Pair1 d = owned_pair1<HY>(L, blockIdx.x);
  const int p = d.q.pb + d.l0;
  const double2v ys = ld2(y + p), rv = ld2(rp + p), dd = ld2(Dp + p);
  double sum0, sum1;
  q1_rows_pair<HY, false>(L, C, d, T, y, ys.v0, ys.v1, nullptr, 0.0, sum0, sum1);
  const double res0 = rv.v0 - (dd.v0 * ys.v0 + sum0), res1 = rv.v1 - (dd.v1 * ys.v1 + sum1);
  // ordered sum over the cell's dofs: ((0 + res_0) + res_1) + ...   (dof 2g, 2g+1 in lane g of the cell's group)
  const int lane0 = (threadIdx.x & 31) & ~(NG - 1);
  double acc = 0.0;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    acc += __shfl_sync(0xffffffffu, res0, lane0 + g);
    acc += __shfl_sync(0xffffffffu, res1, lane0 + g);
  }
  if (d.q.ok && (threadIdx.x & (NG - 1)) == 0) rc[d.q.i + L.NX * (d.q.j + L.NY * d.q.kl)] = acc;
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
