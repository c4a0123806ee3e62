// kernels.cuh -- hand-written sm_100a kernels of the IMPES hot path (RT0/P0/P0).
//
// Arithmetic contract (DESIGN.md "bit-reproducible arithmetic"): this file is compiled with
// -fmad=false and every expression below is written in the exact association order of the
// CPU oracle's assembled-CSR algorithm, so that a full MINRES solve / saturation step is
// bit-identical to the oracle, not merely close: at rtol 1e-6 a Krylov solve amplifies
// 1-ulp differences to 1e-8..1e-7 (tests/test_oracle_invariants.py::test_krylov_drift),
// so "1e-10 per step" (BASELINE north_star) is only reachable this way.
// All reductions use one fixed tree (block_tree_256 + ordered partials), never atomics on
// doubles.  The kernels are HBM/L2-bandwidth or latency bound; giving up FMA contraction
// costs nothing measurable.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "layout.h"

namespace impes {
namespace cg = cooperative_groups;

constexpr int BLK = 256;  // canonical block: 256 consecutive owned cells
#define THIRD (1.0 / 3.0)
#define SIXTH (1.0 / 6.0)

// ---------------------------------------------------------------------------------------
// canonical reduction tree
// ---------------------------------------------------------------------------------------

// xor-butterfly inside each warp (16,8,4,2,1), then the 8 warp sums added in order by thread 0.
// Result valid in thread 0.  sm: 8 doubles.
__device__ __forceinline__ double block_tree_256(double v, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < BLK / 32; ++w) s += sm[w];
  }
  return s;
}

// Ordered sum of entries t, t + 256, ... of p[0, n) by thread t: the loads of 8 terms are issued together (the tail of
// a dot over 65k+ block partials was a chain of dependent L2 round trips), the additions stay sequential.
__device__ __forceinline__ double ordered_partial_sum(const double* p, unsigned n) {
  double acc = 0.0;
  unsigned j = threadIdx.x;
  for (; j + 7 * 256 < n; j += 8 * 256) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(p + j + u * 256);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
  for (; j < n; j += 256) acc += __ldcg(p + j);
  return acc;
}

// The same for an explicit slot t of the canonical 256-thread reduction block (CTAs of another size)
__device__ __forceinline__ double ordered_partial_sum_at(const double* p, unsigned n, unsigned t) {
  double acc = 0.0;
  unsigned j = t;
  for (; j + 7 * 256 < n; j += 8 * 256) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(p + j + u * 256);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
  for (; j < n; j += 256) acc += __ldcg(p + j);
  return acc;
}
// block_tree_256 over 256 values held by a 128-thread CTA as slots t (x0) and t + 128 (x1): warp w of the canonical
// block = slots 32w .. 32w+31 = threads 32 (w & 3) .. +31 with x0 (w < 4) or x1 (w >= 4).  Result valid in thread 0.
__device__ __forceinline__ double block_tree_256_split(double x0, double x1, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x0 += __shfl_xor_sync(0xffffffffu, x0, o);
    x1 += __shfl_xor_sync(0xffffffffu, x1, o);
  }
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = x0; sm[4 + (threadIdx.x >> 5)] = x1; }
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[w];
  }
  return s;
}

// block_tree_256 over 256 values held two per thread by a 128-thread CTA (value 2t, 2t+1 in thread t): the same
// butterfly (xor 16, 8, 4, 2, 1 over the VALUE index inside groups of 32) and the same ordered sum of the 8 group
// results, hence the same bits.  Result valid in thread 0.  sm: 8 doubles.
__device__ __forceinline__ double block_tree_256_pairs(double x0, double x1, double* sm) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {   // value-index xor 16, 8, 4, 2 = lane xor 8, 4, 2, 1
    x0 += __shfl_xor_sync(0xffffffffu, x0, o);
    x1 += __shfl_xor_sync(0xffffffffu, x1, o);
  }
  const double v = x0 + x1;           // value-index xor 1
  if ((threadIdx.x & 15) == 0) sm[threadIdx.x >> 4] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[w];
  }
  return s;
}

// Block partial -> partial[blockIdx.x]; the last block to arrive adds all partials in index
// order (thread t takes t, t+256, ...; then the same tree) and writes *out.  Deterministic:
// which block is last does not matter.
__device__ __forceinline__ void finish_dot(double v, double* partial, double* out, unsigned* ticket) {
  __shared__ double sm[BLK / 32];
  __shared__ int is_last;
  const double s = block_tree_256(v, sm);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s;
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    const double a = ordered_partial_sum(partial, gridDim.x);
    __syncthreads();
    const double tot = block_tree_256(a, sm);
    if (threadIdx.x == 0) {
      *out = tot;
      *ticket = 0u;
    }
  }
}

// n / D, correctly rounded, from the correctly rounded reciprocal rD = RN(1/D): two FMA remainder
// corrections (Markstein): q0 = n*rD is within 2 ulp, q1 = q0 + (n - q0 D) rD is faithful, and the second
// correction of a faithful quotient with an exact remainder rounds correctly.  Bit-identical to IEEE `/`
// for finite, non-denormal operands, with a 5-op dependent chain instead of the ~30-instruction division.
__device__ __forceinline__ double div_exact(double n, double D, double rD) {
  double q = n * rD;
  double r = __fma_rn(-q, D, n);
  q = __fma_rn(r, rD, q);
  r = __fma_rn(-q, D, n);
  return __fma_rn(r, rD, q);
}

// n / D through the correctly rounded reciprocal intrinsic + the two remainder corrections above: the same bits
// as IEEE `/` (finite, non-denormal operands) in roughly half the instructions of the division subroutine.
__device__ __forceinline__ double div_rcp(double n, double D) { return div_exact(n, D, __drcp_rn(D)); }

// ---------------------------------------------------------------------------------------
// physics: Param::Krw / Param::Kro  [REF src/param.cpp:195-215]  (no clamping).
// Divisions by the run-time constants (1 - Sor - Swc), mu_w, mu_o go through div_exact with host-computed
// reciprocals: same bits as `/`, a fifth of the instructions.
// ---------------------------------------------------------------------------------------
// (S - Swc)/(1 - Sor - Swc); with Swc = 0 and denominator 1 both operations are exact identities and are skipped
__device__ __forceinline__ double s_eff(const Phys& ph, double S) {
  return ph.unit_den ? S : div_exact(S - ph.Swc, ph.den, ph.rden);
}
__device__ __forceinline__ double krw_of(const Phys& ph, double S) {
  if (ph.two_phase) {
    const double S_ = s_eff(ph, S);
    return S_ * S_;
  }
  return S;
}
__device__ __forceinline__ double kro_of(const Phys& ph, double S) {
  if (ph.two_phase) {
    const double S_ = s_eff(ph, S);
    return (1. - S_) * (1. - S_);
  }
  return (1. - S);
}
// phase mobilities Krw/mu_w, Kro/mu_o (x/1.0 is x exactly: skipped for the reference's mu = 1)
__device__ __forceinline__ void mobilities(const Phys& ph, double S, double& mw, double& mo) {
  const double kw = krw_of(ph, S), ko = kro_of(ph, S);
  if (ph.unit_mu) { mw = kw; mo = ko; }
  else { mw = div_exact(kw, ph.mu_w, ph.rmu_w); mo = div_exact(ko, ph.mu_o, ph.rmu_o); }
}
// total mobility lambda = Krw/mu_w + Kro/mu_o   (CWVectorCoefficient::Eval [REF src/param.hpp:252-254])
__device__ __forceinline__ double lambda_of(const Phys& ph, double S) {
  double mw, mo;
  mobilities(ph, S, mw, mo);
  return mw + mo;
}
// FS  [REF src/saturation_solver.cpp:8-17]
__device__ __forceinline__ double frac_flow_of(const Phys& ph, double S) {
  double mw, mo;
  mobilities(ph, S, mw, mo);
  return mw / (mw + mo);
}

struct CellIdx {
  int i, j, kl, k;  // kl = local plane, k = global plane
  bool ok;
};
__device__ __forceinline__ CellIdx owned_cell(const Lay& L) {
  CellIdx c;
  const unsigned t = blockIdx.x * (unsigned)BLK + threadIdx.x;  // owned cells per rank < 2^32 (checked at create)
  c.ok = t < (unsigned)L.ncl;
  const unsigned tt = c.ok ? t : 0u;
  const unsigned r = fastdiv(tt, L.fdx);
  c.i = (int)(tt - r * (unsigned)L.NX);
  c.kl = (int)fastdiv(r, L.fdy);
  c.j = (int)(r - (unsigned)c.kl * (unsigned)L.NY);
  c.k = c.kl + L.zlo;
  return c;
}

// ---------------------------------------------------------------------------------------
// K4: S-dependent coefficients (once per pressure solve)
// ---------------------------------------------------------------------------------------

// m_d[c] = geo_d / (lambda(S_c) K_d[c])  -- CWVectorCoefficient::Eval [REF src/param.hpp:246-261]
// times the RT0 Piola factor h_d^2/|cell|.  Runs over the cvec incl. ghost planes.
__global__ void __launch_bounds__(BLK) k_cell_coeff(Lay L, Phys ph, const double* __restrict__ S,
                                                    const double* __restrict__ Kx, const double* __restrict__ Ky,
                                                    const double* __restrict__ Kz, double gx, double gy, double gz,
                                                    double* __restrict__ mx, double* __restrict__ my,
                                                    double* __restrict__ mz) {
  const long c = (long)blockIdx.x * BLK + threadIdx.x;
  if (c >= L.ncv) return;
  const int kk = (int)(c / L.plane);
  const int k = kk - 1 + L.zlo;
  if (k < 0 || k >= L.NZ) {
    mx[c] = 0.0; mz[c] = 0.0;
    if (L.has_y) my[c] = 0.0;
    return;
  }
  const double s = S[c];
  const double lam = lambda_of(ph, s);
  mx[c] = (1. / (lam * Kx[c])) * gx;
  if (L.has_y) my[c] = (1. / (lam * Ky[c])) * gy;
  mz[c] = (1. / (lam * Kz[c])) * gz;
}

// diag(M) on owned faces (1 on boundary faces: eliminated rows [REF src/pressure_solver.cpp:56-60]),
// T_f = 1/diag(M)_f on interior faces (0 on boundary faces), and the diagonal of
// S = B diag(M)^-1 B^T  [REF src/pressure_solver.cpp:76-81] stored in the p-part of Md.
// VAR (S in DG-Q1, k0q_coeff): mx / my / mz point at three arrays each, [a00 | a01 | a11] with stride ncv -- the entries of
// the 2x2 element block of the direction; diag(M)_f = a11(lower cell) + a00(upper cell).
template <bool VAR>
__global__ void __launch_bounds__(BLK) k_face_coeff(Lay L, const double* __restrict__ mx, const double* __restrict__ my,
                                                    const double* __restrict__ mz, double* __restrict__ Md,
                                                    double* __restrict__ T) {
  const CellIdx ci = owned_cell(L);
  if (!ci.ok) return;
  const int i = ci.i, j = ci.j, kl = ci.kl, k = ci.k, kk = kl + 1;
  const long c = ix_c(L, i, j, kk);
  const long n2 = 2L * L.ncv;   // VAR: offset of a11
  double D = 0.0;
  {  // x
    const long f = ix_ux(L, i, j, kl);
    if (i > 0) {
      const double md = VAR ? mx[n2 + c - 1] + mx[c] : mx[c - 1] * THIRD + mx[c] * THIRD;
      const double t = 1.0 / md;
      Md[f] = md; T[f] = t; D += t;
    } else { Md[f] = 1.0; T[f] = 0.0; }
    if (i < L.NX - 1) {
      const double md = VAR ? mx[n2 + c] + mx[c + 1] : mx[c] * THIRD + mx[c + 1] * THIRD;
      D += 1.0 / md;
    } else { Md[f + 1] = 1.0; T[f + 1] = 0.0; }
  }
  if (L.has_y) {
    const long f = ix_uy(L, i, j, kl);
    if (j > 0) {
      const double md = VAR ? my[n2 + c - L.NX] + my[c] : my[c - L.NX] * THIRD + my[c] * THIRD;
      const double t = 1.0 / md;
      Md[f] = md; T[f] = t; D += t;
    } else { Md[f] = 1.0; T[f] = 0.0; }
    if (j < L.NY - 1) {
      const double md = VAR ? my[n2 + c] + my[c + L.NX] : my[c] * THIRD + my[c + L.NX] * THIRD;
      D += 1.0 / md;
    } else { Md[f + L.NX] = 1.0; T[f + L.NX] = 0.0; }
  }
  {  // z
    const long f = ix_uz(L, i, j, kk);
    if (k > 0) {
      const double md = VAR ? mz[n2 + c - L.plane] + mz[c] : mz[c - L.plane] * THIRD + mz[c] * THIRD;
      const double t = 1.0 / md;
      Md[f] = md; T[f] = t; D += t;
    } else { Md[f] = 1.0; T[f] = 0.0; }
    if (k < L.NZ - 1) {
      const double md = VAR ? mz[n2 + c] + mz[c + L.plane] : mz[c] * THIRD + mz[c + L.plane] * THIRD;
      const double t = 1.0 / md;
      D += t;
      if (k == L.zhi - 1) T[f + L.plane] = t;  // ghost-above face plane: needed by the Schur sweeps
    } else { Md[f + L.plane] = 1.0; T[f + L.plane] = 0.0; }
  }
  Md[ix_p(L, i, j, kk)] = D;
}

// ---------------------------------------------------------------------------------------
// K1: y = [M(S) B^T; B 0] x, matrix-free, essential-BC rows/cols eliminated
//     [REF src/pressure_solver.cpp:47-74]; optional fused dot <x, y> (MINRES alpha).
//     One thread per owned cell; the cell computes the rows of the faces it owns (lower
//     faces; upper faces only on the domain boundary) and its pressure row: a gather
//     formulation -- no atomics, fixed summation order (the CSR column order of the oracle).
// ---------------------------------------------------------------------------------------
// One direction of the RT0 apply for one cell, branch-free: the loads were issued by the caller (clamped where the
// neighbour does not exist), skipped CSR entries become `+ 0.0` (bit-identical partial sums).
//   lo / lo2: the lower face is interior / the lower cell's lower face is interior; hi: the upper face is interior
//   cm / cd / cu: the row's coefficients of the lower face of the lower cell, the face itself, the upper face
//   (constant coefficient in the cell: (mL/6), (mL/3 + mR/3), (mR/6); DG-Q1 coefficient: a01(L), a11(L) + a00(R), a01(R))
template <bool DOT>
__device__ __forceinline__ void rt0_dir(bool lo, bool lo2, bool hi, double cm, double cd, double cu, double xm, double xf,
                                        double xf1, double pc, double pm, double& yf, double& yb, double& yp,
                                        double& acc) {
  double s = 0.0;
  s += lo2 ? cm * xm : 0.0;
  s += cd * xf;
  s += hi ? cu * xf1 : 0.0;
  yf = lo ? s + (pc - pm) : xf;  // eliminated boundary row: identity
  if (DOT) acc += xf * yf;
  yp += lo ? xf : 0.0;
  yb = xf1;  // identity row of the upper boundary face (stored and counted only when !hi)
  if (!hi) {
    if (DOT) acc += xf1 * xf1;
  } else {
    yp -= xf1;
  }
}

// VAR: see k_face_coeff
template <bool DOT, bool VAR = false>
__global__ void __launch_bounds__(BLK) k_darcy_apply(Lay L, const double* __restrict__ mx,
                                                     const double* __restrict__ my, const double* __restrict__ mz,
                                                     const double* __restrict__ x, double* __restrict__ y,
                                                     double* partial, double* out, unsigned* ticket,
                                                     const MinresState* st) {
  if (st && st->done) return;
  const CellIdx ci = owned_cell(L);
  double acc = 0.0;
  if (ci.ok) {
    const int i = ci.i, j = ci.j, kl = ci.kl, k = ci.k, kk = kl + 1;
    const int c = ix_c(L, i, j, kk);
    const int ip = L.off_p + c;
    const bool xl = i > 0, xl2 = i > 1, xh = i + 1 < L.NX;
    const bool yl = j > 0, yl2 = j > 1, yh = j + 1 < L.NY;
    const bool zl = k > 0, zl2 = k > 1, zh = k + 1 < L.NZ;
    const int fx = ix_ux(L, i, j, kl), fz = ix_uz(L, i, j, kk);
    const int fy = L.has_y ? ix_uy(L, i, j, kl) : fx;
    // ---- load phase (all unconditional)
    const double pc = x[ip];
    const double pxm = x[xl ? ip - 1 : ip], pym = x[(L.has_y && yl) ? ip - L.NX : ip], pzm = x[zl ? ip - L.plane : ip];
    const double xxf = x[fx], xxf1 = x[fx + 1], xxm = x[xl ? fx - 1 : fx];
    const double xyf = x[fy], xyf1 = x[L.has_y ? fy + L.NX : fy], xym = x[(L.has_y && yl) ? fy - L.NX : fy];
    const double xzf = x[fz], xzf1 = x[fz + L.plane], xzm = x[zl ? fz - L.plane : fz];
    // row coefficients (lower neighbour's face, the face, upper face) per direction
    double cxm, cxd, cxu, czm, czd, czu, cym = 0.0, cyd = 0.0, cyu = 0.0;
    if constexpr (VAR) {
      const long n1 = L.ncv, n2 = 2L * L.ncv;
      const int cxl = xl ? c - 1 : c, czl = zl ? c - L.plane : c;
      cxm = mx[n1 + cxl]; cxd = mx[n2 + cxl] + mx[c]; cxu = mx[n1 + c];
      czm = mz[n1 + czl]; czd = mz[n2 + czl] + mz[c]; czu = mz[n1 + c];
      if (L.has_y) {
        const int cyl = yl ? c - L.NX : c;
        cym = my[n1 + cyl]; cyd = my[n2 + cyl] + my[c]; cyu = my[n1 + c];
      }
    } else {
      const double mxR = mx[c], mxL = mx[xl ? c - 1 : c];
      const double mzR = mz[c], mzL = mz[zl ? c - L.plane : c];
      cxm = mxL * SIXTH; cxd = mxL * THIRD + mxR * THIRD; cxu = mxR * SIXTH;
      czm = mzL * SIXTH; czd = mzL * THIRD + mzR * THIRD; czu = mzR * SIXTH;
      if (L.has_y) {
        const double myR = my[c], myL = my[yl ? c - L.NX : c];
        cym = myL * SIXTH; cyd = myL * THIRD + myR * THIRD; cyu = myR * SIXTH;
      }
    }
    // ---- compute phase
    double yp = 0.0, yf, yb;
    rt0_dir<DOT>(xl, xl2, xh, cxm, cxd, cxu, xxm, xxf, xxf1, pc, pxm, yf, yb, yp, acc);
    y[fx] = yf;
    if (!xh) y[fx + 1] = yb;
    if (L.has_y) {
      rt0_dir<DOT>(yl, yl2, yh, cym, cyd, cyu, xym, xyf, xyf1, pc, pym, yf, yb, yp, acc);
      y[fy] = yf;
      if (!yh) y[fy + L.NX] = yb;
    }
    rt0_dir<DOT>(zl, zl2, zh, czm, czd, czu, xzm, xzf, xzf1, pc, pzm, yf, yb, yp, acc);
    y[fz] = yf;
    if (!zh) y[fz + L.plane] = yb;
    y[ip] = yp;
    if (DOT) acc += pc * yp;
  }
  if (DOT) finish_dot(acc, partial, out, ticket);
}

// Canonical dot <a, b> over the dofs each cell owns, same ownership order as k_darcy_apply.
__global__ void __launch_bounds__(BLK) k_dot_owned(Lay L, const double* __restrict__ a, const double* __restrict__ b,
                                                   double* partial, double* out, unsigned* ticket,
                                                   const MinresState* st) {
  if (st && st->done) return;
  const CellIdx ci = owned_cell(L);
  double acc = 0.0;
  if (ci.ok) {
    const int i = ci.i, j = ci.j, kl = ci.kl, k = ci.k, kk = kl + 1;
    long f = ix_ux(L, i, j, kl);
    acc += a[f] * b[f];
    if (i + 1 == L.NX) acc += a[f + 1] * b[f + 1];
    if (L.has_y) {
      f = ix_uy(L, i, j, kl);
      acc += a[f] * b[f];
      if (j + 1 == L.NY) acc += a[f + L.NX] * b[f + L.NX];
    }
    f = ix_uz(L, i, j, kk);
    acc += a[f] * b[f];
    if (k + 1 == L.NZ) acc += a[f + L.plane] * b[f + L.plane];
    f = ix_p(L, i, j, kk);
    acc += a[f] * b[f];
  }
  finish_dot(acc, partial, out, ticket);
}

// ---------------------------------------------------------------------------------------
// K2: fused MINRES vector passes (MINRESSolver::Mult [MFEM-recall])
// ---------------------------------------------------------------------------------------

// pass B:  v0 <- (q - beta v0) - alpha v1 ;  q_u <- v0_u / diag(M)   (DSmoother, Jacobi block)
__global__ void __launch_bounds__(BLK) k_minres_b(long n, long n_u, const double* q_in,
                                                  double* __restrict__ v0, const double* __restrict__ v1,
                                                  const double* __restrict__ Md, double* q_out,
                                                  const double* alpha_p, const MinresState* st) {
  if (st->done) return;
  const double alpha = *alpha_p, beta = st->beta, m0 = st->m0, m1 = st->m1;
  // two entries per thread and trip (n and n_u are multiples of 32)
  for (long i = 2 * ((long)blockIdx.x * BLK + threadIdx.x); i < n; i += 2 * (long)gridDim.x * BLK) {
    const double2 qi = *reinterpret_cast<const double2*>(q_in + i);
    const double2 a = *reinterpret_cast<const double2*>(v0 + i), b = *reinterpret_cast<const double2*>(v1 + i);
    double2 v;
    const double t0 = qi.x + (-beta) * (a.x * m0), t1 = qi.y + (-beta) * (a.y * m0);
    v.x = t0 + (-alpha) * (b.x * m1);
    v.y = t1 + (-alpha) * (b.y * m1);
    *reinterpret_cast<double2*>(v0 + i) = v;
    if (i < n_u) {
      const double2 md = *reinterpret_cast<const double2*>(Md + i);
      double2 qv;
      qv.x = v.x / md.x; qv.y = v.y / md.y;
      *reinterpret_cast<double2*>(q_out + i) = qv;
    }
  }
}

// The same pass with the velocity part of beta^2 = <v0, P v0> folded in (P_u = diag(M)^-1 is applied right here, so
// v0_u and q_u never have to be read again by a dot kernel: 3/4 of that kernel's traffic).  One CTA per chunk of
// DOT_CHUNK consecutive entries of the vector (a mapping that does not depend on the SM count, so the CPU checker can
// follow it): thread t adds v0 q of entries t, t + 256, ... of its chunk in order, the CTA reduces with the canonical
// tree and stores partial_u[chunk]; the pressure part and the final sum come from the last smoothing sweep of the
// Schur preconditioner (k1_mg_jacobi<.., DOT>).  Entries in [x0, x1) u [x2, x3) are ghost layers (stale on several
// GPUs) and padding; they are updated like everything else but left out of the sum.
constexpr int DOT_CHUNK = 16384;
constexpr int MINRES_B_T = 128;   // threads of a pass-B CTA: thread t owns slots 2t, 2t + 1 of the canonical 256-slot block,
                                  // i.e. entries 2t + 256 k (+1) of its chunk -- one 16-byte access per vector and trip
                                  // (the 8-byte form ran at 0.79 of the HBM roofline; pass C gained 11 % the same way)
__global__ void __launch_bounds__(MINRES_B_T) k_minres_b_dot(long n, long n_u, const double* q_in,
                                                             double* __restrict__ v0, const double* __restrict__ v1,
                                                             const double* __restrict__ Md, double* q_out,
                                                             const double* alpha_p, const MinresState* st,
                                                             long x0, long x1, long x2, long x3, double* __restrict__ partial_u) {
  if (st->done) return;
  const double alpha = *alpha_p, beta = st->beta, m0 = st->m0, m1 = st->m1;
  const long base = (long)blockIdx.x * DOT_CHUNK;
  double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 4
  for (int k = 0; k < DOT_CHUNK / BLK; ++k) {
    const long i = base + 2 * threadIdx.x + k * BLK;   // even; n and n_u are multiples of 32
    if (i < n) {
      const double2 qi = *reinterpret_cast<const double2*>(q_in + i);
      const double2 a = *reinterpret_cast<const double2*>(v0 + i), b = *reinterpret_cast<const double2*>(v1 + i);
      double2 v;
      const double t0 = qi.x + (-beta) * (a.x * m0), t1 = qi.y + (-beta) * (a.y * m0);
      v.x = t0 + (-alpha) * (b.x * m1);
      v.y = t1 + (-alpha) * (b.y * m1);
      *reinterpret_cast<double2*>(v0 + i) = v;
      if (i < n_u) {
        const double2 md = *reinterpret_cast<const double2*>(Md + i);
        double2 qv;
        qv.x = v.x / md.x; qv.y = v.y / md.y;
        *reinterpret_cast<double2*>(q_out + i) = qv;
        const bool g0 = (i >= x0 && i < x1) || (i >= x2 && i < x3);
        const bool g1 = (i + 1 >= x0 && i + 1 < x1) || (i + 1 >= x2 && i + 1 < x3);
        acc0 += g0 ? 0.0 : v.x * qv.x;
        acc1 += g1 ? 0.0 : v.y * qv.y;
      }
    }
  }
  __shared__ double sm[8];
  const double s = block_tree_256_pairs(acc0, acc1, sm);
  if (threadIdx.x == 0) partial_u[blockIdx.x] = s;
}

// pass C: scalar recurrence + w/x update + scaling of the next Lanczos pair.
//   w0 <- (-rho3/rho1) w0 + (-rho2/rho1) w1 + (1/rho1) u1 ; x += gamma1 eta w0 ;
//   q *= 1/beta_new (becomes u1 of the next iteration: K1 and the halo exchange read it as is);
//   the scaling v0 *= 1/beta_new (becomes v1) is lazy: only the factor is recorded (MinresState::m1)
template <bool LAZYU>
__global__ void __launch_bounds__(BLK) k_minres_c(long n, double* __restrict__ w0, const double* __restrict__ w1,
                                                  const double* __restrict__ u1, double* __restrict__ x,
                                                  double* __restrict__ v0, double* __restrict__ q,
                                                  const double* alpha_p, const double* bb_p, const MinresState* st,
                                                  MinresState* st_next, int maxiter) {
  if (st->done) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *st_next = *st;
    return;
  }
  const double alpha = *alpha_p, beta_old = st->beta, eta = st->eta;
  const double gamma0 = st->gamma0, gamma1 = st->gamma1, sigma0 = st->sigma0, sigma1 = st->sigma1;
  const double delta = gamma1 * alpha - gamma0 * sigma1 * beta_old;
  const double rho3 = sigma0 * beta_old;
  const double rho2 = sigma1 * alpha + gamma0 * gamma1 * beta_old;
  const double beta = sqrt(*bb_p);
  const double rho1 = sqrt(delta * delta + beta * beta);  // hypot(delta, beta)
  const double a = -rho3 / rho1, c = -rho2 / rho1, e = 1. / rho1;
  const double gamma1n = delta / rho1;
  const double xf = gamma1n * eta;
  const double sigma1n = beta / rho1;
  const double etan = -sigma1n * eta;
  const double m = 1.0 / beta;
  const double mu = st->mu;
  // two entries per thread and trip (16-byte accesses; n is a multiple of 32, the vectors are 256-byte aligned)
  const long n2 = n >> 1;
  double2* w0v = reinterpret_cast<double2*>(w0);
  const double2* w1v = reinterpret_cast<const double2*>(w1);
  const double2* u1v = reinterpret_cast<const double2*>(u1);
  double2* xv = reinterpret_cast<double2*>(x);
  double2* qv = reinterpret_cast<double2*>(q);
#pragma unroll 2
  for (long i = (long)blockIdx.x * BLK + threadIdx.x; i < n2; i += (long)gridDim.x * BLK) {
    const double2 a0 = w0v[i], a1 = w1v[i], uu = u1v[i];
    double2 xx = xv[i], w;
    const double t0 = a * a0.x + c * a1.x, t1 = a * a0.y + c * a1.y;
    w.x = t0 + e * (LAZYU ? uu.x * mu : uu.x);
    w.y = t1 + e * (LAZYU ? uu.y * mu : uu.y);
    w0v[i] = w;
    xx.x += xf * w.x; xx.y += xf * w.y;
    xv[i] = xx;
    if (!LAZYU) { double2 qq = qv[i]; qq.x *= m; qq.y *= m; qv[i] = qq; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    MinresState s = *st;
    s.beta = beta; s.eta = etan;
    s.gamma0 = gamma1; s.gamma1 = gamma1n;
    s.sigma0 = sigma1; s.sigma1 = sigma1n;
    s.m0 = st->m1; s.m1 = m;
    s.mu = LAZYU ? m : 1.0;
    s.it = st->it + 1;
    s.done = (fabs(etan) <= s.norm_goal) ? 1 : (s.it >= maxiter ? 2 : 0);
    *st_next = s;
  }
}

// start of a solve: eta = beta = sqrt(<u1, v1>), goal = max(rtol*eta, atol); v1 /= beta, u1 /= beta.
__global__ void __launch_bounds__(BLK) k_minres_init(long n, double* __restrict__ v1, double* __restrict__ u1,
                                                     const double* bb_p, MinresState* st0, double rtol, double atol,
                                                     int lazy_u) {
  const double beta = sqrt(*bb_p);
  const double goal = fmax(rtol * beta, atol);
  const bool done = beta <= goal;
  if (!done && !lazy_u) {  // u1 /= beta (or lazily through mu); v1 /= beta lazily through m1
    const double m = 1.0 / beta;
    for (long i = (long)blockIdx.x * BLK + threadIdx.x; i < n; i += (long)gridDim.x * BLK) u1[i] *= m;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    MinresState s;
    s.beta = beta; s.eta = beta; s.eta0 = beta;
    s.gamma0 = 1.; s.gamma1 = 1.; s.sigma0 = 0.; s.sigma1 = 0.;
    s.norm_goal = goal;
    s.m0 = 1.0; s.m1 = done ? 1.0 : 1.0 / beta;
    s.mu = (lazy_u && !done) ? 1.0 / beta : 1.0;
    s.it = 0;
    s.done = done ? 1 : 0;
    *st0 = s;
  }
}

// y_u = r_u / diag(M)  (whole u-part, padding included: Md is 1 there)
__global__ void __launch_bounds__(BLK) k_jacobi_u(long n_u, const double* __restrict__ r, const double* __restrict__ Md,
                                                  double* __restrict__ y) {
  for (long i = (long)blockIdx.x * BLK + threadIdx.x; i < n_u; i += (long)gridDim.x * BLK) y[i] = r[i] / Md[i];
}

// ---------------------------------------------------------------------------------------
// K3: Schur block, parity mode -- one symmetric Gauss-Seidel sweep on S = B diag(M)^-1 B^T in
// lexicographic order (GSSmoother, type 0, 1 iteration, zero initial guess [MFEM-recall];
// [REF src/pressure_solver.cpp:84]), executed as hyperplane wavefronts i+j+k = const:
// identical dependency order, hence identical arithmetic, to the sequential sweep.
// Cooperative launch; one grid.sync per hyperplane.
// ---------------------------------------------------------------------------------------
// MODE 0: one CTA, __syncthreads per hyperplane;  MODE 1: one thread-block cluster, hardware
// cluster barrier per hyperplane;  MODE 2: whole grid (cooperative launch), grid.sync per hyperplane.
template <int MODE>
__device__ __forceinline__ void wave_sync(cg::grid_group& grid) {
  if (MODE == 0) __syncthreads();
  else if (MODE == 1) cg::this_cluster().sync();
  else grid.sync();
}

template <int MODE>
__global__ void __launch_bounds__(1024) k_sgs_wavefront(Lay L, const double* __restrict__ T,
                                                        const double* __restrict__ Md, const double* __restrict__ r,
                                                        double* y, int do_fwd, int do_bwd, const MinresState* st) {
  if (st && st->done) return;
  cg::grid_group grid = cg::this_grid();
  const int nstage = L.NX + L.NY + L.nzl - 2;
  const long npair = (long)L.NY * L.nzl;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long)gridDim.x * blockDim.x;
  const double* D = Md + L.off_p;
  const double* rp = r + L.off_p;
  double* yp = y + L.off_p;
  if (do_fwd) {
    for (int d = 0; d < nstage; ++d) {
      for (long jk = tid; jk < npair; jk += nth) {
        const int j = (int)(jk % L.NY), kl = (int)(jk / L.NY);
        const int i = d - j - kl;
        if (i < 0 || i >= L.NX) continue;
        const int kk = kl + 1, k = kl + L.zlo;
        const long c = ix_c(L, i, j, kk);
        double sum = 0.0;
        if (k > 0) sum += yp[c - L.plane] * (-T[ix_uz(L, i, j, kk)]);
        if (L.has_y && j > 0) sum += yp[c - L.NX] * (-T[ix_uy(L, i, j, kl)]);
        if (i > 0) sum += yp[c - 1] * (-T[ix_ux(L, i, j, kl)]);
        yp[c] = (rp[c] - sum) / D[c];
      }
      wave_sync<MODE>(grid);
    }
  }
  if (do_bwd) {
    for (int d = nstage - 1; d >= 0; --d) {
      for (long jk = tid; jk < npair; jk += nth) {
        const int j = (int)(jk % L.NY), kl = (int)(jk / L.NY);
        const int i = d - j - kl;
        if (i < 0 || i >= L.NX) continue;
        const int kk = kl + 1, k = kl + L.zlo;
        const long c = ix_c(L, i, j, kk);
        double sum = 0.0;
        if (k > 0) sum += yp[c - L.plane] * (-T[ix_uz(L, i, j, kk)]);
        if (L.has_y && j > 0) sum += yp[c - L.NX] * (-T[ix_uy(L, i, j, kl)]);
        if (i > 0) sum += yp[c - 1] * (-T[ix_ux(L, i, j, kl)]);
        if (i + 1 < L.NX) sum += yp[c + 1] * (-T[ix_ux(L, i + 1, j, kl)]);
        if (L.has_y && j + 1 < L.NY) sum += yp[c + L.NX] * (-T[ix_uy(L, i, j + 1, kl)]);
        if (k + 1 < L.NZ) sum += yp[c + L.plane] * (-T[ix_uz(L, i, j, kk + 1)]);
        yp[c] = (rp[c] - sum) / D[c];
      }
      if (d > 0) wave_sync<MODE>(grid);
    }
  }
}

// ---------------------------------------------------------------------------------------
// K3 (fast path): the same lexicographic symmetric Gauss-Seidel sweep on a SKEWED (stage-major)
// copy of the Schur stencil.  Pencil p = j + NY*kl runs along x; cell (i, p) sits in stage
// d = i + j + kl at slot d*NPs + p.  At stage d, thread p touches slot (d, p) and its neighbours
// (d-1, p), (d-1, p-1), (d-1, p-NY) / (d+1, p), (d+1, p+1), (d+1, p+NY): every access is
// coalesced across the threads of a stage and known one stage ahead, so it is prefetched into
// registers before the barrier.  Neighbour y values travel through shared memory (one CTA) or
// L2 (cluster).  The arithmetic per cell is unchanged, so the result is bit-identical to the
// sequential sweep.
// ---------------------------------------------------------------------------------------
struct SgsPack {
  // Stage-major block layout: stage d owns one contiguous block of 6 rows of NPs doubles
  //   row 0 Rs (skewed rhs), 1 D, 2 rD = RN(1/D), 3 -Tx_lo, 4 -Tz_lo, 5 Yf (forward result), 6 -Ty_lo;
  // (rows ordered so that a 2-D sweep streams only the first 5 (forward) / 6 (backward) rows)
  // the pointers below address row r of stage 0, consecutive stages are SS = 7*NPs apart.
  double *Rs, *D, *rD, *nTx, *nTy, *nTz, *Yf;
  double* Yb;  // backward result, [stage][NPs]
  int NPs;     // pencils padded to a multiple of 32
  int SS;      // stage stride of the block layout
};

__device__ __forceinline__ long skew_blk(const Lay& L, const SgsPack& P, int i, int j, int kl) {
  return (long)(i + j + kl) * P.SS + (j + L.NY * kl);
}
__device__ __forceinline__ long skew_row(const Lay& L, const SgsPack& P, int i, int j, int kl) {
  return (long)(i + j + kl) * P.NPs + (j + L.NY * kl);
}

// once per pressure solve: natural -> skewed copy of the stencil
__global__ void __launch_bounds__(BLK) k_sgs_pack(Lay L, SgsPack P, const double* __restrict__ T,
                                                  const double* __restrict__ Md) {
  const CellIdx ci = owned_cell(L);
  if (!ci.ok) return;
  const int i = ci.i, j = ci.j, kl = ci.kl, kk = kl + 1;
  const long s = skew_blk(L, P, i, j, kl);
  const double dg = Md[ix_p(L, i, j, kk)];
  P.D[s] = dg;
  P.rD[s] = 1.0 / dg;
  P.nTx[s] = -T[ix_ux(L, i, j, kl)];
  P.nTy[s] = L.has_y ? -T[ix_uy(L, i, j, kl)] : 0.0;
  P.nTz[s] = -T[ix_uz(L, i, j, kk)];
}
// per application: r (natural p-part) -> Rs ; Y (skewed) -> y (natural p-part)
__global__ void __launch_bounds__(BLK) k_sgs_skew_in(Lay L, SgsPack P, const double* __restrict__ r,
                                                     const MinresState* st) {
  if (st && st->done) return;
  const CellIdx ci = owned_cell(L);
  if (!ci.ok) return;
  P.Rs[skew_blk(L, P, ci.i, ci.j, ci.kl)] = r[ix_p(L, ci.i, ci.j, ci.kl + 1)];
}
__global__ void __launch_bounds__(BLK) k_sgs_skew_out(Lay L, SgsPack P, int from_bwd, double* __restrict__ y,
                                                      const MinresState* st) {
  if (st && st->done) return;
  const CellIdx ci = owned_cell(L);
  if (!ci.ok) return;
  y[ix_p(L, ci.i, ci.j, ci.kl + 1)] =
      from_bwd ? P.Yb[skew_row(L, P, ci.i, ci.j, ci.kl)] : P.Yf[skew_blk(L, P, ci.i, ci.j, ci.kl)];
}

constexpr int SGS_T = 512;  // threads per CTA of the register-prefetch skewed sweep

// Register-prefetch variant: MODE 0 one CTA (neighbours through shared memory), MODE 1 one cluster
// (neighbours through L2 after the hardware cluster barrier).  PPT pencils per thread.
template <int MODE, int PPT>
__global__ void __launch_bounds__(SGS_T) k_sgs_skew(Lay L, SgsPack P, const double* __restrict__ T, double* y,
                                                    int do_fwd, int do_bwd, const MinresState* st) {
  if (st && st->done) return;
  __shared__ double sh[2][MODE == 0 ? SGS_T * PPT : 1];
  const int nth = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int NP = L.NY * L.nzl, NPs = P.NPs;
  const long SS = P.SS;
  const int nstage = L.NX + L.NY + L.nzl - 2;
  double* yp = y + L.off_p;
  int pj[PPT], pk[PPT], pp[PPT];
  bool pv[PPT];
#pragma unroll
  for (int s = 0; s < PPT; ++s) {
    pp[s] = tid + s * nth;
    pv[s] = pp[s] < NP;
    const int q = pv[s] ? pp[s] : 0;
    pj[s] = q % L.NY;
    pk[s] = q / L.NY;
  }
  auto sync = [&]() {
    if (MODE == 0) __syncthreads();
    else cg::this_cluster().sync();
  };
  if (do_fwd) {
    double R[PPT], D[PPT], nx[PPT], ny[PPT], nz[PPT];
    auto load = [&](int d) {
#pragma unroll
      for (int s = 0; s < PPT; ++s) {
        const int i = d - pj[s] - pk[s];
        if (pv[s] && i >= 0 && i < L.NX) {
          const long o = (long)d * SS + pp[s];
          R[s] = P.Rs[o]; D[s] = P.D[o]; nx[s] = P.nTx[o]; ny[s] = P.nTy[o]; nz[s] = P.nTz[o];
        }
      }
    };
    load(0);
    for (int d = 0; d < nstage; ++d) {
      const double* prev = P.Yf + (long)(d - 1) * SS;
#pragma unroll
      for (int s = 0; s < PPT; ++s) {
        const int i = d - pj[s] - pk[s];
        if (pv[s] && i >= 0 && i < L.NX) {
          const int p = pp[s], k = pk[s] + L.zlo;
          double sum = 0.0;
          if (k > 0) {
            const double yl = pk[s] > 0 ? (MODE == 0 ? sh[(d - 1) & 1][p - L.NY] : __ldcg(prev + p - L.NY))
                                        : __ldcg(yp + ix_c(L, i, pj[s], 0));
            sum += yl * nz[s];
          }
          if (L.has_y && pj[s] > 0) sum += (MODE == 0 ? sh[(d - 1) & 1][p - 1] : __ldcg(prev + p - 1)) * ny[s];
          if (i > 0) sum += (MODE == 0 ? sh[(d - 1) & 1][p] : __ldcg(prev + p)) * nx[s];
          const double v = (R[s] - sum) / D[s];
          P.Yf[(long)d * SS + p] = v;
          if (MODE == 0) sh[d & 1][p] = v;
        }
      }
      if (d + 1 < nstage) load(d + 1);
      sync();
    }
  }
  if (do_bwd) {
    double R[PPT], D[PPT], nx[PPT], ny[PPT], nz[PPT], hx[PPT], hy[PPT], hz[PPT], fx[PPT], fy[PPT], fz[PPT];
    auto load = [&](int d) {
#pragma unroll
      for (int s = 0; s < PPT; ++s) {
        const int i = d - pj[s] - pk[s];
        if (pv[s] && i >= 0 && i < L.NX) {
          const int p = pp[s], k = pk[s] + L.zlo;
          const long o = (long)d * SS + p;
          R[s] = P.Rs[o]; D[s] = P.D[o]; nx[s] = P.nTx[o]; ny[s] = P.nTy[o]; nz[s] = P.nTz[o];
          const long up = o + SS, lo = o - SS;
          if (i + 1 < L.NX) hx[s] = P.nTx[up];
          if (L.has_y && pj[s] + 1 < L.NY) hy[s] = P.nTy[up + 1];
          if (k + 1 < L.NZ) hz[s] = pk[s] + 1 < L.nzl ? P.nTz[up + L.NY] : -T[ix_uz(L, i, pj[s], pk[s] + 2)];
          if (i > 0) fx[s] = P.Yf[lo];
          if (L.has_y && pj[s] > 0) fy[s] = P.Yf[lo - 1];
          if (k > 0) fz[s] = pk[s] > 0 ? P.Yf[lo - L.NY] : __ldcg(yp + ix_c(L, i, pj[s], 0));
        }
      }
    };
    load(nstage - 1);
    for (int d = nstage - 1; d >= 0; --d) {
      const double* next = P.Yb + (long)(d + 1) * NPs;
#pragma unroll
      for (int s = 0; s < PPT; ++s) {
        const int i = d - pj[s] - pk[s];
        if (pv[s] && i >= 0 && i < L.NX) {
          const int p = pp[s], k = pk[s] + L.zlo;
          double sum = 0.0;
          if (k > 0) sum += fz[s] * nz[s];
          if (L.has_y && pj[s] > 0) sum += fy[s] * ny[s];
          if (i > 0) sum += fx[s] * nx[s];
          if (i + 1 < L.NX) sum += (MODE == 0 ? sh[(d + 1) & 1][p] : __ldcg(next + p)) * hx[s];
          if (L.has_y && pj[s] + 1 < L.NY) sum += (MODE == 0 ? sh[(d + 1) & 1][p + 1] : __ldcg(next + p + 1)) * hy[s];
          if (k + 1 < L.NZ) {
            const double yu = pk[s] + 1 < L.nzl ? (MODE == 0 ? sh[(d + 1) & 1][p + L.NY] : __ldcg(next + p + L.NY))
                                                : __ldcg(yp + ix_c(L, i, pj[s], pk[s] + 2));
            sum += yu * hz[s];
          }
          const double v = (R[s] - sum) / D[s];
          P.Yb[(long)d * NPs + p] = v;
          if (MODE == 0) sh[d & 1][p] = v;
        }
      }
      if (d > 0) {
        load(d - 1);
        sync();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// K3 (TMA path, one CTA): the same skewed sweep with all operands staged by the TMA engine.
// Stage d's block (6 rows x NPs doubles, contiguous) is brought into a shared-memory ring by ONE
// cp.async.bulk (mbarrier complete_tx) issued NS-2 stages ahead by a dedicated producer warp, which
// is also the only waiter on the mbarriers; the compute warps see "slot ready" through the per-stage
// __syncthreads.  Per-stage critical path of a compute warp: LDS operands + neighbour values,
// <= 6 mul/add, one division, STS (+ a fire-and-forget coalesced global store), barrier.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// blockDim.x = 32 (producer warp) + NPs (one compute thread per pencil).
// Shared memory: ring[NS][7][NPs] | ybuf[2][YP + NPs + YP] | mbarriers[NS],  YP = padding >= NY+1 (zeros).
// The stage body is branch-free: neighbours that do not exist carry a zero coefficient (T = 0 on boundary
// faces, packed slots outside the mesh are zero) and the y they multiply is always finite.
// NPS_T > 0 fixes the row length at compile time (row offsets become immediates); HAS_Y_T = 3-D.
// All ring / phase bookkeeping is incremental (no runtime division or 64-bit multiply inside the loop):
// the producer lane's per-stage work is what every compute warp waits for at the barrier.
template <int NPS_T, bool HAS_Y_T>
__global__ void __launch_bounds__(1024) k_sgs_tma(Lay L, SgsPack P, const double* __restrict__ T, double* y, int NS,
                                                  int YP, int do_fwd, int do_bwd, const MinresState* st) {
  if (st && st->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int NSa = NS < 0 ? -NS : NS;  // NS < 0: timing experiment without the per-stage global store
  const bool no_tma = YP < 0;         // YP < 0: timing experiment without the producer's TMA traffic
  YP = YP < 0 ? -YP : YP;
  const int NPs = NPS_T > 0 ? NPS_T : P.NPs;
  const int NY = HAS_Y_T ? L.NY : 1;
  const int NP = NY * L.nzl;
  const int SS = 7 * NPs;
  const int nstage = L.NX + NY + L.nzl - 2;
  const int YR = NPs + 2 * YP;                                                               // ybuf row length
  double* ring = reinterpret_cast<double*>(smem_raw);                                        // [NS][7][NPs]
  double* ybuf = ring + (size_t)NSa * SS;                                                     // [2][YR]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ybuf + (size_t)2 * YR);   // [NS]
  const bool producer = threadIdx.x < 32;
  const int p = (int)threadIdx.x - 32;
  const bool pv = !producer && p < NP;
  const int pj = pv ? p % NY : 0, pk = pv ? p / NY : 0;
  const int s0 = pj + pk;                       // first stage in which this pencil is active
  const int k = pk + L.zlo;
  const bool ghost_lo = pv && pk == 0 && k > 0;                  // multi-rank: lower neighbour on another GPU
  const bool ghost_hi = pv && pk + 1 == L.nzl && k + 1 < L.NZ;   // multi-rank: upper neighbour on another GPU
  const bool in_lo = pk > 0, in_hi = pk + 1 < L.nzl, has_jlo = HAS_Y_T && pj > 0, has_jhi = HAS_Y_T && pj + 1 < NY;
  double* yp = y + L.off_p;
  const unsigned rowb = (unsigned)NPs * 8u;
  const int ringn = NSa * SS;

  // ring and ybuf start as zeros: never-loaded slots / padding must hold finite values (they only ever meet
  // zero coefficients)
  for (int q = threadIdx.x; q < ringn + 2 * YR; q += blockDim.x) ring[q] = 0.0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSa; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // running ring state shared by both phases (slot index and mbarrier parity of the next issue / wait / use)
  int sl = 0, par = 0;                 // compute: slot offset (elements) and ybuf row of the current iteration
  int is_slot = 0;                     // producer: slot of the next block to issue
  int wt_slot = 0;
  unsigned wt_par = 0;                 // producer: slot / parity of the next block to wait for

  for (int dir = 0; dir < 2; ++dir) {
    if (dir == 0 ? !do_fwd : !do_bwd) continue;
    const unsigned bytes = rowb * (HAS_Y_T ? 7u : (dir == 0 ? 5u : 6u));  // 2-D: skip the unused rows
    const long gstep = dir == 0 ? (long)SS : -(long)SS;
    const double* is_src = P.Rs + (dir == 0 ? 0 : (size_t)(nstage - 1) * SS);  // next block to issue (global)
    int n_issued = 0, n_waited = 0;
    auto issue1 = [&]() {  // producer lane 0 only
      unsigned long long* bar = bars + is_slot;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(ring + is_slot * SS, is_src, bytes, bar);
      is_src += gstep;
      is_slot = is_slot + 1 == NSa ? 0 : is_slot + 1;
      ++n_issued;
    };
    auto wait1 = [&]() {
      mbar_wait(bars + wt_slot, wt_par);
      if (wt_slot + 1 == NSa) { wt_slot = 0; wt_par ^= 1u; }
      else ++wt_slot;
      ++n_waited;
    };
    if (dir == 1 && do_fwd) {
      // forward rows (row Yf of every block) were written with ordinary stores: order them before the
      // async-proxy reads of the backward pass
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (threadIdx.x == 0 && !no_tma) {
      while (n_issued < NSa - 2 && n_issued < nstage) issue1();
      wait1();
      if (nstage > 1) wait1();
    }
    __syncthreads();
    if (producer) {
      for (int n = 0; n < nstage; ++n) {
        if (threadIdx.x == 0 && !no_tma) {
          // the slot of iteration n-2 is free (last read, as upper-face rows, during n-1)
          if (n_issued < nstage) issue1();
          if (n_waited < nstage) wait1();
        }
        __syncthreads();
      }
    } else if (dir == 0) {
      int i = -s0;
      double* gout = P.Yf + p;
      for (int n = 0; n < nstage; ++n, ++i) {
        if ((unsigned)i < (unsigned)L.NX && pv) {
          const double* slot = ring + sl + p;
          double* ycur = ybuf + par * YR + YP + p;
          const double* yadj = ybuf + (par ^ 1) * YR + YP + p;  // previous iteration's row
          const double R = slot[0], D = slot[NPs], rD = slot[2 * NPs], nx = slot[3 * NPs], nz = slot[4 * NPs];
          const double yz = ghost_lo ? __ldcg(yp + ix_c(L, i, pj, 0)) : yadj[-NY];
          double sum = yz * nz;
          if (HAS_Y_T) sum += yadj[-1] * slot[6 * NPs];
          sum += yadj[0] * nx;
          const double v = div_exact(R - sum, D, rD);
          ycur[0] = v;
          if (NS > 0) *gout = v;
        }
        __syncthreads();
        gout += SS;
        sl = sl + SS == ringn ? 0 : sl + SS;
        par ^= 1;
      }
    } else {
      int i = nstage - 1 - s0;
      double* gout = P.Yb + p + (size_t)(nstage - 1) * NPs;
      for (int n = 0; n < nstage; ++n, --i) {
        if ((unsigned)i < (unsigned)L.NX && pv) {
          const int sn = sl + SS == ringn ? 0 : sl + SS, su = sl == 0 ? ringn - SS : sl - SS;
          const double* slot = ring + sl + p;
          const double* F = ring + sn + 5 * NPs + p;  // block of stage d-1, row Yf
          const double* up = ring + su + p;           // block of stage d+1
          double* ycur = ybuf + par * YR + YP + p;
          const double* yadj = ybuf + (par ^ 1) * YR + YP + p;  // previous iteration's row (stage d+1)
          const double R = slot[0], D = slot[NPs], rD = slot[2 * NPs], nx = slot[3 * NPs], nz = slot[4 * NPs];
          const double fz = ghost_lo ? __ldcg(yp + ix_c(L, i, pj, 0)) : (in_lo ? F[-NY] : 0.0);
          double sum = fz * nz;
          if (HAS_Y_T) sum += (has_jlo ? F[-1] : 0.0) * slot[6 * NPs];
          sum += F[0] * nx;
          sum += yadj[0] * (i + 1 < L.NX ? up[3 * NPs] : 0.0);
          if (HAS_Y_T) sum += yadj[1] * (has_jhi ? up[6 * NPs + 1] : 0.0);
          const double hz = ghost_hi ? -T[ix_uz(L, i, pj, pk + 2)] : (in_hi ? up[4 * NPs + NY] : 0.0);
          const double yu = ghost_hi ? __ldcg(yp + ix_c(L, i, pj, pk + 2)) : yadj[NY];
          sum += yu * hz;
          const double v = div_exact(R - sum, D, rD);
          ycur[0] = v;
          if (NS > 0) *gout = v;
        }
        __syncthreads();
        gout -= NPs;
        sl = sl + SS == ringn ? 0 : sl + SS;
        par ^= 1;
      }
    }
  }
}

// small PTX helpers shared by the staged kernels
__device__ __forceinline__ int ld_volatile_s32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_volatile_s32(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src_gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------
// K6: one forward-Euler sub-step  S <- S + dt M_phi^-1 (Kt f(S) + b), fused
//     (FS, Kt.Mult, z += b, M solve, y *= dt, S += y  [REF src/saturation_solver.cpp:76-91]);
//     P0/RT0 closed form of the DG upwind operator:
//       z_c = sum_F upwind flux, summed in the CSR column order of the assembled Kt.
//     RATE = true writes dS/dt = M_phi^-1 (Kt f(S) + b) instead (stage of an RK scheme).
// ---------------------------------------------------------------------------------------
struct SatParam {
  int two_phase;       // FS applied (two-phase) or identity (single-phase) [REF src/saturation_solver.cpp:78-86]
  double dt, vol;
  double hx, hy, hz;   // internal spacings
  int well_2d;         // injection box test in (x, internal z) only (2-D) or 3-D rule
  double inj_c0, inj_c1, inj_c2, inj_r, inj_h, src;
  // the same box as inclusive INTERNAL cell-index ranges (cell centres tested on the host with the formulas of
  // sat_source; membership is monotone along each axis); empty box: lo > hi
  int bi0, bi1, bj0, bj1, bk0, bk1;
  // DG-Q1: inclusive cell-index ranges of the cells with at least one DG node inside the box (per-axis predicates at
  // the node positions (i + sn) h, evaluated on the host with the kernel's expression); empty: lo > hi
  int ni0, ni1, nj0, nj1, nk0, nk1;
};
__device__ __forceinline__ double sat_source_box(const SatParam& sp, int i, int j, int k) {
  return (i >= sp.bi0 && i <= sp.bi1 && j >= sp.bj0 && j <= sp.bj1 && k >= sp.bk0 && k <= sp.bk1) ? sp.src : 0.;
}

__device__ __forceinline__ double sat_source(const SatParam& sp, int i, int j, int k) {
  // SaturationSourceCoefficient::eval2D/eval3D at the cell centre [REF src/param.hpp:382-406]
  const double px = (i + 0.5) * sp.hx;
  if (sp.well_2d) {
    const double py = (k + 0.5) * sp.hz;  // 2-D y == internal z
    if (fabs(px - sp.inj_c0) <= sp.inj_r && fabs(py - sp.inj_c1) <= sp.inj_r) return sp.src;
    return 0.;
  }
  const double py = (j + 0.5) * sp.hy, pz = (k + 0.5) * sp.hz;
  if (fabs(px - sp.inj_c0) <= sp.inj_r && fabs(pz - sp.inj_c2) <= sp.inj_r && py >= sp.inj_c1 &&
      py <= sp.inj_c1 + sp.inj_h)
    return sp.src;
  return 0.;
}

// K6, streaming form (default): f(S) is carried as a second resident array.  A sub-step reads f through the
// 7-point stencil (neighbours hit L1/L2), its own S, the face fluxes and phi, and writes S_new together with
// f(S_new): exactly one fractional-flow evaluation per cell and step, no shared memory, no barrier.
// Actual traffic 64 B/cell-step in 3-D (reads fs,
// S, 3 fluxes, phi; writes S, fs) against 48 B algorithmic.
__global__ void __launch_bounds__(BLK) k_frac_flow(Lay L, Phys ph, int two_phase, const double* __restrict__ S,
                                                   double* __restrict__ F) {
  // over the whole cvec (ghost planes included: they were exchanged for S already)
  const long c = (long)blockIdx.x * BLK + threadIdx.x;
  if (c >= L.ncv) return;
  const double s = S[c];
  F[c] = two_phase ? frac_flow_of(ph, s) : s;
}

template <bool RATE>
__global__ void __launch_bounds__(BLK) k_sat_stream(Lay L, Phys ph, SatParam sp, const double* __restrict__ u,
                                                    const double* __restrict__ phi, const double* __restrict__ Sin,
                                                    const double* __restrict__ Fin, double* __restrict__ Sout,
                                                    double* __restrict__ Fout) {
  const CellIdx ci = owned_cell(L);
  if (!ci.ok) return;
  const int i = ci.i, j = ci.j, kl = ci.kl, k = ci.k, kk = kl + 1;
  const int c = ix_c(L, i, j, kk);
  const bool xl = i > 0, xh = i + 1 < L.NX, yl = L.has_y && j > 0, yh = L.has_y && j + 1 < L.NY, zl = k > 0,
             zh = k + 1 < L.NZ;
  // Branch-free: every load is unconditional (a missing neighbour is clamped to the cell itself and its flux is
  // replaced by +0), so all 15 loads are issued before any arithmetic; a skipped term becomes `+ 0 * f`, which
  // leaves the partial sums bit-identical (f finite; zeros only enter sums that start from +0).
  const int fx = ix_ux(L, i, j, kl), fz = ix_uz(L, i, j, kk);
  const int fy = L.has_y ? ix_uy(L, i, j, kl) : fx;
  const double uxl = u[fx], uxh = u[fx + 1], uyl = u[fy], uyh = u[L.has_y ? fy + L.NX : fx], uzl = u[fz],
               uzh = u[fz + L.plane];
  const double f_c = Fin[c];
  const double f_zl = Fin[zl ? c - L.plane : c], f_yl = Fin[yl ? c - L.NX : c], f_xl = Fin[xl ? c - 1 : c];
  const double f_xh = Fin[xh ? c + 1 : c], f_yh = Fin[yh ? c + L.NX : c], f_zh = Fin[zh ? c + L.plane : c];
  const double ph_c = phi[c];
  const double s_c = RATE ? 0.0 : Sin[c];
  // (F)+ = F > 0 ? F : 0 and (F)- = F < 0 ? F : 0 exactly as the oracle writes them, with the "neighbour exists"
  // predicate folded into the comparison: one DSETP + two FSEL each (fp64 fmax/fmin cost ~8 instructions apiece
  // because of their NaN handling -- ncu: they were a third of this kernel's instructions)
  const double pxl = (xl && uxl > 0.0) ? uxl : 0.0, nxl = (xl && uxl < 0.0) ? uxl : 0.0;
  const double pxh = (xh && uxh > 0.0) ? uxh : 0.0, nxh = (xh && uxh < 0.0) ? uxh : 0.0;
  const double pyl = (yl && uyl > 0.0) ? uyl : 0.0, nyl = (yl && uyl < 0.0) ? uyl : 0.0;
  const double pyh = (yh && uyh > 0.0) ? uyh : 0.0, nyh = (yh && uyh < 0.0) ? uyh : 0.0;
  const double pzl = (zl && uzl > 0.0) ? uzl : 0.0, nzl = (zl && uzl < 0.0) ? uzl : 0.0;
  const double pzh = (zh && uzh > 0.0) ? uzh : 0.0, nzh = (zh && uzh < 0.0) ? uzh : 0.0;
  double z = 0.0;
  z += pzl * f_zl;
  if (L.has_y) z += pyl * f_yl;
  z += pxl * f_xl;
  {
    double dg = 0.0;
    dg += nxl;
    dg += -pxh;
    if (L.has_y) {
      dg += nyl;
      dg += -pyh;
    }
    dg += nzl;
    dg += -pzh;
    z += dg * f_c;
  }
  z += (-nxh) * f_xh;
  if (L.has_y) z += (-nyh) * f_yh;
  z += (-nzh) * f_zh;
  z += sat_source_box(sp, i, j, k) * sp.vol;
  double yv = div_rcp(z, ph_c * sp.vol);
  if (RATE) {
    Sout[c] = yv;
  } else {
    yv *= sp.dt;
    const double sn = s_c + yv;
    Sout[c] = sn;
    double fn = sn;
    if (sp.two_phase) {  // FS [REF src/saturation_solver.cpp:8-17]
      double mw, mo;
      mobilities(ph, sn, mw, mo);
      fn = div_rcp(mw, mw + mo);
    }
    Fout[c] = fn;
  }
}

// RK stage combinations (MPI path's -ode variants): out = a*x + b*y (owned + ghosts)
__global__ void __launch_bounds__(BLK) k_axpby(long n, double a, const double* __restrict__ x, double b,
                                               const double* __restrict__ y, double* __restrict__ out) {
  for (long i = (long)blockIdx.x * BLK + threadIdx.x; i < n; i += (long)gridDim.x * BLK) out[i] = a * x[i] + b * y[i];
}
// out = x + b*y
__global__ void __launch_bounds__(BLK) k_xpby(long n, const double* __restrict__ x, double b,
                                              const double* __restrict__ y, double* __restrict__ out) {
  for (long i = (long)blockIdx.x * BLK + threadIdx.x; i < n; i += (long)gridDim.x * BLK) out[i] = x[i] + b * y[i];
}

// K7: guards of the global loop [REF src/run_serial.cpp:132-142]: #non-finite, max(S) over owned cells.
// Warp-aggregated: lanes combine their count / maximum with shuffles and lane 0 issues at most two atomics (one
// atomicMax per dof on a single address cost 88 ms for the 1.3e8 dofs of a 256^3 DG-Q1 field).
__device__ __forceinline__ unsigned long long guard_key(double s) {
  // order-preserving map double -> uint64 so that an unsigned maximum works for negative values too
  const unsigned long long b = (unsigned long long)__double_as_longlong(s);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ void guard_commit(unsigned bad, unsigned long long key, unsigned long long* n_bad,
                                             unsigned long long* max_bits) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
    const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
    key = k2 > key ? k2 : key;
  }
  if ((threadIdx.x & 31) == 0) {
    if (bad) atomicAdd(n_bad, (unsigned long long)bad);
    if (key) atomicMax(max_bits, key);
  }
}
__global__ void __launch_bounds__(BLK) k_guards(Lay L, const double* __restrict__ S, unsigned long long* n_bad,
                                                unsigned long long* max_bits) {
  const CellIdx ci = owned_cell(L);
  unsigned bad = 0;
  unsigned long long key = 0ull;   // 0 = "no finite value seen" (below the key of every double)
  if (ci.ok) {
    const double s = S[ix_c(L, ci.i, ci.j, ci.kl + 1)];
    if (!isfinite(s)) bad = 1;
    else key = guard_key(s);
  }
  guard_commit(bad, key, n_bad, max_bits);
}

// L2 flush helper for impes_time_kernel
__global__ void k_fill(double* p, long n, double v) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace impes
