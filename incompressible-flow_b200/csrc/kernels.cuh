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
    double a = 0.0;
    for (unsigned j = threadIdx.x; j < gridDim.x; j += BLK) a += __ldcg(partial + j);
    __syncthreads();
    const double tot = block_tree_256(a, sm);
    if (threadIdx.x == 0) {
      *out = tot;
      *ticket = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------
// physics: Param::Krw / Param::Kro  [REF src/param.cpp:195-215]  (no clamping)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double krw_of(const Phys& ph, double S) {
  if (ph.two_phase) {
    const double S_ = (S - ph.Swc) / (1. - ph.Sor - ph.Swc);
    return S_ * S_;
  }
  return S;
}
__device__ __forceinline__ double kro_of(const Phys& ph, double S) {
  if (ph.two_phase) {
    const double S_ = (S - ph.Swc) / (1. - ph.Sor - ph.Swc);
    return (1. - S_) * (1. - S_);
  }
  return (1. - S);
}
// FS  [REF src/saturation_solver.cpp:8-17]
__device__ __forceinline__ double frac_flow_of(const Phys& ph, double S) {
  const double mw = krw_of(ph, S) / ph.mu_w;
  const double mo = kro_of(ph, S) / ph.mu_o;
  return mw / (mw + mo);
}

struct CellIdx {
  int i, j, kl, k;  // kl = local plane, k = global plane
  bool ok;
};
__device__ __forceinline__ CellIdx owned_cell(const Lay& L) {
  CellIdx c;
  const long t = (long)blockIdx.x * BLK + threadIdx.x;
  c.ok = t < L.ncl;
  const long tt = c.ok ? t : 0;
  c.i = (int)(tt % L.NX);
  const long r = tt / L.NX;
  c.j = (int)(r % L.NY);
  c.kl = (int)(r / L.NY);
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
  const double lam = krw_of(ph, s) / ph.mu_w + kro_of(ph, s) / ph.mu_o;
  mx[c] = (1. / (lam * Kx[c])) * gx;
  if (L.has_y) my[c] = (1. / (lam * Ky[c])) * gy;
  mz[c] = (1. / (lam * Kz[c])) * gz;
}

// diag(M) on owned faces (1 on boundary faces: eliminated rows [REF src/pressure_solver.cpp:56-60]),
// T_f = 1/diag(M)_f on interior faces (0 on boundary faces), and the diagonal of
// S = B diag(M)^-1 B^T  [REF src/pressure_solver.cpp:76-81] stored in the p-part of Md.
__global__ void __launch_bounds__(BLK) k_face_coeff(Lay L, const double* __restrict__ mx, const double* __restrict__ my,
                                                    const double* __restrict__ mz, double* __restrict__ Md,
                                                    double* __restrict__ T) {
  const CellIdx ci = owned_cell(L);
  if (!ci.ok) return;
  const int i = ci.i, j = ci.j, kl = ci.kl, k = ci.k, kk = kl + 1;
  const long c = ix_c(L, i, j, kk);
  double D = 0.0;
  {  // x
    const long f = ix_ux(L, i, j, kl);
    if (i > 0) {
      const double md = mx[c - 1] * THIRD + mx[c] * THIRD;
      const double t = 1.0 / md;
      Md[f] = md; T[f] = t; D += t;
    } else { Md[f] = 1.0; T[f] = 0.0; }
    if (i < L.NX - 1) {
      const double md = mx[c] * THIRD + mx[c + 1] * THIRD;
      D += 1.0 / md;
    } else { Md[f + 1] = 1.0; T[f + 1] = 0.0; }
  }
  if (L.has_y) {
    const long f = ix_uy(L, i, j, kl);
    if (j > 0) {
      const double md = my[c - L.NX] * THIRD + my[c] * THIRD;
      const double t = 1.0 / md;
      Md[f] = md; T[f] = t; D += t;
    } else { Md[f] = 1.0; T[f] = 0.0; }
    if (j < L.NY - 1) {
      const double md = my[c] * THIRD + my[c + L.NX] * THIRD;
      D += 1.0 / md;
    } else { Md[f + L.NX] = 1.0; T[f + L.NX] = 0.0; }
  }
  {  // z
    const long f = ix_uz(L, i, j, kk);
    if (k > 0) {
      const double md = mz[c - L.plane] * THIRD + mz[c] * THIRD;
      const double t = 1.0 / md;
      Md[f] = md; T[f] = t; D += t;
    } else { Md[f] = 1.0; T[f] = 0.0; }
    if (k < L.NZ - 1) {
      const double md = mz[c] * THIRD + mz[c + L.plane] * THIRD;
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
template <bool DOT>
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
    const long c = ix_c(L, i, j, kk);
    const long ip = L.off_p + c;
    const double pc = x[ip];
    double yp = 0.0;
    {  // ---- x faces
      const long f = ix_ux(L, i, j, kl);
      const double xf = x[f], xf1 = x[f + 1];
      if (i > 0) {
        const double mL = mx[c - 1], mR = mx[c];
        double s = 0.0;
        if (i - 1 > 0) s += (mL * SIXTH) * x[f - 1];
        s += (mL * THIRD + mR * THIRD) * xf;
        if (i + 1 < L.NX) s += (mR * SIXTH) * xf1;
        const double yf = s + (pc - x[ip - 1]);
        y[f] = yf;
        if (DOT) acc += xf * yf;
        yp += xf;
      } else {
        y[f] = xf;
        if (DOT) acc += xf * xf;
      }
      if (i + 1 < L.NX) yp -= xf1;
      else {
        y[f + 1] = xf1;
        if (DOT) acc += xf1 * xf1;
      }
    }
    if (L.has_y) {  // ---- y faces
      const long f = ix_uy(L, i, j, kl);
      const double xf = x[f], xf1 = x[f + L.NX];
      if (j > 0) {
        const double mL = my[c - L.NX], mR = my[c];
        double s = 0.0;
        if (j - 1 > 0) s += (mL * SIXTH) * x[f - L.NX];
        s += (mL * THIRD + mR * THIRD) * xf;
        if (j + 1 < L.NY) s += (mR * SIXTH) * xf1;
        const double yf = s + (pc - x[ip - L.NX]);
        y[f] = yf;
        if (DOT) acc += xf * yf;
        yp += xf;
      } else {
        y[f] = xf;
        if (DOT) acc += xf * xf;
      }
      if (j + 1 < L.NY) yp -= xf1;
      else {
        y[f + L.NX] = xf1;
        if (DOT) acc += xf1 * xf1;
      }
    }
    {  // ---- z faces
      const long f = ix_uz(L, i, j, kk);
      const double xf = x[f], xf1 = x[f + L.plane];
      if (k > 0) {
        const double mL = mz[c - L.plane], mR = mz[c];
        double s = 0.0;
        if (k - 1 > 0) s += (mL * SIXTH) * x[f - L.plane];
        s += (mL * THIRD + mR * THIRD) * xf;
        if (k + 1 < L.NZ) s += (mR * SIXTH) * xf1;
        const double yf = s + (pc - x[ip - L.plane]);
        y[f] = yf;
        if (DOT) acc += xf * yf;
        yp += xf;
      } else {
        y[f] = xf;
        if (DOT) acc += xf * xf;
      }
      if (k + 1 < L.NZ) yp -= xf1;
      else {
        y[f + L.plane] = xf1;
        if (DOT) acc += xf1 * xf1;
      }
    }
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
  const double alpha = *alpha_p, beta = st->beta;
  for (long i = (long)blockIdx.x * BLK + threadIdx.x; i < n; i += (long)gridDim.x * BLK) {
    const double t = q_in[i] + (-beta) * v0[i];
    const double v = t + (-alpha) * v1[i];
    v0[i] = v;
    if (i < n_u) q_out[i] = v / Md[i];
  }
}

// pass C: scalar recurrence + w/x update + scaling of the next Lanczos pair.
//   w0 <- (-rho3/rho1) w0 + (-rho2/rho1) w1 + (1/rho1) u1 ; x += gamma1 eta w0 ;
//   v0 *= 1/beta_new ; q *= 1/beta_new      (become v1, u1 of the next iteration)
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
  for (long i = (long)blockIdx.x * BLK + threadIdx.x; i < n; i += (long)gridDim.x * BLK) {
    const double t = a * w0[i] + c * w1[i];
    const double w = t + e * u1[i];
    w0[i] = w;
    x[i] += xf * w;
    v0[i] *= m;
    q[i] *= m;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    MinresState s = *st;
    s.beta = beta; s.eta = etan;
    s.gamma0 = gamma1; s.gamma1 = gamma1n;
    s.sigma0 = sigma1; s.sigma1 = sigma1n;
    s.it = st->it + 1;
    s.done = (fabs(etan) <= s.norm_goal) ? 1 : (s.it >= maxiter ? 2 : 0);
    *st_next = s;
  }
}

// start of a solve: eta = beta = sqrt(<u1, v1>), goal = max(rtol*eta, atol); v1 /= beta, u1 /= beta.
__global__ void __launch_bounds__(BLK) k_minres_init(long n, double* __restrict__ v1, double* __restrict__ u1,
                                                     const double* bb_p, MinresState* st0, double rtol, double atol) {
  const double beta = sqrt(*bb_p);
  const double goal = fmax(rtol * beta, atol);
  const bool done = beta <= goal;
  if (!done) {
    const double m = 1.0 / beta;
    for (long i = (long)blockIdx.x * BLK + threadIdx.x; i < n; i += (long)gridDim.x * BLK) {
      v1[i] *= m;
      u1[i] *= m;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    MinresState s;
    s.beta = beta; s.eta = beta; s.eta0 = beta;
    s.gamma0 = 1.; s.gamma1 = 1.; s.sigma0 = 0.; s.sigma1 = 0.;
    s.norm_goal = goal;
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
};

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

template <bool RATE>
__global__ void __launch_bounds__(BLK) k_sat_substep(Lay L, Phys ph, SatParam sp, const double* __restrict__ u,
                                                     const double* __restrict__ phi, const double* __restrict__ Sin,
                                                     double* __restrict__ Sout) {
  const CellIdx ci = owned_cell(L);
  if (!ci.ok) return;
  const int i = ci.i, j = ci.j, kl = ci.kl, k = ci.k, kk = kl + 1;
  const long c = ix_c(L, i, j, kk);
  const bool xl = i > 0, xh = i + 1 < L.NX, yl = L.has_y && j > 0, yh = L.has_y && j + 1 < L.NY, zl = k > 0,
             zh = k + 1 < L.NZ;
  // face fluxes in +d direction (RT0 dof = total flux through the face)
  const double Fxl = xl ? u[ix_ux(L, i, j, kl)] : 0.0, Fxh = xh ? u[ix_ux(L, i + 1, j, kl)] : 0.0;
  const double Fyl = yl ? u[ix_uy(L, i, j, kl)] : 0.0, Fyh = yh ? u[ix_uy(L, i, j + 1, kl)] : 0.0;
  const double Fzl = zl ? u[ix_uz(L, i, j, kk)] : 0.0, Fzh = zh ? u[ix_uz(L, i, j, kk + 1)] : 0.0;
#define FPOS(F) ((F) > 0.0 ? (F) : 0.0)
#define FNEG(F) ((F) < 0.0 ? (F) : 0.0)
#define GS(idx) (sp.two_phase ? frac_flow_of(ph, Sin[idx]) : Sin[idx])
  double z = 0.0;
  if (zl) z += FPOS(Fzl) * GS(c - L.plane);
  if (yl) z += FPOS(Fyl) * GS(c - L.NX);
  if (xl) z += FPOS(Fxl) * GS(c - 1);
  {
    double dg = 0.0;  // diagonal of Kt, accumulated in assembly (face-loop) order
    if (xl) dg += FNEG(Fxl);
    if (xh) dg += -FPOS(Fxh);
    if (yl) dg += FNEG(Fyl);
    if (yh) dg += -FPOS(Fyh);
    if (zl) dg += FNEG(Fzl);
    if (zh) dg += -FPOS(Fzh);
    z += dg * GS(c);
  }
  if (xh) z += (-FNEG(Fxh)) * GS(c + 1);
  if (yh) z += (-FNEG(Fyh)) * GS(c + L.NX);
  if (zh) z += (-FNEG(Fzh)) * GS(c + L.plane);
#undef FPOS
#undef FNEG
#undef GS
  z += sat_source(sp, i, j, k) * sp.vol;  // z += b
  double yv = z / (phi[c] * sp.vol);       // M_phi^-1 (diagonal)
  if (RATE) {
    Sout[c] = yv;
  } else {
    yv *= sp.dt;
    Sout[c] = Sin[c] + yv;
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
__global__ void __launch_bounds__(BLK) k_guards(Lay L, const double* __restrict__ S, unsigned long long* n_bad,
                                                unsigned long long* max_bits) {
  const CellIdx ci = owned_cell(L);
  if (!ci.ok) return;
  const double s = S[ix_c(L, ci.i, ci.j, ci.kl + 1)];
  if (!isfinite(s)) {
    atomicAdd(n_bad, 1ull);
    return;
  }
  // order-preserving map double -> uint64 so that atomicMax works for negative values too
  unsigned long long b = (unsigned long long)__double_as_longlong(s);
  b = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
  atomicMax(max_bits, b);
}

// L2 flush helper for impes_time_kernel
__global__ void k_fill(double* p, long n, double v) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace impes
