// mg.cuh -- throughput-mode Schur-block preconditioner (IMPES_SCHUR_MG): one geometric multigrid V(nu,nu)
// cycle on S = B diag(M)^-1 B^T, the same matrix the serial reference smooths with one symmetric Gauss-Seidel
// sweep [REF src/pressure_solver.cpp:76-91] and the MPI reference hands to BoomerAMG
// [REF src/par_pressure_solver.cpp:83-90].  BoomerAMG is neither available nor restatable, so this
// mode is checked against the oracle's restatement of THIS algorithm (the CPU checker under tests), bit for bit.
//
// S is a two-point-flux stencil: S[c,c'] = -T_f, S[c,c] = sum_f T_f.  Hierarchy: 2x2(x2) cell aggregates;
// coarse transmissibility = sum of the fine ones across the aggregate interface (= Galerkin R S P with
// piecewise-constant P, R = P^T); damped-Jacobi smoothing (order independent => deterministic and symmetric);
// coarse-grid correction scaled by `cfac` (piecewise-constant Galerkin operators are too stiff by ~2x).
// MINRES iterations become mesh independent (~40-45 at rtol 1e-6 on the log-normal benchmark fields, against
// 800-2900 with the reference's preconditioner).
#pragma once
#include <cuda_runtime.h>

namespace impes {

struct MgLevel {
  int nx, ny, nz;              // cells held by this rank (2-D problems: ny = 1)
  int zoff, nzg;               // global index of local plane 0 and global plane count (one GPU: 0, nz).
                               // Distributed levels carry one ghost plane below / above in r, ya, yb.
  double *Tx, *Ty, *Tz, *D;    // compact face / cell arrays: Tx[(nx+1)*ny*nz], Ty[nx*(ny+1)*nz], Tz[nx*ny*(nz+1)]
  double *r, *ya, *yb;         // right-hand side and two iterate buffers (level 0: r and yb alias Krylov vectors)
};

constexpr int MG_B = 256;

struct MgIdx { int i, j, k; bool ok; };
__device__ __forceinline__ MgIdx mg_cell(const MgLevel& v) {
  const unsigned t = blockIdx.x * (unsigned)MG_B + threadIdx.x;
  const unsigned n = (unsigned)v.nx * v.ny * v.nz;
  MgIdx c;
  c.ok = t < n;
  const unsigned tt = c.ok ? t : 0u;
  const unsigned r = tt / (unsigned)v.nx;
  c.i = (int)(tt - r * (unsigned)v.nx);
  c.k = (int)(r / (unsigned)v.ny);
  c.j = (int)(r - (unsigned)c.k * (unsigned)v.ny);
  return c;
}

// (S y)_c in the fixed order: D*y, then minus T*y of the neighbours z-lo, y-lo, x-lo, x-hi, y-hi, z-hi
__device__ __forceinline__ double mg_apply(const MgLevel& v, const double* __restrict__ y, int i, int j, int k, int c) {
  const int pl = v.nx * v.ny;
  double s = v.D[c] * y[c];
  if (k + v.zoff > 0) s -= v.Tz[c] * y[c - pl];
  if (j > 0) s -= v.Ty[i + v.nx * (j + (v.ny + 1) * k)] * y[c - v.nx];
  if (i > 0) s -= v.Tx[i + (v.nx + 1) * (j + v.ny * k)] * y[c - 1];
  if (i + 1 < v.nx) s -= v.Tx[i + 1 + (v.nx + 1) * (j + v.ny * k)] * y[c + 1];
  if (j + 1 < v.ny) s -= v.Ty[i + v.nx * (j + 1 + (v.ny + 1) * k)] * y[c + v.nx];
  if (k + v.zoff + 1 < v.nzg) s -= v.Tz[c + pl] * y[c + pl];
  return s;
}

// level 0: compact copy of the fine stencil from the solver's face arrays (local planes of the slab)
__global__ void __launch_bounds__(MG_B) k_mg_copy_fine(Lay L, MgLevel v, const double* __restrict__ T,
                                                       const double* __restrict__ Md) {
  const MgIdx ci = mg_cell(v);
  if (!ci.ok) return;
  const int i = ci.i, j = ci.j, k = ci.k, c = i + v.nx * (j + v.ny * k), pl = v.nx * v.ny;
  v.D[c] = Md[ix_p(L, i, j, k + 1)];
  v.Tx[i + (v.nx + 1) * (j + v.ny * k)] = T[ix_ux(L, i, j, k)];
  if (i + 1 == v.nx) v.Tx[i + 1 + (v.nx + 1) * (j + v.ny * k)] = 0.0;
  if (L.has_y) {
    v.Ty[i + v.nx * (j + (v.ny + 1) * k)] = T[ix_uy(L, i, j, k)];
    if (j + 1 == v.ny) v.Ty[i + v.nx * (j + 1 + (v.ny + 1) * k)] = 0.0;
  }
  v.Tz[c] = T[ix_uz(L, i, j, k + 1)];
  if (k + 1 == v.nz) v.Tz[c + pl] = (k + v.zoff + 1 == v.nzg) ? 0.0 : T[ix_uz(L, i, j, k + 2)];  // slab top: ghost face
}

// coarse transmissibilities: sum of the fine ones across the aggregate interface (k, then j, then i ascending)
__global__ void __launch_bounds__(MG_B) k_mg_coarsen(MgLevel f, MgLevel c) {
  const MgIdx ci = mg_cell(c);
  if (!ci.ok) return;
  const int I = ci.i, J = ci.j, K = ci.k;
  const int i0 = 2 * I, j0 = f.ny > 1 ? 2 * J : 0, k0 = 2 * K;
  const int i1 = min(i0 + 2, f.nx), j1 = f.ny > 1 ? min(j0 + 2, f.ny) : 1, k1 = min(k0 + 2, f.nz);
  double tx = 0.0, ty = 0.0, tz = 0.0;
  if (I > 0)
    for (int k = k0; k < k1; ++k)
      for (int j = j0; j < j1; ++j) tx += f.Tx[i0 + (f.nx + 1) * (j + f.ny * k)];
  if (J > 0)
    for (int k = k0; k < k1; ++k)
      for (int i = i0; i < i1; ++i) ty += f.Ty[i + f.nx * (j0 + (f.ny + 1) * k)];
  if (K + c.zoff > 0)
    for (int j = j0; j < j1; ++j)
      for (int i = i0; i < i1; ++i) tz += f.Tz[i + f.nx * (j + f.ny * k0)];
  const int cc = I + c.nx * (J + c.ny * K);
  c.Tx[I + (c.nx + 1) * (J + c.ny * K)] = tx;
  if (I + 1 == c.nx) c.Tx[I + 1 + (c.nx + 1) * (J + c.ny * K)] = 0.0;
  if (c.ny > 1) {
    c.Ty[I + c.nx * (J + (c.ny + 1) * K)] = ty;
    if (J + 1 == c.ny) c.Ty[I + c.nx * (J + 1 + (c.ny + 1) * K)] = 0.0;
  }
  c.Tz[cc] = tz;
  if (K + 1 == c.nz) {  // upper face plane of the slab: the upper rank's lowest faces (aggregates are slab aligned)
    double tt = 0.0;
    if (K + c.zoff + 1 < c.nzg)
      for (int j = j0; j < j1; ++j)
        for (int i = i0; i < i1; ++i) tt += f.Tz[i + f.nx * (j + f.ny * (k0 + 2))];
    c.Tz[cc + c.nx * c.ny] = tt;
  }
}
// coarse diagonal: sum of the level's own face transmissibilities, order x-lo, x-hi, y-lo, y-hi, z-lo, z-hi
__global__ void __launch_bounds__(MG_B) k_mg_diag(MgLevel v) {
  const MgIdx ci = mg_cell(v);
  if (!ci.ok) return;
  const int i = ci.i, j = ci.j, k = ci.k, c = i + v.nx * (j + v.ny * k);
  double d = 0.0;
  d += v.Tx[i + (v.nx + 1) * (j + v.ny * k)];
  d += v.Tx[i + 1 + (v.nx + 1) * (j + v.ny * k)];
  if (v.ny > 1) {
    d += v.Ty[i + v.nx * (j + (v.ny + 1) * k)];
    d += v.Ty[i + v.nx * (j + 1 + (v.ny + 1) * k)];
  }
  d += v.Tz[c];
  d += v.Tz[c + v.nx * v.ny];
  v.D[c] = d;
}

// damped Jacobi: first = 1: y_out = omega * (r / D)   (zero initial guess)
//                else     : y_out = y_in + omega * ((r - S y_in) / D)
__global__ void __launch_bounds__(MG_B) k_mg_jacobi(MgLevel v, const double* __restrict__ r, const double* __restrict__ yin,
                                                    double* __restrict__ yout, double omega, int first,
                                                    const MinresState* st) {
  if (st && st->done) return;
  const MgIdx ci = mg_cell(v);
  if (!ci.ok) return;
  const int c = ci.i + v.nx * (ci.j + v.ny * ci.k);
  if (first) {
    yout[c] = omega * (r[c] / v.D[c]);
  } else {
    const double s = mg_apply(v, yin, ci.i, ci.j, ci.k, c);
    yout[c] = yin[c] + omega * ((r[c] - s) / v.D[c]);
  }
}

// coarse rhs = sum over the aggregate's children (k, j, i ascending) of the fine residual r - S y
__global__ void __launch_bounds__(MG_B) k_mg_restrict(MgLevel f, MgLevel c, const double* __restrict__ rf,
                                                      const double* __restrict__ yf, double* __restrict__ rc,
                                                      const MinresState* st) {
  if (st && st->done) return;
  const MgIdx ci = mg_cell(c);
  if (!ci.ok) return;
  const int i0 = 2 * ci.i, j0 = f.ny > 1 ? 2 * ci.j : 0, k0 = 2 * ci.k;
  const int i1 = min(i0 + 2, f.nx), j1 = f.ny > 1 ? min(j0 + 2, f.ny) : 1, k1 = min(k0 + 2, f.nz);
  double acc = 0.0;
  for (int k = k0; k < k1; ++k)
    for (int j = j0; j < j1; ++j)
      for (int i = i0; i < i1; ++i) {
        const int cf = i + f.nx * (j + f.ny * k);
        acc += rf[cf] - mg_apply(f, yf, i, j, k, cf);
      }
  rc[ci.i + c.nx * (ci.j + c.ny * ci.k)] = acc;
}

// y_fine += cfac * e_coarse(parent)
__global__ void __launch_bounds__(MG_B) k_mg_prolong(MgLevel f, MgLevel c, const double* __restrict__ ec,
                                                     double* __restrict__ yf, double cfac, const MinresState* st) {
  if (st && st->done) return;
  const MgIdx ci = mg_cell(f);
  if (!ci.ok) return;
  const int cf = ci.i + f.nx * (ci.j + f.ny * ci.k);
  const int cc = (ci.i >> 1) + c.nx * ((f.ny > 1 ? ci.j >> 1 : 0) + c.ny * (ci.k >> 1));
  yf[cf] = yf[cf] + cfac * ec[cc];
}

}  // namespace impes
