// layout.h -- device data layout of the IMPES hot path (RT0 / P0 / P0 on a Cartesian grid).
//
// Internal grid: always 3-D (NX, NY, NZ).  A 2-D problem (nx, ny) is stored as (nx, 1, ny):
// the 2-D y-direction plays the role of internal z and the internal y-component is absent
// (has_y = 0).  Native dof order [ux | uy | uz | p] of include/impes_b200.h is preserved by
// this mapping, so host<->device transfers are per-component contiguous copies.
//
// Slab decomposition (one rank per GPU) along internal z.  A rank owns cell planes
// [zlo, zhi) and, by the rule "a cell owns its lower faces", the z-face planes [zlo, zhi)
// (the last rank also owns the top boundary plane NZ).
//
// One Darcy-sized device vector ("dvec") holds, each component padded to 32 doubles:
//   ux : nzl planes of (NX+1)*NY        owned only
//   uy : nzl planes of NX*(NY+1)        owned only          (absent when !has_y)
//   uz : (nzl+2) planes of NX*NY        face planes zlo-1 .. zhi   (ghost below / above)
//   p  : (nzl+2) planes of NX*NY        cell planes zlo-1 .. zhi   (ghost below / above)
// Cell-centred coefficient arrays (mobility-weighted m_d, saturation, ...) use the p layout
// ("cvec", (nzl+2) planes).
#pragma once
#include <cstdint>

namespace impes {

// Exact unsigned division by a run-time constant through one 64-bit multiply and a shift (Granlund-Montgomery
// round-up method): q = (n * M) >> s with s = 32 + ceil(log2 d), M = ceil(2^s / d); exact for n < 2^31.
// Replaces the ~20-instruction 32-bit division sequence in every thread's cell-index computation.
struct FastDiv {
  unsigned long long M;
  int s;
};
inline FastDiv make_fastdiv(unsigned d) {
  FastDiv f;
  int L = 0;
  while ((1ull << L) < d) ++L;
  f.s = 32 + L;
  f.M = ((1ull << f.s) + d - 1) / d;
  return f;
}

struct Lay {
  FastDiv fdx, fdy;    // division by NX, NY
  int NX, NY, NZ;      // internal global grid
  int has_y;           // 0 for 2-D problems
  int zlo, zhi, nzl;   // owned cell planes [zlo, zhi)
  int rank, nranks;
  // all element offsets fit in 32 bits (a dvec per rank is < 2^31 doubles = 16 GiB; checked at create)
  int plane;           // NX*NY
  int off_x, off_y, off_z, off_p;  // component offsets inside a dvec
  int ntot;            // padded dvec length
  int ncl;             // owned cells = nzl*plane
  int ncv;             // cvec length = (nzl+2)*plane
  int nblk;            // blocks of 256 owned cells (canonical reduction granularity)
};

// Physics needed by device code: Param::Krw/Kro [REF src/param.cpp:195-215], MU_W/MU_O
// [REF config.hpp.in:37,43].
struct Phys {
  int two_phase;
  double Sor, Swc, mu_w, mu_o;
  // constant denominators of the rel-perm / fractional-flow formulas and their correctly rounded
  // reciprocals (for div_exact): den = 1 - Sor - Swc
  double den, rden, rmu_w, rmu_o;
  int unit_den;  // Swc == 0 and den == 1: (S - Swc)/den is S exactly
  int unit_mu;   // mu_w == mu_o == 1 (the reference's compile-time values): x/mu is x exactly
};

#if defined(__CUDACC__)
#define IMPES_HD __host__ __device__ __forceinline__
#else
#define IMPES_HD inline
#endif

IMPES_HD unsigned fastdiv(unsigned n, const FastDiv& f) { return (unsigned)(((unsigned long long)n * f.M) >> f.s); }

IMPES_HD int ix_ux(const Lay& L, int i, int j, int kl) { return L.off_x + i + (L.NX + 1) * (j + L.NY * kl); }
IMPES_HD int ix_uy(const Lay& L, int i, int j, int kl) { return L.off_y + i + L.NX * (j + (L.NY + 1) * kl); }
// kk = (face plane) - zlo + 1
IMPES_HD int ix_uz(const Lay& L, int i, int j, int kk) { return L.off_z + i + L.NX * (j + L.NY * kk); }
// kk = (cell plane) - zlo + 1
IMPES_HD int ix_p(const Lay& L, int i, int j, int kk) { return L.off_p + i + L.NX * (j + L.NY * kk); }
IMPES_HD int ix_c(const Lay& L, int i, int j, int kk) { return i + L.NX * (j + L.NY * kk); }

// MINRES scalar recurrence state, device resident (double buffered by iteration parity).
// Lazy scaling of the Lanczos vectors v: they are stored UNscaled and k_minres_b, their only consumer, multiplies
// on load by m1 = 1/beta of the current v1 or m0 = 1/beta of the previous one (v0).  The product is the double the
// explicit scaling used to store, so nothing changes numerically; one read + one write per dof and iteration are
// gone.  (u = P v stays explicitly scaled: scaling K1's many input loads instead costs K1 ~30 %, measured.)
struct MinresState {
  double beta, eta, gamma0, gamma1, sigma0, sigma1, norm_goal, eta0;
  double m0, m1;
  double mu;  // lazy scaling of u = P v as well (DG-Q1 path only, 1 otherwise): K1 and pass C multiply the stored,
              // UNscaled vector by mu on load -- the same rounded products the explicit scaling stored; pass C then
              // neither reads nor writes q (a quarter of its traffic).  Round 1 rejected this for the RT0 K1 (its 13
              // loads per cell became 30 % slower); the DG-Q1 K1 does ~1000 fp64 operations per thread and does not notice.
  int it;     // iterations completed
  int done;   // 0 running, 1 converged, 2 hit maxiter
};

}  // namespace impes
