/*
 * impes_oracle.c -- CPU restatement of the reference's per-time-step IMPES hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (incompressible-flow_b200/,
 * include/) may import, link or call this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker and as
 * the CPU baseline -- never as the thing shipped.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 * (SURVEY 8c), and its arithmetic lives in un-vendored, un-pinned MFEM (~3.0-3.3) /
 * HYPRE, which cannot be built in this image (no mfem.hpp, no MPI, no network).
 * This file therefore restates the published MFEM algorithms at the reference's own
 * call sites; every MFEM convention is tagged [MFEM-recall] and is unverifiable here.
 * What pins it instead: the analytic invariants of SURVEY Appendix C (tests/).
 *
 * Deliberately written the way the reference works, NOT the way the GPU path works:
 * every operator is ASSEMBLED from element matrices computed by numerical quadrature
 * (like MFEM's integrators) into CSR matrices, re-assembled on every call
 * [REF src/pressure_solver.cpp:47-81] [REF src/saturation_solver.cpp:30-55], so it
 * is an independent formulation of the operators the GPU applies matrix-free.
 *
 * Dof numbering: the "native" layout documented in include/impes_b200.h (the math
 * is invariant under MFEM's permutation/sign convention).
 *
 * Covered: RT_k / L2_k, k <= 1, with the saturation in DG_0 or DG_1 (Q1 at the Gauss nodes);
 * DG_1 saturation operator = Convection(v,1)^T - DGTrace(v,1,1/2) with MFEM's rule orders.
 * arith == 1 ("canonical") pins round-off-level choices so that the GPU path can be compared
 * bit for bit: closed-form RT0 / RT1 (cell-constant coefficient) element matrices, closed-form
 * RT1/Q1 divergence matrix and right-hand side, the GPU's dot-product tree.  The DG_1-coefficient
 * mass matrix stays quadrature-assembled in both modes (the GPU agrees to round-off there).
 */
#include "../include/impes_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

typedef struct {
  long nr, nc, nnz;
  long *I;
  int *J;
  double *A;
} csr;

static void csr_free(csr *m) {
  free(m->I); free(m->J); free(m->A);
  memset(m, 0, sizeof(*m));
}

typedef struct { int nq; double x[8], w[8]; } qrule;

typedef struct orc {
  impes_desc d;
  int dim, n[3], kv, kp, ks;
  double h[3], vol;
  long ncell, nv, np, ns, nvc[3], voff[3];
  int vg[3][3];            /* node grid of velocity component d along e */
  int npl, nsl;            /* local pressure / saturation dofs per cell */
  double *K[3], *phi;
  double cnode[4], onode[3], pnode[3], snode[3]; /* closed / open / L2_p / DG_s nodes on [0,1] */
  unsigned char *ess;      /* essential (boundary) velocity dof flags */
  csr B, BT;               /* constant: B = -(div u, q), boundary columns eliminated */
  double *rhs;             /* [0 | -(Q,psi)] */
  /* rebuilt per pressure solve */
  csr M, S;
  double *Md;
  int have_system;
  int dot_mode;            /* 0 = sequential (MFEM), 1 = pairwise (drift experiments) */
  int arith;               /* 0 = MFEM-like (quadrature element matrices, sequential dots, hypot);
                              1 = canonical: the arithmetic contract shared with the GPU path (closed-form RT0
                                  element matrices, the GPU's fixed dot-product tree, rho1 = sqrt(d^2+b^2)).
                                  Same algorithm, same assembled-CSR data flow; only round-off-level choices
                                  are pinned so that GPU and oracle agree bit for bit. */
  /* canonical RT1/Q1 constants (arith == 1): CW = Mc (x) Mo (x) Mo entries, Bv = -G (x) Mo (x) Mo */
  double c_wt, c_CW[3][3], c_Bv[2][3];
  /* canonical DG-Q1 path (arith == 1, kv = kp = ks = 1): w_q / lambda(S(x_q)) per cell and the dot of the last
   * canonical K1 application (orc_q1.inc mirrors csrc/rt1q.cuh operation by operation) */
  double *rlam;
  double q1_alpha;
  int threads;             /* >1: OpenMP in SpMV / vector ops (CPU baseline only) */
  int det;                 /* threads > 1 with deterministic arithmetic: the canonical paths keep their bits */
  int mass_cg;             /* 1: invert M_phi with Jacobi-PCG like the reference */
  /* MG hierarchy for IMPES_SCHUR_MG (restated in orc_mg.inc) */
  void *mg;
  char err[256];
} orc;

/* ---------------------------------------------------------------- utilities */

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* Gauss-Legendre rule with n points on [0,1]  [MFEM-recall: IntRules.Get(Segment, order), n = order/2+1] */
static void gauss01(int n, qrule *q) {
  q->nq = n;
  for (int i = 0; i < n; ++i) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= n; ++j) {
        double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    /* z descends from +1; map to ascending points on [0,1] */
    q->x[n - 1 - i] = 0.5 * (1.0 + z);
    q->w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
  }
}

static double lag(const double *nodes, int nn, int a, double x) {
  double v = 1.0;
  for (int m = 0; m < nn; ++m)
    if (m != a) v *= (x - nodes[m]) / (nodes[a] - nodes[m]);
  return v;
}

static double dlag(const double *nodes, int nn, int a, double x) {
  double s = 0.0;
  for (int l = 0; l < nn; ++l) {
    if (l == a) continue;
    double t = 1.0 / (nodes[a] - nodes[l]);
    for (int m = 0; m < nn; ++m)
      if (m != a && m != l) t *= (x - nodes[m]) / (nodes[a] - nodes[m]);
    s += t;
  }
  return s;
}

/* COO -> CSR: rows grouped (stable), columns sorted ascending (stable), duplicates summed
 * in insertion order -- deterministic stand-in for MFEM's element-by-element AddSubMatrix. */
static int coo_to_csr(long nr, long nc, long nnz, const long *ri, const int *ci, const double *v, csr *out) {
  long *I = (long *)calloc(nr + 1, sizeof(long));
  int *J = (int *)malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
  double *A = (double *)malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
  if (!I || !J || !A) return -1;
  for (long e = 0; e < nnz; ++e) I[ri[e] + 1]++;
  for (long r = 0; r < nr; ++r) I[r + 1] += I[r];
  long *pos = (long *)malloc(sizeof(long) * (nr > 0 ? nr : 1));
  memcpy(pos, I, sizeof(long) * nr);
  for (long e = 0; e < nnz; ++e) {
    long p = pos[ri[e]]++;
    J[p] = ci[e];
    A[p] = v[e];
  }
  free(pos);
  long w = 0;
  long rs = 0;
  for (long r = 0; r < nr; ++r) {
    long re = I[r + 1];
    /* stable insertion sort of [rs,re) by column */
    for (long a = rs + 1; a < re; ++a) {
      int cj = J[a];
      double cv = A[a];
      long b = a - 1;
      while (b >= rs && J[b] > cj) { J[b + 1] = J[b]; A[b + 1] = A[b]; --b; }
      J[b + 1] = cj;
      A[b + 1] = cv;
    }
    long wstart = w;
    for (long a = rs; a < re; ++a) {
      if (w > wstart && J[w - 1] == J[a]) A[w - 1] += A[a];
      else { J[w] = J[a]; A[w] = A[a]; ++w; }
    }
    rs = re;
    I[r + 1] = w;
  }
  out->nr = nr; out->nc = nc; out->nnz = w; out->I = I; out->J = J; out->A = A;
  return 0;
}

static int csr_transpose(const csr *a, csr *t) {
  long *I = (long *)calloc(a->nc + 1, sizeof(long));
  int *J = (int *)malloc(sizeof(int) * (a->nnz > 0 ? a->nnz : 1));
  double *A = (double *)malloc(sizeof(double) * (a->nnz > 0 ? a->nnz : 1));
  if (!I || !J || !A) return -1;
  for (long e = 0; e < a->nnz; ++e) I[a->J[e] + 1]++;
  for (long c = 0; c < a->nc; ++c) I[c + 1] += I[c];
  long *pos = (long *)malloc(sizeof(long) * (a->nc > 0 ? a->nc : 1));
  memcpy(pos, I, sizeof(long) * a->nc);
  for (long r = 0; r < a->nr; ++r)
    for (long e = a->I[r]; e < a->I[r + 1]; ++e) {
      long p = pos[a->J[e]]++;
      J[p] = (int)r;
      A[p] = a->A[e];
    }
  free(pos);
  t->nr = a->nc; t->nc = a->nr; t->nnz = a->nnz; t->I = I; t->J = J; t->A = A;
  return 0;
}

/* y (+)= A x : SparseMatrix::AddMult row loop [MFEM-recall] */
static void csr_mult(const orc *o, const csr *a, const double *x, double *y, int add) {
  const long nr = a->nr;
#pragma omp parallel for schedule(static) if (o->threads > 1)
  for (long r = 0; r < nr; ++r) {
    double s = 0.0;
    for (long e = a->I[r]; e < a->I[r + 1]; ++e) s += a->A[e] * x[a->J[e]];
    y[r] = add ? y[r] + s : s;
  }
}

static double pairwise_dot(const double *a, const double *b, long n) {
  if (n <= 64) {
    double s = 0.0;
    for (long i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
  }
  long h = n / 2;
  return pairwise_dot(a, b, h) + pairwise_dot(a + h, b + h, n - h);
}

/* The GPU's reduction tree (csrc/kernels.cuh block_tree_256): xor-butterfly 16,8,4,2,1 inside each warp of 32,
 * then the 8 warp sums added in order. */
static double tree256(const double *v) {
  double ws[8];
  for (int w = 0; w < 8; ++w) {
    double a[32], t[32];
    memcpy(a, v + 32 * w, sizeof a);
    for (int off = 16; off > 0; off >>= 1) {
      for (int l = 0; l < 32; ++l) t[l] = a[l] + a[l ^ off];
      memcpy(a, t, sizeof a);
    }
    ws[w] = a[0];
  }
  double s = 0.0;
  for (int w = 0; w < 8; ++w) s += ws[w];
  return s;
}

/* Canonical dot over a Darcy vector: blocks of 256 consecutive cells; each cell adds the products of the
 * dofs it owns (x-lo face, x-hi if on the boundary, y.., z.., pressure) -- k_dot_owned / k_darcy_apply. */
static double canon_dot(const orc *o, const double *a, const double *b) {
  const int NX = o->n[0], NY = o->dim == 3 ? o->n[1] : 1, NZ = o->dim == 3 ? o->n[2] : o->n[1];
  const int has_y = o->dim == 3;
  const long nvx = (long)(NX + 1) * NY * NZ, nvy = has_y ? (long)NX * (NY + 1) * NZ : 0;
  const long ncl = (long)NX * NY * NZ, nblk = (ncl + 255) / 256;
  double *partial = (double *)malloc(sizeof(double) * nblk);
#pragma omp parallel for schedule(static) if (o->threads > 1)
  for (long blk = 0; blk < nblk; ++blk) {
    double v[256];
    for (int t = 0; t < 256; ++t) {
      const long c = blk * 256 + t;
      double acc = 0.0;
      if (c < ncl) {
        const int i = (int)(c % NX), j = (int)((c / NX) % NY), k = (int)(c / ((long)NX * NY));
        long f = i + (long)(NX + 1) * (j + (long)NY * k);
        acc += a[f] * b[f];
        if (i + 1 == NX) acc += a[f + 1] * b[f + 1];
        if (has_y) {
          f = nvx + i + (long)NX * (j + (long)(NY + 1) * k);
          acc += a[f] * b[f];
          if (j + 1 == NY) acc += a[f + NX] * b[f + NX];
        }
        f = nvx + nvy + i + (long)NX * (j + (long)NY * k);
        acc += a[f] * b[f];
        if (k + 1 == NZ) acc += a[f + (long)NX * NY] * b[f + (long)NX * NY];
        f = o->nv + c;
        acc += a[f] * b[f];
      }
      v[t] = acc;
    }
    partial[blk] = tree256(v);
  }
  double v[256];
  for (int t = 0; t < 256; ++t) {
    double acc = 0.0;
    for (long j = t; j < nblk; j += 256) acc += partial[j];
    v[t] = acc;
  }
  free(partial);
  return tree256(v);
}

/* Canonical dot for RT1/Q1 Darcy vectors (csrc/rt1.cuh k1_darcy_apply / k1_dot_owned): one "thread" per L2 dof,
 * blocks of 256 consecutive dofs (cell-major).  The thread of local index (a_0, a_1, a_2) adds, per direction
 * d = x, y, z, the product of closed node a_d (0 = lower face node, 1 = interior node) of its line through the
 * open nodes (a_e, e != d) -- and of node 2 when a_d = 1 on the upper boundary --, then its own pressure dof. */
static inline long vidx(const orc *o, int d, int i0, int i1, int i2);
static double canon_dot_rt1(const orc *o, const double *a, const double *b) {
  const int dim = o->dim, npl = o->npl;
  const long nd = o->ncell * npl, nblk = (nd + 255) / 256;
  double *partial = (double *)malloc(sizeof(double) * nblk);
#pragma omp parallel for schedule(static) if (o->threads > 1)
  for (long blk = 0; blk < nblk; ++blk) {
    double v[256];
    for (int t = 0; t < 256; ++t) {
      const long gt = blk * 256 + t;
      double acc = 0.0;
      if (gt < nd) {
        const long c = gt / npl;
        const int l = (int)(gt % npl);
        const int cc[3] = {(int)(c % o->n[0]), (int)((c / o->n[0]) % o->n[1]), (int)(c / ((long)o->n[0] * o->n[1]))};
        const int al[3] = {l & 1, (l >> 1) & 1, (l >> 2) & 1};
        for (int d = 0; d < dim; ++d) {
          int g[3] = {0, 0, 0};
          for (int e = 0; e < dim; ++e) g[e] = 2 * cc[e] + al[e];
          long f = vidx(o, d, g[0], g[1], g[2]);
          acc += a[f] * b[f];
          if (al[d] == 1 && cc[d] + 1 == o->n[d]) {
            g[d] = 2 * cc[d] + 2;
            f = vidx(o, d, g[0], g[1], g[2]);
            acc += a[f] * b[f];
          }
        }
        const long f = o->nv + gt;
        acc += a[f] * b[f];
      }
      v[t] = acc;
    }
    partial[blk] = tree256(v);
  }
  double v[256];
  for (int t = 0; t < 256; ++t) {
    double acc = 0.0;
    for (long j = t; j < nblk; j += 256) acc += partial[j];
    v[t] = acc;
  }
  free(partial);
  return tree256(v);
}

/* beta^2 = <v0, P v0> of the RT1/Q1 + multigrid path, where the GPU folds the dot into other kernels
 * (csrc/kernels.cuh k_minres_b_dot, csrc/rt1.cuh k1_mg_jacobi<.., DOT>):
 *   velocity part: the device vector [ux | pad | uy | pad | ghost, uz, ghost | pad] is cut into chunks of 16384 entries;
 *     "thread" t of a chunk adds a b of entries t, t + 256, ... in order, tree256 per chunk, then the ordered sum of the
 *     chunk partials (padding and ghost entries contribute nothing);
 *   pressure part: blocks of 256 consecutive L2 dofs, tree256 per block, ordered sum of the block partials;
 *   result = velocity part + pressure part. */
static double strided_tree(const double *partial, long n) {
  double v[256];
  for (int t = 0; t < 256; ++t) {
    double acc = 0.0;
    for (long j = t; j < n; j += 256) acc += partial[j];
    v[t] = acc;
  }
  return tree256(v);
}
static double canon_dot_fold(const orc *o, const double *a, const double *b) {
  const int HY = o->dim == 3;
  const long NX = o->n[0], NY = HY ? o->n[1] : 1, NZ = HY ? o->n[2] : o->n[1], oy = HY ? 2 : 1;
  const long nx_ = (2 * NX + 1) * (oy * NY) * 2 * NZ, ny_ = HY ? 2 * NX * (2 * NY + 1) * 2 * NZ : 0;
  const long pz = 2 * NX * oy * NY, nz_ = pz * (2 * NZ + 3);
#define PAD32(v) (((v) + 31) / 32 * 32)
  const long off_y = PAD32(nx_), off_z = PAD32(off_y + ny_), off_p = PAD32(off_z + nz_);
#undef PAD32
  const long own_z0 = off_z + 2 * pz, own_z1 = own_z0 + pz * (2 * NZ + 1);
  const long nvx = nx_, nvy = ny_;
  const long nch = (off_p + 16383) / 16384;
  double *partial = (double *)malloc(sizeof(double) * (nch > 0 ? nch : 1));
#pragma omp parallel for schedule(static) if (o->threads > 1)
  for (long c = 0; c < nch; ++c) {
    double v[256];
    for (int t = 0; t < 256; ++t) {
      double acc = 0.0;
      for (int k = 0; k < 64; ++k) {
        const long i = c * 16384 + t + 256L * k;
        long f = -1;
        if (i < nx_) f = i;
        else if (i >= off_y && i < off_y + ny_) f = nvx + (i - off_y);
        else if (i >= own_z0 && i < own_z1) f = nvx + nvy + (i - own_z0);
        if (i < off_p) acc += f >= 0 ? a[f] * b[f] : 0.0;
      }
      v[t] = acc;
    }
    partial[c] = tree256(v);
  }
  const double tu = strided_tree(partial, nch);
  free(partial);
  const long nb = (o->np + 255) / 256;
  partial = (double *)malloc(sizeof(double) * (nb > 0 ? nb : 1));
#pragma omp parallel for schedule(static) if (o->threads > 1)
  for (long blk = 0; blk < nb; ++blk) {
    double v[256];
    for (int t = 0; t < 256; ++t) {
      const long p = blk * 256 + t;
      v[t] = p < o->np ? a[o->nv + p] * b[o->nv + p] : 0.0;
    }
    partial[blk] = tree256(v);
  }
  const double tp = strided_tree(partial, nb);
  free(partial);
  return tu + tp;
}

/* Vector::operator* : plain sequential sum [MFEM-recall] */
static double vdot(const orc *o, const double *a, const double *b, long n) {
  const int canon = o->arith == 1 && (o->threads <= 1 || o->det) && n == o->nv + o->np;
  if (canon && o->kv == 0 && o->kp == 0) return canon_dot(o, a, b);
  if (canon && o->kv == 1 && o->kp == 1) return canon_dot_rt1(o, a, b);
  if (o->threads > 1) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (long i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
  }
  if (o->dot_mode == 1) return pairwise_dot(a, b, n);
  double s = 0.0;
  for (long i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* ---------------------------------------------------------------- physics */

/* Param::Krw / Param::Kro  [REF src/param.cpp:195-215]; no clamping of S */
static double krw(const impes_desc *d, double S) {
  if (d->two_phase) {
    double S_ = (S - d->Swc) / (1. - d->Sor - d->Swc);
    return S_ * S_;
  }
  return S;
}
static double kro(const impes_desc *d, double S) {
  if (d->two_phase) {
    double S_ = (S - d->Swc) / (1. - d->Sor - d->Swc);
    return (1. - S_) * (1. - S_);
  }
  return (1. - S);
}

/* WellFunctionCoefficient::eval2D/eval3D  [REF src/param.hpp:306-345] */
static double well_Q(const impes_desc *d, const double *p) {
  const impes_well *in = &d->injection, *pr = &d->production;
  if (d->dim == 2) {
    if (fabs(p[0] - in->center[0]) <= in->radius && fabs(p[1] - in->center[1]) <= in->radius) return d->inflow;
    if (fabs(p[0] - pr->center[0]) <= pr->radius && fabs(p[1] - pr->center[1]) <= pr->radius) return d->outflow;
    return 0.;
  }
  if (fabs(p[0] - in->center[0]) <= in->radius && fabs(p[2] - in->center[2]) <= in->radius &&
      p[1] >= in->center[1] && p[1] <= in->center[1] + in->height)
    return d->inflow;
  if (fabs(p[0] - pr->center[0]) <= pr->radius && fabs(p[2] - pr->center[2]) <= pr->radius &&
      p[1] >= pr->center[1] && p[1] <= pr->center[1] + pr->height)
    return d->outflow;
  return 0.;
}

/* SaturationSourceCoefficient::eval2D/eval3D  [REF src/param.hpp:382-406] */
static double sat_R(const impes_desc *d, const double *p) {
  const impes_well *in = &d->injection;
  if (d->dim == 2) {
    if (fabs(p[0] - in->center[0]) <= in->radius && fabs(p[1] - in->center[1]) <= in->radius) return d->saturation_source;
    return 0.;
  }
  if (fabs(p[0] - in->center[0]) <= in->radius && fabs(p[2] - in->center[2]) <= in->radius &&
      p[1] >= in->center[1] && p[1] <= in->center[1] + in->height)
    return d->saturation_source;
  return 0.;
}

static inline long vidx(const orc *o, int d, int i0, int i1, int i2) {
  return o->voff[d] + i0 + (long)o->vg[d][0] * (i1 + (long)o->vg[d][1] * i2);
}

/* ---------------------------------------------------------------- setup */

static int assemble_B_and_rhs(orc *o);

ORC_API void orc_destroy(orc *o);

ORC_API const char *orc_last_error(orc *o) { return o ? o->err : "null oracle handle"; }

ORC_API orc *orc_create(const impes_desc *d, const double *Kx, const double *Ky, const double *Kz, const double *phi) {
  if (!d || d->struct_size != (int)sizeof(impes_desc)) return NULL;
  if (d->dim != 2 && d->dim != 3) return NULL;
  if (d->order_v < 0 || d->order_v > 1 || d->order_p < 0 || d->order_p > 1 || d->order_s < 0 || d->order_s > 1) return NULL;
  orc *o = (orc *)calloc(1, sizeof(orc));
  o->d = *d;
  o->dim = d->dim;
  o->n[0] = d->nx; o->n[1] = d->ny; o->n[2] = d->dim == 3 ? d->nz : 1;
  o->h[0] = d->sx / d->nx; o->h[1] = d->sy / d->ny; o->h[2] = d->dim == 3 ? d->sz / d->nz : 1.0;
  o->vol = o->h[0] * o->h[1] * (d->dim == 3 ? o->h[2] : 1.0);
  o->kv = d->order_v; o->kp = d->order_p; o->ks = d->order_s;
  o->ncell = (long)o->n[0] * o->n[1] * o->n[2];
  o->npl = 1; o->nsl = 1;
  for (int e = 0; e < o->dim; ++e) { o->npl *= o->kp + 1; o->nsl *= o->ks + 1; }
  o->np = o->ncell * o->npl;
  o->ns = o->ncell * o->nsl;
  long off = 0;
  for (int c = 0; c < 3; ++c) {
    if (c >= o->dim) { o->nvc[c] = 0; o->voff[c] = off; o->vg[c][0] = o->vg[c][1] = o->vg[c][2] = 0; continue; }
    long sz = 1;
    for (int e = 0; e < 3; ++e) {
      int g = e >= o->dim ? 1 : (e == c ? o->n[e] * (o->kv + 1) + 1 : o->n[e] * (o->kv + 1));
      o->vg[c][e] = g;
      sz *= g;
    }
    o->nvc[c] = sz; o->voff[c] = off; off += sz;
  }
  o->nv = off;
  /* nodal points [MFEM-recall]: RT closed = Gauss-Lobatto (k+2), open = Gauss-Legendre (k+1);
   * L2/DG = Gauss-Legendre (k+1) */
  const double g1 = 0.5 / sqrt(3.0);
  if (o->kv == 0) { o->cnode[0] = 0; o->cnode[1] = 1; o->onode[0] = 0.5; }
  else { o->cnode[0] = 0; o->cnode[1] = 0.5; o->cnode[2] = 1; o->onode[0] = 0.5 - g1; o->onode[1] = 0.5 + g1; }
  if (o->kp == 0) o->pnode[0] = 0.5; else { o->pnode[0] = 0.5 - g1; o->pnode[1] = 0.5 + g1; }
  if (o->ks == 0) o->snode[0] = 0.5; else { o->snode[0] = 0.5 - g1; o->snode[1] = 0.5 + g1; }

  const double *Kin[3] = {Kx, Ky, Kz};
  for (int c = 0; c < o->dim; ++c) {
    if (!Kin[c]) { orc_destroy(o); return NULL; }
    o->K[c] = (double *)malloc(sizeof(double) * o->ncell);
    memcpy(o->K[c], Kin[c], sizeof(double) * o->ncell);
  }
  o->phi = (double *)malloc(sizeof(double) * o->ncell);
  if (phi) memcpy(o->phi, phi, sizeof(double) * o->ncell);
  else for (long i = 0; i < o->ncell; ++i) o->phi[i] = 1.0;

  /* all boundary attributes essential  [REF src/pressure_solver.cpp:56-59] */
  o->ess = (unsigned char *)calloc(o->nv, 1);
  for (int c = 0; c < o->dim; ++c)
    for (int i2 = 0; i2 < o->vg[c][2]; ++i2)
      for (int i1 = 0; i1 < o->vg[c][1]; ++i1)
        for (int i0 = 0; i0 < o->vg[c][0]; ++i0) {
          int id = c == 0 ? i0 : (c == 1 ? i1 : i2);
          if (id == 0 || id == o->vg[c][c] - 1) o->ess[vidx(o, c, i0, i1, i2)] = 1;
        }
  o->threads = 1;
  {
    static const int MCI[3][3] = {{4, 2, -1}, {2, 16, 2}, {-1, 2, 4}};
    const double r3 = sqrt(3.0) / 3.0;
    const double G[2][3] = {{-0.5 - r3, 2.0 * r3, 0.5 - r3}, {-0.5 + r3, -2.0 * r3, 0.5 + r3}};
    o->c_wt = o->dim == 3 ? 0.25 : 0.5;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) o->c_CW[i][j] = (MCI[i][j] / 30.0) * o->c_wt;
    for (int a2 = 0; a2 < 2; ++a2)
      for (int i = 0; i < 3; ++i) o->c_Bv[a2][i] = -(G[a2][i] * o->c_wt);
  }
  if (assemble_B_and_rhs(o)) { orc_destroy(o); return NULL; }
  return o;
}

extern void orc_mg_free(orc *o);

ORC_API void orc_destroy(orc *o) {
  if (!o) return;
  for (int c = 0; c < 3; ++c) free(o->K[c]);
  free(o->phi); free(o->ess); free(o->rhs); free(o->Md); free(o->rlam);
  csr_free(&o->B); csr_free(&o->BT); csr_free(&o->M); csr_free(&o->S);
  orc_mg_free(o);
  free(o);
}

ORC_API void orc_set_threads(orc *o, int t) {
  o->threads = t < 1 ? 1 : t;
#ifdef _OPENMP
  omp_set_num_threads(o->threads);
#endif
}
/* OpenMP threads for the canonical (arith == 1) paths as well: every loop that is threaded there runs over
 * independent unknowns / reduction blocks, so the results are bit-identical to one thread. */
ORC_API void orc_set_threads_det(orc *o, int t) {
  orc_set_threads(o, t);
  o->det = 1;
}
ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
ORC_API void orc_set_dot_mode(orc *o, int m) { o->dot_mode = m; }
ORC_API void orc_set_arith(orc *o, int m) {
  if (o->arith == m) return;
  o->arith = m;
  assemble_B_and_rhs(o); /* B and the rhs have canonical closed forms for RT1/Q1 */
  o->have_system = 0;
}
ORC_API void orc_set_mass_cg(orc *o, int on) { o->mass_cg = on; }
ORC_API void orc_get_dims(orc *o, impes_dims *out) {
  memset(out, 0, sizeof(*out));
  out->n_cells = o->ncell; out->n_v = o->nv; out->n_p = o->np; out->n_s = o->ns;
  for (int c = 0; c < 3; ++c) out->n_v_comp[c] = o->nvc[c];
  out->slab_lo = 0; out->slab_hi = o->n[o->dim - 1];
}

/* B = -(div u, q)  [REF src/pressure_solver.cpp:48,53-54,68], boundary trial columns eliminated [REF :61];
 * rhs block 1 = -(Q, psi)  [REF :41-45].  Both are S-independent; the reference rebuilds them every
 * call, the oracle's timed path does too (orc_pressure_solve calls this again when timing_faithful). */
static int assemble_B_and_rhs(orc *o) {
  const int kv = o->kv, kp = o->kp, dim = o->dim;
  /* VectorFEDivergenceIntegrator order = trial.order + test.order - 1 = (kv+1)+kp-1 [MFEM-recall] */
  int ord = (kv + 1) + kp - 1;
  if (ord < 0) ord = 0;
  qrule qd; gauss01(ord / 2 + 1, &qd);
  qrule one = {1, {0.5}, {1.0}};
  const qrule *Q[3] = {&qd, &qd, dim == 3 ? &qd : &one};
  int cntp[3], cntv[3];
  long cap = o->ncell * (long)o->npl * dim * (kv + 2) * (kv + 1) * (dim == 3 ? (kv + 1) : 1);
  long *ri = (long *)malloc(sizeof(long) * cap);
  int *ci = (int *)malloc(sizeof(int) * cap);
  double *vv = (double *)malloc(sizeof(double) * cap);
  long nnz = 0;
  for (int e = 0; e < 3; ++e) cntp[e] = e < dim ? kp + 1 : 1;
  /* element matrices are identical for all cells (Piola: mesh independent) -- compute per component once */
  for (int cz = 0; cz < o->n[2]; ++cz)
    for (int cy = 0; cy < o->n[1]; ++cy)
      for (int cx = 0; cx < o->n[0]; ++cx) {
        const int cc[3] = {cx, cy, cz};
        const long cell = cx + (long)o->n[0] * (cy + (long)o->n[1] * cz);
        for (int d = 0; d < dim; ++d) {
          for (int e = 0; e < 3; ++e) cntv[e] = e >= dim ? 1 : (e == d ? kv + 2 : kv + 1);
          for (int b2 = 0; b2 < cntp[2]; ++b2)
            for (int b1 = 0; b1 < cntp[1]; ++b1)
              for (int b0 = 0; b0 < cntp[0]; ++b0) {
                const int bb[3] = {b0, b1, b2};
                const long prow = cell * o->npl + b0 + (kp + 1) * (b1 + (kp + 1) * b2);
                for (int a2 = 0; a2 < cntv[2]; ++a2)
                  for (int a1 = 0; a1 < cntv[1]; ++a1)
                    for (int a0 = 0; a0 < cntv[0]; ++a0) {
                      const int aa[3] = {a0, a1, a2};
                      double s = 0.0;
                      if (o->arith == 1 && kv == 1 && kp == 1) {
                        /* canonical closed form: B|cell,d = -G(d) (x) Mo (x) Mo, Mo = diag(1/2) */
                        int match = 1;
                        for (int e = 0; e < dim; ++e) if (e != d && aa[e] != bb[e]) match = 0;
                        if (!match) continue;
                        int gi[3];
                        for (int e = 0; e < 3; ++e) gi[e] = e < dim ? cc[e] * (kv + 1) + aa[e] : 0;
                        const long col = vidx(o, d, gi[0], gi[1], gi[2]);
                        ri[nnz] = prow; ci[nnz] = (int)col; vv[nnz] = o->ess[col] ? 0.0 : o->c_Bv[bb[d]][aa[d]]; ++nnz;
                        continue;
                      }
                      for (int q2 = 0; q2 < Q[2]->nq; ++q2)
                        for (int q1 = 0; q1 < Q[1]->nq; ++q1)
                          for (int q0 = 0; q0 < Q[0]->nq; ++q0) {
                            const int qq[3] = {q0, q1, q2};
                            double w = 1.0, psi = 1.0, dphi = 1.0;
                            for (int e = 0; e < dim; ++e) {
                              const double xq = Q[e]->x[qq[e]];
                              w *= Q[e]->w[qq[e]];
                              psi *= lag(o->pnode, kp + 1, bb[e], xq);
                              dphi *= (e == d) ? dlag(o->cnode, kv + 2, aa[e], xq) : lag(o->onode, kv + 1, aa[e], xq);
                            }
                            s += w * psi * dphi;
                          }
                      int gi[3];
                      for (int e = 0; e < 3; ++e) gi[e] = e < dim ? cc[e] * (kv + 1) + aa[e] : 0;
                      const long col = vidx(o, d, gi[0], gi[1], gi[2]);
                      /* (div u, q) assembled, then B *= -1 [REF :68]; eliminated trial dofs -> 0 [REF :61] */
                      double val = -s;
                      if (o->ess[col]) val = 0.0;
                      ri[nnz] = prow; ci[nnz] = (int)col; vv[nnz] = val; ++nnz;
                    }
              }
        }
      }
  csr_free(&o->B); csr_free(&o->BT);
  if (coo_to_csr(o->np, o->nv, nnz, ri, ci, vv, &o->B)) return -1;
  free(ri); free(ci); free(vv);
  if (csr_transpose(&o->B, &o->BT)) return -1;

  /* rhs: block 0 = 0; block 1 = -(Q,psi), DomainLFIntegrator order 2*kp -> kp+1 points [MFEM-recall] */
  free(o->rhs);
  o->rhs = (double *)calloc(o->nv + o->np, sizeof(double));
  qrule qr; gauss01(kp + 1, &qr);
  const qrule *R[3] = {&qr, &qr, dim == 3 ? &qr : &one};
  for (int cz = 0; cz < o->n[2]; ++cz)
    for (int cy = 0; cy < o->n[1]; ++cy)
      for (int cx = 0; cx < o->n[0]; ++cx) {
        const int cc[3] = {cx, cy, cz};
        const long cell = cx + (long)o->n[0] * (cy + (long)o->n[1] * cz);
        for (int b2 = 0; b2 < cntp[2]; ++b2)
          for (int b1 = 0; b1 < cntp[1]; ++b1)
            for (int b0 = 0; b0 < cntp[0]; ++b0) {
              const int bb[3] = {b0, b1, b2};
              double s = 0.0;
              if (o->arith == 1 && kp == 1) {
                /* canonical: the 2-point Gauss rule coincides with the L2 nodes => g_m = -Q(x_m) |cell| / 2^dim */
                double p[3] = {0, 0, 0};
                for (int e = 0; e < dim; ++e) p[e] = (cc[e] + o->pnode[bb[e]]) * o->h[e];
                o->rhs[o->nv + cell * o->npl + b0 + (kp + 1) * (b1 + (kp + 1) * b2)] =
                    -(well_Q(&o->d, p) * (o->vol * (dim == 3 ? 0.125 : 0.25)));
                continue;
              }
              for (int q2 = 0; q2 < R[2]->nq; ++q2)
                for (int q1 = 0; q1 < R[1]->nq; ++q1)
                  for (int q0 = 0; q0 < R[0]->nq; ++q0) {
                    const int qq[3] = {q0, q1, q2};
                    double w = o->vol, psi = 1.0, p[3] = {0, 0, 0};
                    for (int e = 0; e < dim; ++e) {
                      const double xq = R[e]->x[qq[e]];
                      w *= R[e]->w[qq[e]];
                      psi *= lag(o->pnode, kp + 1, bb[e], xq);
                      p[e] = (cc[e] + xq) * o->h[e];
                    }
                    s += w * well_Q(&o->d, p) * psi;
                  }
              o->rhs[o->nv + cell * o->npl + b0 + (kp + 1) * (b1 + (kp + 1) * b2)] = -s; /* gform *= -1 [REF :45] */
            }
      }
  return 0;
}

/* Saturation at reference point xi of a cell (DG_ks nodal interpolation):
 * GridFunctionCoefficient::Eval [REF src/param.hpp:250] */
static double sat_at(const orc *o, const double *S, long cell, const double *xi) {
  if (o->ks == 0) return S[cell];
  double v = 0.0;
  const int n1 = o->ks + 1;
  for (int m2 = 0; m2 < (o->dim == 3 ? n1 : 1); ++m2)
    for (int m1 = 0; m1 < n1; ++m1)
      for (int m0 = 0; m0 < n1; ++m0) {
        double b = lag(o->snode, n1, m0, xi[0]) * lag(o->snode, n1, m1, xi[1]);
        if (o->dim == 3) b *= lag(o->snode, n1, m2, xi[2]);
        v += S[cell * o->nsl + m0 + n1 * (m1 + n1 * m2)] * b;
      }
  return v;
}

static int q1_canonical(const orc *o) { return o->arith == 1 && o->kv == 1 && o->kp == 1 && o->ks == 1 && (o->threads <= 1 || o->det); }
static void q1_coefficients(orc *o, const double *S);
static void q1_darcy_mult(orc *o, const double *x, double *y);
static int q1_saturation_rate(const orc *o, const double *u, const double *S, const double *F, double *out, double dt, int rate);

/* M = VectorFEMassIntegrator(CWVectorCoefficient)  [REF src/pressure_solver.cpp:47,50-51]
 * with boundary rows/cols eliminated, diagonal 1 [REF :60]; then diag(M) and
 * S = B diag(M)^-1 B^T  [REF :76-81]. */
static int assemble_system(orc *o, const double *S) {
  const int kv = o->kv, dim = o->dim;
  /* order = OrderW + 2*el.order = (dim-1) + 2(kv+1)  ->  n = order/2 + 1  [MFEM-recall] */
  const int ord = (dim - 1) + 2 * (kv + 1);
  qrule qm; gauss01(ord / 2 + 1, &qm);
  qrule one = {1, {0.5}, {1.0}};
  const qrule *Q[3] = {&qm, &qm, dim == 3 ? &qm : &one};
  const int nl = (kv + 2) * (kv + 1) * (dim == 3 ? (kv + 1) : 1);
  long cap = o->ncell * (long)dim * nl * nl;
  long *ri = (long *)malloc(sizeof(long) * cap);
  int *ci = (int *)malloc(sizeof(int) * cap);
  double *vv = (double *)malloc(sizeof(double) * cap);
  if (!ri || !ci || !vv) return -1;
  long nnz = 0;
  /* basis tables at the quadrature points */
  double C[8][4], O[8][3];
  for (int q = 0; q < qm.nq; ++q) {
    for (int a = 0; a < kv + 2; ++a) C[q][a] = lag(o->cnode, kv + 2, a, qm.x[q]);
    for (int a = 0; a < kv + 1; ++a) O[q][a] = lag(o->onode, kv + 1, a, qm.x[q]);
  }
  const int nqt = Q[0]->nq * Q[1]->nq * Q[2]->nq;
  /* Element loop.  Every cell writes its element-matrix entries into its own slot of the COO arrays (slot =
   * dim*nl*nl entries, the closed forms use fewer) and the slots are compacted in cell order afterwards, so the
   * COO sequence -- hence the CSR and every sum built from it -- is the same with one thread or many (OpenMP is
   * only for the CPU baseline: MFEM's own assembly is element-parallel-friendly too). */
  const long slot = (long)dim * nl * nl;
  long *cellcnt = (long *)malloc(sizeof(long) * o->ncell);
#pragma omp parallel if (o->threads > 1)
  {
  double *coef = (double *)malloc(sizeof(double) * nqt);
  long *gl = (long *)malloc(sizeof(long) * nl);
  double *phi = (double *)malloc(sizeof(double) * nl * nqt);
#pragma omp for schedule(static)
  for (long cell = 0; cell < o->ncell; ++cell) {
      {
        const int cx = (int)(cell % o->n[0]), cy = (int)((cell / o->n[0]) % o->n[1]), cz = (int)(cell / ((long)o->n[0] * o->n[1]));
        const int cc[3] = {cx, cy, cz};
        long nnz = cell * slot;
        /* total mobility at the quadrature points: CWVectorCoefficient::Eval [REF src/param.hpp:246-261] */
        int qi = 0;
        for (int q2 = 0; q2 < Q[2]->nq; ++q2)
          for (int q1 = 0; q1 < Q[1]->nq; ++q1)
            for (int q0 = 0; q0 < Q[0]->nq; ++q0, ++qi) {
              const double xi[3] = {Q[0]->x[q0], Q[1]->x[q1], Q[2]->x[q2]};
              const double s = sat_at(o, S, cell, xi);
              coef[qi] = krw(&o->d, s) / o->d.mu_w + kro(&o->d, s) / o->d.mu_o;
            }
        for (int d = 0; d < dim; ++d) {
          int cnt[3];
          for (int e = 0; e < 3; ++e) cnt[e] = e >= dim ? 1 : (e == d ? kv + 2 : kv + 1);
          int li = 0;
          for (int a2 = 0; a2 < cnt[2]; ++a2)
            for (int a1 = 0; a1 < cnt[1]; ++a1)
              for (int a0 = 0; a0 < cnt[0]; ++a0, ++li) {
                const int aa[3] = {a0, a1, a2};
                int gi[3];
                for (int e = 0; e < 3; ++e) gi[e] = e < dim ? cc[e] * (kv + 1) + aa[e] : 0;
                gl[li] = vidx(o, d, gi[0], gi[1], gi[2]);
                int q = 0;
                for (int q2 = 0; q2 < Q[2]->nq; ++q2)
                  for (int q1 = 0; q1 < Q[1]->nq; ++q1)
                    for (int q0 = 0; q0 < Q[0]->nq; ++q0, ++q) {
                      const int qq[3] = {q0, q1, q2};
                      double b = 1.0;
                      for (int e = 0; e < dim; ++e) b *= (e == d) ? C[qq[e]][aa[e]] : O[qq[e]][aa[e]];
                      phi[li * nqt + q] = b;
                    }
              }
          /* contravariant Piola: phys = J ref / detJ -> component d scaled by h_d/vol; dV = vol dxi */
          const double geo = o->h[d] * o->h[d] / o->vol;
          const double Kd = o->K[d][cell];
          const int closed_form = o->arith == 1 && kv == 0 && o->ks == 0;
          const int closed_form1 = o->arith == 1 && kv == 1 && o->ks == 0;
          for (int i = 0; i < li; ++i)
            for (int j = 0; j < li; ++j) {
              double s = 0.0;
              int q = 0;
              if (closed_form1) { /* exact RT1 element matrix, cell-constant coefficient: m Mc (x) Mo (x) Mo */
                int ai[3], aj[3], ti = i, tj = j, match = 1;
                for (int e = 0; e < 3; ++e) { ai[e] = ti % cnt[e]; ti /= cnt[e]; aj[e] = tj % cnt[e]; tj /= cnt[e]; }
                for (int e = 0; e < dim; ++e) if (e != d && ai[e] != aj[e]) match = 0;
                if (!match) continue;
                const double mm = (1. / (coef[0] * Kd)) * geo;
                const double val0 = mm * o->c_CW[ai[d]][aj[d]];
                const long r0 = gl[i], c0 = gl[j];
                ri[nnz] = r0; ci[nnz] = (int)c0; vv[nnz] = (o->ess[r0] || o->ess[c0]) ? 0.0 : val0; ++nnz;
                continue;
              }
              if (closed_form) { /* exact RT0 element matrix m [[1/3,1/6],[1/6,1/3]] */
                const double mm = (1. / (coef[0] * Kd)) * geo;
                const double val0 = (i == j) ? mm * (1.0 / 3.0) : mm * (1.0 / 6.0);
                const long r0 = gl[i], c0 = gl[j];
                ri[nnz] = r0; ci[nnz] = (int)c0; vv[nnz] = (o->ess[r0] || o->ess[c0]) ? 0.0 : val0; ++nnz;
                continue;
              }
              for (int q2 = 0; q2 < Q[2]->nq; ++q2)
                for (int q1 = 0; q1 < Q[1]->nq; ++q1)
                  for (int q0 = 0; q0 < Q[0]->nq; ++q0, ++q) {
                    const double w = Q[0]->w[q0] * Q[1]->w[q1] * Q[2]->w[q2];
                    const double a = 1. / (coef[q] * Kd); /* V(d) = 1./(val*val_array[index]) [REF src/param.hpp:257-260] */
                    s += w * a * geo * phi[i * nqt + q] * phi[j * nqt + q];
                  }
              double val = s;
              const long r = gl[i], c = gl[j];
              if (o->ess[r] || o->ess[c]) val = 0.0; /* EliminateRowCol */
              ri[nnz] = r; ci[nnz] = (int)c; vv[nnz] = val; ++nnz;
            }
        }
        cellcnt[cell] = nnz - cell * slot;
      }
  }
  free(coef); free(gl); free(phi);
  }
  for (long cell = 0; cell < o->ncell; ++cell) { /* compact the slots in cell order */
    const long from = cell * slot, cnt = cellcnt[cell];
    if (from != nnz) {
      memmove(ri + nnz, ri + from, sizeof(long) * cnt);
      memmove(ci + nnz, ci + from, sizeof(int) * cnt);
      memmove(vv + nnz, vv + from, sizeof(double) * cnt);
    }
    nnz += cnt;
  }
  free(cellcnt);
  csr_free(&o->M);
  if (coo_to_csr(o->nv, o->nv, nnz, ri, ci, vv, &o->M)) return -1;
  free(ri); free(ci); free(vv);
  /* diagonal of eliminated rows := 1  [MFEM-recall: default DIAG_ONE policy] */
  free(o->Md);
  o->Md = (double *)malloc(sizeof(double) * o->nv);
  for (long r = 0; r < o->nv; ++r) {
    double dg = 0.0;
    for (long e = o->M.I[r]; e < o->M.I[r + 1]; ++e)
      if (o->M.J[e] == r) {
        if (o->ess[r]) o->M.A[e] = 1.0;
        dg = o->M.A[e];
      }
    o->Md[r] = dg;
  }
  if (q1_canonical(o)) q1_coefficients(o, S); /* canonical diag(M): the GPU's quadrature sums, bit for bit */
  /* S = B * (diag(Md)^-1 B^T): row-by-row SpGEMM, columns sorted ascending */
  {
    long cap2 = 0;
    for (long r = 0; r < o->B.nr; ++r)
      for (long e = o->B.I[r]; e < o->B.I[r + 1]; ++e) cap2 += o->BT.I[o->B.J[e] + 1] - o->BT.I[o->B.J[e]];
    long *r2 = (long *)malloc(sizeof(long) * (cap2 > 0 ? cap2 : 1));
    int *c2 = (int *)malloc(sizeof(int) * (cap2 > 0 ? cap2 : 1));
    double *v2 = (double *)malloc(sizeof(double) * (cap2 > 0 ? cap2 : 1));
    long k = 0;
    for (long r = 0; r < o->B.nr; ++r)
      for (long e = o->B.I[r]; e < o->B.I[r + 1]; ++e) {
        const long f = o->B.J[e];
        const double bf = o->B.A[e];
        if (bf == 0.0) continue; /* eliminated columns */
        for (long g = o->BT.I[f]; g < o->BT.I[f + 1]; ++g) {
          r2[k] = r; c2[k] = o->BT.J[g];
          v2[k] = bf * (o->BT.A[g] * (1.0 / o->Md[f])); /* ScaleRow(i, 1.0/Md(i)) [REF :79-80] */
          ++k;
        }
      }
    csr_free(&o->S);
    if (coo_to_csr(o->np, o->np, k, r2, c2, v2, &o->S)) return -1;
    free(r2); free(c2); free(v2);
  }
  o->have_system = 1;
  return 0;
}

/* ---------------------------------------------------------------- operators */

/* BlockMatrix::Mult: y = 0; y_u += M x_u; y_u += BT x_p; y_p += B x_u [MFEM-recall] */
static void darcy_mult(const orc *o, const double *x, double *y) {
  if (q1_canonical(o)) { q1_darcy_mult((orc *)o, x, y); return; }
  csr_mult(o, &o->M, x, y, 0);
  csr_mult(o, &o->BT, x + o->nv, y, 1);
  csr_mult(o, &o->B, x, y + o->nv, 0);
}

/* GSSmoother(S), type 0, 1 iteration, iterative_mode=false: x=0; forward sweep; backward sweep
 * SparseMatrix::Gauss_Seidel_forw/back [MFEM-recall] */
static void sgs_apply(const csr *S, const double *b, double *x) {
  const long n = S->nr;
  for (long i = 0; i < n; ++i) x[i] = 0.0;
  for (long i = 0; i < n; ++i) {
    double sum = 0.0, dg = 0.0;
    for (long e = S->I[i]; e < S->I[i + 1]; ++e) {
      if (S->J[e] == i) dg = S->A[e];
      else sum += x[S->J[e]] * S->A[e];
    }
    x[i] = (b[i] - sum) / dg;
  }
  for (long i = n - 1; i >= 0; --i) {
    double sum = 0.0, dg = 0.0;
    for (long e = S->I[i]; e < S->I[i + 1]; ++e) {
      if (S->J[e] == i) dg = S->A[e];
      else sum += x[S->J[e]] * S->A[e];
    }
    x[i] = (b[i] - sum) / dg;
  }
}

extern int orc_mg_setup(orc *o);
extern void orc_mg_apply(orc *o, const double *r, double *y);

/* BlockDiagonalPreconditioner{DSmoother(M), GSSmoother(S)}  [REF src/pressure_solver.cpp:83-91] */
static void prec_mult(orc *o, const double *r, double *y) {
  const long nv = o->nv;
#pragma omp parallel for schedule(static) if (o->threads > 1)
  for (long i = 0; i < nv; ++i) y[i] = r[i] / o->Md[i]; /* DSmoother: x = b / diag, 1 Jacobi iteration from 0 */
  if (o->d.schur_prec == IMPES_SCHUR_MG) orc_mg_apply(o, r + nv, y + nv);
  else sgs_apply(&o->S, r + nv, y + nv);
}

/* MINRESSolver::Mult with SPD preconditioner  [MFEM-recall: linalg/solvers.cpp, van der Vorst Fig 6.9]
 * [REF src/pressure_solver.cpp:93-102] */
static void minres(orc *o, const double *b, double *x, impes_solve_stats *st) {
  const long n = o->nv + o->np;
  double *v0 = (double *)calloc(n, sizeof(double)), *v1 = (double *)malloc(sizeof(double) * n);
  double *w0 = (double *)calloc(n, sizeof(double)), *w1 = (double *)calloc(n, sizeof(double));
  double *q = (double *)calloc(n, sizeof(double)), *u1 = (double *)calloc(n, sizeof(double));
  const int T = o->threads > 1;
  int it, converged = 1;
  double beta, eta, gamma0, gamma1, sigma0, sigma1, alpha, delta, rho1, rho2, rho3, norm_goal;
  memcpy(v1, b, sizeof(double) * n); /* x = 0 => v1 = b - A x = b */
  memset(x, 0, sizeof(double) * n);
  prec_mult(o, v1, u1);
  eta = beta = sqrt(vdot(o, u1, v1, n));
  gamma0 = gamma1 = 1.;
  sigma0 = sigma1 = 0.;
  norm_goal = fmax(o->d.darcy_rtol * eta, o->d.darcy_atol);
  st->initial_norm = eta;
  if (eta <= norm_goal) { it = 0; goto loop_end; }
  if (o->d.darcy_print_level == 1) printf("MINRES: iteration %3d: ||r||_B = %g\n", 0, eta);
  for (it = 1; it <= o->d.darcy_maxiter; it++) {
    {
      const double m = 1.0 / beta; /* Vector::operator/= multiplies by 1/c [MFEM-recall] */
#pragma omp parallel for schedule(static) if (T)
      for (long i = 0; i < n; ++i) { v1[i] *= m; u1[i] *= m; }
    }
    darcy_mult(o, u1, q);
    alpha = q1_canonical(o) ? o->q1_alpha : vdot(o, u1, q, n); /* canonical DG-Q1: the dot fused into K1 (two passes) */
    if (it > 1) {
#pragma omp parallel for schedule(static) if (T)
      for (long i = 0; i < n; ++i) q[i] += -beta * v0[i]; /* q.Add(-beta, v0) */
    }
#pragma omp parallel for schedule(static) if (T)
    for (long i = 0; i < n; ++i) v0[i] = q[i] + (-alpha) * v1[i]; /* add(q, -alpha, v1, v0) */
    delta = gamma1 * alpha - gamma0 * sigma1 * beta;
    rho3 = sigma0 * beta;
    rho2 = sigma1 * alpha + gamma0 * gamma1 * beta;
    prec_mult(o, v0, q);
    {
      const int fold = o->arith == 1 && (o->threads <= 1 || o->det) && o->kv == 1 && o->kp == 1 && o->d.schur_prec == IMPES_SCHUR_MG;
      beta = sqrt(fold ? canon_dot_fold(o, v0, q) : vdot(o, v0, q, n));
    }
    rho1 = o->arith == 1 ? sqrt(delta * delta + beta * beta) : hypot(delta, beta);
    if (it == 1) {
      const double a = 1. / rho1;
#pragma omp parallel for schedule(static) if (T)
      for (long i = 0; i < n; ++i) w0[i] = a * u1[i]; /* w0.Set(1./rho1, z) */
    } else if (it == 2) {
      const double a = 1. / rho1, c = -rho2 / rho1;
#pragma omp parallel for schedule(static) if (T)
      for (long i = 0; i < n; ++i) w0[i] = a * u1[i] + c * w1[i]; /* add(1./rho1, z, -rho2/rho1, w1, w0) */
    } else {
      const double a = -rho3 / rho1, c = -rho2 / rho1, e = 1. / rho1;
#pragma omp parallel for schedule(static) if (T)
      for (long i = 0; i < n; ++i) {
        const double t = a * w0[i] + c * w1[i]; /* add(-rho3/rho1, w0, -rho2/rho1, w1, w0) */
        w0[i] = t + e * u1[i];                  /* w0.Add(1./rho1, z) */
      }
    }
    gamma0 = gamma1;
    gamma1 = delta / rho1;
    {
      const double a = gamma1 * eta;
#pragma omp parallel for schedule(static) if (T)
      for (long i = 0; i < n; ++i) x[i] += a * w0[i]; /* x.Add(gamma1*eta, w0) */
    }
    sigma0 = sigma1;
    sigma1 = beta / rho1;
    eta = -sigma1 * eta;
    if (o->d.darcy_print_level == 1) printf("MINRES: iteration %3d: ||r||_B = %g\n", it, fabs(eta));
    if (fabs(eta) <= norm_goal) goto loop_end;
    { double *t = u1; u1 = q; q = t; }
    { double *t = v0; v0 = v1; v1 = t; }
    { double *t = w0; w0 = w1; w1 = t; }
  }
  converged = 0;
  it--;
loop_end:
  st->iterations = it;
  st->converged = converged;
  st->final_norm = fabs(eta);
  free(v0); free(v1); free(w0); free(w1); free(q); free(u1);
}

/* ---------------------------------------------------------------- public: pressure */

/* PressureSolver  [REF src/pressure_solver.cpp:7-125].  S: saturation dofs (ns); x: out [V|P]. */
ORC_API int orc_pressure_solve(orc *o, const double *S, double *x, impes_solve_stats *st) {
  impes_solve_stats loc;
  if (!st) st = &loc;
  memset(st, 0, sizeof(*st));
  const double t0 = now_s();
  /* the reference re-assembles B and the rhs on every call too [REF :41-54] */
  if (assemble_B_and_rhs(o)) return IMPES_ERR_BAD_ARG;
  if (assemble_system(o, S)) return IMPES_ERR_BAD_ARG;
  if (o->d.schur_prec == IMPES_SCHUR_MG && orc_mg_setup(o)) return IMPES_ERR_UNSUPPORTED;
  minres(o, o->rhs, x, st);
  st->seconds = now_s() - t0;
  return st->converged ? IMPES_OK : IMPES_ERR_NOT_CONVERGED;
}

ORC_API int orc_update_coefficients(orc *o, const double *S) {
  if (assemble_system(o, S)) return IMPES_ERR_BAD_ARG;
  if (o->d.schur_prec == IMPES_SCHUR_MG && orc_mg_setup(o)) return IMPES_ERR_UNSUPPORTED;
  return IMPES_OK;
}
ORC_API int orc_darcy_apply(orc *o, const double *x, double *y) {
  if (!o->have_system) return IMPES_ERR_BAD_ARG;
  darcy_mult(o, x, y);
  return IMPES_OK;
}
ORC_API int orc_prec_apply(orc *o, const double *r, double *y) {
  if (!o->have_system) return IMPES_ERR_BAD_ARG;
  prec_mult(o, r, y);
  return IMPES_OK;
}
ORC_API int orc_get_rhs(orc *o, double *rhs) {
  memcpy(rhs, o->rhs, sizeof(double) * (o->nv + o->np));
  return IMPES_OK;
}
ORC_API int orc_get_diag_M(orc *o, double *md) {
  if (!o->have_system) return IMPES_ERR_BAD_ARG;
  memcpy(md, o->Md, sizeof(double) * o->nv);
  return IMPES_OK;
}
/* which: 0 = M, 1 = B, 2 = S.  Pass NULL arrays to query nnz. */
ORC_API long orc_get_csr(orc *o, int which, long *I, int *J, double *A) {
  const csr *m = which == 0 ? &o->M : (which == 1 ? &o->B : &o->S);
  if (which != 1 && !o->have_system) return -1;
  if (I) memcpy(I, m->I, sizeof(long) * (m->nr + 1));
  if (J) memcpy(J, m->J, sizeof(int) * m->nnz);
  if (A) memcpy(A, m->A, sizeof(double) * m->nnz);
  return m->nnz;
}

/* ---------------------------------------------------------------- public: saturation */

/* FS  [REF src/saturation_solver.cpp:8-17] */
static void frac_flow(const orc *o, const double *S, double *fs, long n) {
#pragma omp parallel for schedule(static) if (o->threads > 1)
  for (long i = 0; i < n; ++i) {
    const double mw = krw(&o->d, S[i]) / o->d.mu_w;
    const double mo = kro(&o->d, S[i]) / o->d.mu_o;
    fs[i] = mw / (mw + mo);
  }
}

/* Kt = -(K^T) with K = Convection(v,-1) + InteriorFace Transpose(DGTrace(v,1,1/2))
 * [REF src/saturation_solver.cpp:32-35,53-55]; for DG_0 the volume term vanishes (grad w = 0)
 * and the face term is
 *   (Kt g)_i = - sum_F [ (v.n)+ g_1 + (v.n)- g_2 ] (w_i1 - w_i2),   n from element 1 to element 2,
 * with  int_F v.n = the RT0 dof of F (RT_k: sum over the face's open-node dofs with Gauss weights).
 * b_i = (R, w_i) [REF :41-44]; M_phi = (phi w, s) [REF :28-31]. */
/* Reference-element RT_kv velocity component d at xi (all dofs +d oriented; contravariant Piola factors cancel
 * against the geometry in both the volume and the face terms, see assemble_saturation_q1) */
static double vhat(const orc *o, const double *u, const int *cc, int d, const double *xi) {
  const int kv = o->kv, dim = o->dim;
  int cnt[3];
  for (int e = 0; e < 3; ++e) cnt[e] = e >= dim ? 1 : (e == d ? kv + 2 : kv + 1);
  double v = 0.0;
  for (int a2 = 0; a2 < cnt[2]; ++a2)
    for (int a1 = 0; a1 < cnt[1]; ++a1)
      for (int a0 = 0; a0 < cnt[0]; ++a0) {
        const int aa[3] = {a0, a1, a2};
        int gi[3];
        double bs = 1.0;
        for (int e = 0; e < 3; ++e) {
          gi[e] = e < dim ? cc[e] * (kv + 1) + aa[e] : 0;
          if (e < dim) bs *= (e == d) ? lag(o->cnode, kv + 2, aa[e], xi[e]) : lag(o->onode, kv + 1, aa[e], xi[e]);
        }
        v += u[vidx(o, d, gi[0], gi[1], gi[2])] * bs;
      }
  return v;
}

/* DG_1 (Q1 at the Gauss nodes) saturation operator, assembled by quadrature like MFEM's integrators:
 *   Kt = -(Convection(v,-1) + Transpose(DGTrace(v,1,1/2)))^T = Convection(v,1)^T - DGTrace(v,1,1/2)
 * [REF src/saturation_solver.cpp:32-35,53-55] ([MFEM-recall]):
 *   Convection(Q,a)_ij = a int phi_i (Q . grad phi_j): adj(J) Q = the reference-element velocity, so the element
 *     matrix is mesh independent; rule order = OrderGrad + Order + el.order = dim + 2 ks - 1;
 *   DGTrace(u,1,1/2): a = un/2, b = |un|/2 with un = u . (area-scaled normal from element 1 to 2) = the
 *     reference normal component, u evaluated from element 1; blocks  (1,1) += w(a+b) s1 s1, (2,1) -= w(a+b) s2 s1,
 *     (2,2) += w(b-a) s2 s2, (1,2) -= w(b-a) s1 s2;  rule order = OrderW + 2 ks = dim - 1 + 2 ks.
 * b = (R, w) with DomainLFIntegrator's rule of order 2 ks; M_phi = (phi w, s) is diagonal in this basis (A.6). */
static int assemble_saturation_q1(orc *o, const double *u, csr *Kt, double *b, double *Mphi) {
  const int dim = o->dim, ks = o->ks, n1 = ks + 1, nsl = o->nsl;
  qrule qv, qf, qb;
  gauss01((dim + 2 * ks - 1) / 2 + 1, &qv);
  gauss01((dim - 1 + 2 * ks) / 2 + 1, &qf);
  gauss01(ks + 1, &qb);
  qrule one = {1, {0.5}, {1.0}};
  const long cap = o->ncell * (long)nsl * nsl * (1 + 4 * dim);
  long *ri = (long *)malloc(sizeof(long) * cap);
  int *ci = (int *)malloc(sizeof(int) * cap);
  double *vv = (double *)malloc(sizeof(double) * cap);
  if (!ri || !ci || !vv) return -1;
  long nnz = 0;
  double sh[8], dsh[8][3], el[8][8];
  /* volume term: Kt_ij += sum_q w_q psi_j(q) (vhat(q) . grad psi_i(q)) */
  const qrule *QV[3] = {&qv, &qv, dim == 3 ? &qv : &one};
  for (int cz = 0; cz < o->n[2]; ++cz)
    for (int cy = 0; cy < o->n[1]; ++cy)
      for (int cx = 0; cx < o->n[0]; ++cx) {
        const int cc[3] = {cx, cy, cz};
        const long cell = cx + (long)o->n[0] * (cy + (long)o->n[1] * cz);
        memset(el, 0, sizeof el);
        for (int q2 = 0; q2 < QV[2]->nq; ++q2)
          for (int q1 = 0; q1 < QV[1]->nq; ++q1)
            for (int q0 = 0; q0 < QV[0]->nq; ++q0) {
              const double xi[3] = {QV[0]->x[q0], QV[1]->x[q1], QV[2]->x[q2]};
              const double w = QV[0]->w[q0] * QV[1]->w[q1] * QV[2]->w[q2];
              double vh[3] = {0, 0, 0};
              for (int d = 0; d < dim; ++d) vh[d] = vhat(o, u, cc, d, xi);
              for (int m = 0; m < nsl; ++m) {
                const int mm[3] = {m % n1, (m / n1) % n1, m / (n1 * n1)};
                sh[m] = 1.0;
                for (int e = 0; e < dim; ++e) sh[m] *= lag(o->snode, n1, mm[e], xi[e]);
                for (int d = 0; d < dim; ++d) {
                  double g = 1.0;
                  for (int e = 0; e < dim; ++e) g *= (e == d) ? dlag(o->snode, n1, mm[e], xi[e]) : lag(o->snode, n1, mm[e], xi[e]);
                  dsh[m][d] = g;
                }
              }
              for (int i = 0; i < nsl; ++i) {
                double vg = 0.0;
                for (int d = 0; d < dim; ++d) vg += vh[d] * dsh[i][d];
                for (int j = 0; j < nsl; ++j) el[i][j] += w * sh[j] * vg;
              }
            }
        for (int i = 0; i < nsl; ++i)
          for (int j = 0; j < nsl; ++j) { ri[nnz] = cell * nsl + i; ci[nnz] = (int)(cell * nsl + j); vv[nnz] = el[i][j]; ++nnz; }
      }
  /* interior faces: Kt -= DGTrace */
  const qrule *QF2 = dim == 3 ? &qf : &one;
  for (int d = 0; d < dim; ++d) {
    int tr[2], nt = 0;
    for (int e = 0; e < dim; ++e) if (e != d) tr[nt++] = e;
    for (int cz = 0; cz < o->n[2]; ++cz)
      for (int cy = 0; cy < o->n[1]; ++cy)
        for (int cx = 0; cx < o->n[0]; ++cx) {
          const int c2c[3] = {cx, cy, cz};
          if (c2c[d] == 0) continue; /* boundary faces omitted [REF :36-39] */
          int c1c[3] = {cx, cy, cz};
          c1c[d] -= 1;
          const long cell2 = cx + (long)o->n[0] * (cy + (long)o->n[1] * cz);
          const long stride = d == 0 ? 1 : (d == 1 ? o->n[0] : (long)o->n[0] * o->n[1]);
          const long cell1 = cell2 - stride;
          double e11[8][8], e12[8][8], e21[8][8], e22[8][8];
          memset(e11, 0, sizeof e11); memset(e12, 0, sizeof e12); memset(e21, 0, sizeof e21); memset(e22, 0, sizeof e22);
          for (int qb2 = 0; qb2 < QF2->nq; ++qb2)
            for (int qa = 0; qa < qf.nq; ++qa) {
              double xi1[3] = {0, 0, 0}, xi2[3] = {0, 0, 0};
              xi1[d] = 1.0; xi2[d] = 0.0;
              xi1[tr[0]] = xi2[tr[0]] = qf.x[qa];
              double w = qf.w[qa];
              if (nt == 2) { xi1[tr[1]] = xi2[tr[1]] = QF2->x[qb2]; w *= QF2->w[qb2]; }
              const double un = vhat(o, u, c1c, d, xi1); /* velocity from element 1, reference normal component */
              const double a = 0.5 * un, bb = 0.5 * fabs(un);
              double s1[8], s2[8];
              for (int m = 0; m < nsl; ++m) {
                const int mm[3] = {m % n1, (m / n1) % n1, m / (n1 * n1)};
                s1[m] = s2[m] = 1.0;
                for (int e = 0; e < dim; ++e) { s1[m] *= lag(o->snode, n1, mm[e], xi1[e]); s2[m] *= lag(o->snode, n1, mm[e], xi2[e]); }
              }
              const double wp = w * (a + bb), wm = w * (bb - a);
              for (int i = 0; i < nsl; ++i)
                for (int j = 0; j < nsl; ++j) {
                  if (wp != 0.0) { e11[i][j] += wp * s1[i] * s1[j]; e21[i][j] -= wp * s2[i] * s1[j]; }
                  if (wm != 0.0) { e22[i][j] += wm * s2[i] * s2[j]; e12[i][j] -= wm * s1[i] * s2[j]; }
                }
            }
          for (int i = 0; i < nsl; ++i)
            for (int j = 0; j < nsl; ++j) {
              ri[nnz] = cell1 * nsl + i; ci[nnz] = (int)(cell1 * nsl + j); vv[nnz] = -e11[i][j]; ++nnz;
              ri[nnz] = cell1 * nsl + i; ci[nnz] = (int)(cell2 * nsl + j); vv[nnz] = -e12[i][j]; ++nnz;
              ri[nnz] = cell2 * nsl + i; ci[nnz] = (int)(cell1 * nsl + j); vv[nnz] = -e21[i][j]; ++nnz;
              ri[nnz] = cell2 * nsl + i; ci[nnz] = (int)(cell2 * nsl + j); vv[nnz] = -e22[i][j]; ++nnz;
            }
        }
  }
  if (coo_to_csr(o->ns, o->ns, nnz, ri, ci, vv, Kt)) return -1;
  free(ri); free(ci); free(vv);
  /* b and the (diagonal) mass matrix */
  const qrule *QB[3] = {&qb, &qb, dim == 3 ? &qb : &one};
  for (int cz = 0; cz < o->n[2]; ++cz)
    for (int cy = 0; cy < o->n[1]; ++cy)
      for (int cx = 0; cx < o->n[0]; ++cx) {
        const int cc[3] = {cx, cy, cz};
        const long cell = cx + (long)o->n[0] * (cy + (long)o->n[1] * cz);
        for (int m = 0; m < nsl; ++m) {
          const int mm[3] = {m % n1, (m / n1) % n1, m / (n1 * n1)};
          double sb = 0.0, sm = 0.0;
          for (int q2 = 0; q2 < QB[2]->nq; ++q2)
            for (int q1 = 0; q1 < QB[1]->nq; ++q1)
              for (int q0 = 0; q0 < QB[0]->nq; ++q0) {
                const int qq[3] = {q0, q1, q2};
                double w = o->vol, psi = 1.0, p[3] = {0, 0, 0};
                for (int e = 0; e < dim; ++e) {
                  w *= QB[e]->w[qq[e]];
                  psi *= lag(o->snode, n1, mm[e], QB[e]->x[qq[e]]);
                  p[e] = (cc[e] + QB[e]->x[qq[e]]) * o->h[e];
                }
                sb += w * sat_R(&o->d, p) * psi;
                sm += w * o->phi[cell] * psi * psi; /* exact: the basis is orthogonal under the (ks+1)-point Gauss rule */
              }
          b[cell * nsl + m] = sb;
          Mphi[cell * nsl + m] = sm;
        }
      }
  return 0;
}

static int assemble_saturation(orc *o, const double *u, csr *Kt, double *b, double *Mphi) {
  if (o->ks == 1) return assemble_saturation_q1(o, u, Kt, b, Mphi);
  if (o->ks != 0) { snprintf(o->err, sizeof o->err, "oracle: order_s=%d saturation not restated yet", o->ks); return -1; }
  const int dim = o->dim, kv = o->kv;
  long nf = 0;
  for (int d = 0; d < dim; ++d) nf += o->ncell; /* upper bound on interior faces */
  long *ri = (long *)malloc(sizeof(long) * nf * 4);
  int *ci = (int *)malloc(sizeof(int) * nf * 4);
  double *vv = (double *)malloc(sizeof(double) * nf * 4);
  long nnz = 0;
  for (int d = 0; d < dim; ++d) {
    int lim[3] = {o->n[0], o->n[1], o->n[2]};
    for (int cz = 0; cz < lim[2]; ++cz)
      for (int cy = 0; cy < lim[1]; ++cy)
        for (int cx = 0; cx < lim[0]; ++cx) {
          const int cc[3] = {cx, cy, cz};
          if (cc[d] == 0) continue; /* face between cell-1 (element 1) and cell (element 2); boundary faces omitted [REF :36-39] */
          const long c2 = cx + (long)o->n[0] * (cy + (long)o->n[1] * cz);
          const long stride = d == 0 ? 1 : (d == 1 ? o->n[0] : (long)o->n[0] * o->n[1]);
          const long c1 = c2 - stride;
          /* flux through the face in +d direction = sum of the face's RT dofs x Gauss weights */
          double F = 0.0;
          const int no = kv + 1;
          const double wq[2][2] = {{1.0, 0.0}, {0.5, 0.5}};
          for (int a2 = 0; a2 < (dim == 3 ? no : 1); ++a2)
            for (int a1 = 0; a1 < no; ++a1) {
              int gi[3], t = 0;
              const int aa[2] = {a1, a2};
              double w = 1.0;
              for (int e = 0; e < 3; ++e) {
                if (e >= dim) { gi[e] = 0; continue; }
                if (e == d) gi[e] = cc[e] * (kv + 1);
                else { gi[e] = cc[e] * (kv + 1) + aa[t]; w *= wq[kv][aa[t]]; ++t; }
              }
              F += w * u[vidx(o, d, gi[0], gi[1], gi[2])];
            }
          const double Fp = F > 0.0 ? F : 0.0, Fm = F < 0.0 ? F : 0.0;
          ri[nnz] = c1; ci[nnz] = (int)c1; vv[nnz] = -Fp; ++nnz;
          ri[nnz] = c1; ci[nnz] = (int)c2; vv[nnz] = -Fm; ++nnz;
          ri[nnz] = c2; ci[nnz] = (int)c1; vv[nnz] = Fp; ++nnz;
          ri[nnz] = c2; ci[nnz] = (int)c2; vv[nnz] = Fm; ++nnz;
        }
  }
  if (coo_to_csr(o->ns, o->ns, nnz, ri, ci, vv, Kt)) return -1;
  free(ri); free(ci); free(vv);
  for (int cz = 0; cz < o->n[2]; ++cz)
    for (int cy = 0; cy < o->n[1]; ++cy)
      for (int cx = 0; cx < o->n[0]; ++cx) {
        const long c = cx + (long)o->n[0] * (cy + (long)o->n[1] * cz);
        const double p[3] = {(cx + 0.5) * o->h[0], (cy + 0.5) * o->h[1], (cz + 0.5) * o->h[2]};
        b[c] = sat_R(&o->d, p) * o->vol; /* DomainLFIntegrator, order 0: cell-centre rule [MFEM-recall] */
        Mphi[c] = o->phi[c] * o->vol;    /* MassIntegrator(Phi) on DG_0 */
      }
  return 0;
}

/* y = M_phi^-1 z.  mass_cg: Jacobi-preconditioned CG exactly as configured in the reference
 * (rtol 1e-9, atol 0, maxit 100, zero initial guess) [REF src/saturation_solver.cpp:60-68]
 * [MFEM-recall: CGSolver::Mult]; otherwise the exact diagonal inverse. */
static void mass_solve(const orc *o, const double *Mphi, const double *z, double *y, long n) {
  if (!o->mass_cg) {
#pragma omp parallel for schedule(static) if (o->threads > 1)
    for (long i = 0; i < n; ++i) y[i] = z[i] / Mphi[i];
    return;
  }
  double *r = (double *)malloc(sizeof(double) * n), *dd = (double *)malloc(sizeof(double) * n),
         *zz = (double *)malloc(sizeof(double) * n);
  for (long i = 0; i < n; ++i) { y[i] = 0.0; r[i] = z[i]; zz[i] = r[i] / Mphi[i]; dd[i] = zz[i]; }
  double nom0 = vdot(o, dd, r, n), nom = nom0;
  const double r0 = fmax(nom * 1e-9 * 1e-9, 0.0);
  if (nom > r0) {
    for (int it = 1; it <= 100; ++it) {
      for (long i = 0; i < n; ++i) zz[i] = Mphi[i] * dd[i];
      const double den = vdot(o, zz, dd, n);
      const double alpha = nom / den;
      for (long i = 0; i < n; ++i) { y[i] += alpha * dd[i]; r[i] -= alpha * zz[i]; }
      for (long i = 0; i < n; ++i) zz[i] = r[i] / Mphi[i];
      const double betanom = vdot(o, r, zz, n);
      if (betanom < r0 || !(betanom == betanom)) break;
      const double beta = betanom / nom;
      for (long i = 0; i < n; ++i) dd[i] = zz[i] + beta * dd[i];
      nom = betanom;
    }
  }
  free(r); free(dd); free(zz);
}

static int n_substeps(const orc *o, double global_dt, double dt) {
  if (o->d.substep_rule == IMPES_SUBSTEPS_PARALLEL) {
    int n = 0;
    double t = 0.0;
    while (!(t >= global_dt - dt / 2)) { t += dt; ++n; } /* [REF src/par_saturation_solver.cpp:149-155] */
    return n;
  }
  return (int)(global_dt / dt); /* int nt = global_dt / dt  [REF src/saturation_solver.cpp:74-75] */
}

/* F(S) = M_phi^-1 (Kt f(S) + b)   FE_Evolution::Mult [REF src/par_saturation_solver.cpp:46-62] */
static const double *g_q1_u; /* velocity of the running orc_saturation_advance call (canonical DG-Q1 stages) */
static void sat_rate(const orc *o, const csr *Kt, const double *b, const double *Mphi, const double *S, double *fs,
                     double *z, double *y) {
  const long n = o->ns;
  if (q1_canonical(o)) { /* the GPU's fused stage, operation by operation (orc_q1.inc) */
    if (o->d.two_phase) frac_flow(o, S, fs, n);
    else memcpy(fs, S, sizeof(double) * n);
    q1_saturation_rate(o, g_q1_u, S, fs, y, 0.0, 1);
    return;
  }
  if (o->d.two_phase) { frac_flow(o, S, fs, n); csr_mult(o, Kt, fs, z, 0); }
  else csr_mult(o, Kt, S, z, 0);
#pragma omp parallel for schedule(static) if (o->threads > 1)
  for (long i = 0; i < n; ++i) z[i] += b[i];
  mass_solve(o, Mphi, z, y, n);
}

/* SaturationSolver [REF src/saturation_solver.cpp:19-112]; -ode variants of the MPI path
 * [REF src/par_saturation_solver.cpp:113-124] ([MFEM-recall]). */
ORC_API int orc_saturation_advance(orc *o, const double *u, double *S, double global_dt, double dt, int *nsub_out) {
  const long n = o->ns;
  g_q1_u = u;
  csr Kt; memset(&Kt, 0, sizeof Kt);
  double *b = (double *)malloc(sizeof(double) * n), *Mphi = (double *)malloc(sizeof(double) * n);
  if (assemble_saturation(o, u, &Kt, b, Mphi)) { free(b); free(Mphi); return IMPES_ERR_UNSUPPORTED; }
  double *y = (double *)malloc(sizeof(double) * n), *z = (double *)malloc(sizeof(double) * n),
         *fs = (double *)malloc(sizeof(double) * n);
  double *k2 = NULL, *k3 = NULL, *k4 = NULL, *tmp = NULL;
  const int ode = (o->d.substep_rule == IMPES_SUBSTEPS_SERIAL) ? 1 : o->d.ode_solver_type; /* serial ignores -ode */
  int rc = IMPES_OK;
  if (ode != 1) {
    k2 = (double *)malloc(sizeof(double) * n); k3 = (double *)malloc(sizeof(double) * n);
    k4 = (double *)malloc(sizeof(double) * n); tmp = (double *)malloc(sizeof(double) * n);
  }
  const int nt = n_substeps(o, global_dt, dt);
  for (int ti = 1; ti <= nt && rc == IMPES_OK; ++ti) {
    switch (ode) {
      case 1: /* forward Euler: y = F(S); y *= dt; S += y */
        sat_rate(o, &Kt, b, Mphi, S, fs, z, y);
        for (long i = 0; i < n; ++i) { y[i] *= dt; S[i] += y[i]; }
        break;
      case 2: { /* RK2Solver(a=1), b = 0.5/a: k=F(x); x1 = x+(1-b)dt k; x += a dt k; k=F(x); x = x1 + b dt k */
        const double a = 1.0, bb = 0.5 / a;
        sat_rate(o, &Kt, b, Mphi, S, fs, z, y);
        for (long i = 0; i < n; ++i) { tmp[i] = S[i] + ((1. - bb) * dt) * y[i]; S[i] += (a * dt) * y[i]; }
        sat_rate(o, &Kt, b, Mphi, S, fs, z, y);
        for (long i = 0; i < n; ++i) S[i] = tmp[i] + (bb * dt) * y[i];
        break;
      }
      case 3: /* RK3SSPSolver */
        sat_rate(o, &Kt, b, Mphi, S, fs, z, k2);
        for (long i = 0; i < n; ++i) y[i] = S[i] + dt * k2[i];
        sat_rate(o, &Kt, b, Mphi, y, fs, z, k2);
        for (long i = 0; i < n; ++i) { y[i] += dt * k2[i]; y[i] = 0.75 * S[i] + 0.25 * y[i]; }
        sat_rate(o, &Kt, b, Mphi, y, fs, z, k2);
        for (long i = 0; i < n; ++i) { y[i] += dt * k2[i]; S[i] = (1. / 3) * S[i] + (2. / 3) * y[i]; }
        break;
      case 4: /* RK4Solver: z accumulates x + dt/6 k1 + dt/3 k2 + dt/3 k3, then x = z + dt/6 k4 */
        sat_rate(o, &Kt, b, Mphi, S, fs, z, y);
        for (long i = 0; i < n; ++i) { tmp[i] = S[i] + (dt / 2) * y[i]; k4[i] = S[i] + (dt / 6) * y[i]; }
        sat_rate(o, &Kt, b, Mphi, tmp, fs, z, y);
        for (long i = 0; i < n; ++i) { tmp[i] = S[i] + (dt / 2) * y[i]; k4[i] += (dt / 3) * y[i]; }
        sat_rate(o, &Kt, b, Mphi, tmp, fs, z, y);
        for (long i = 0; i < n; ++i) { tmp[i] = S[i] + dt * y[i]; k4[i] += (dt / 3) * y[i]; }
        sat_rate(o, &Kt, b, Mphi, tmp, fs, z, y);
        for (long i = 0; i < n; ++i) S[i] = k4[i] + (dt / 6) * y[i];
        break;
      default:
        snprintf(o->err, sizeof o->err, "Uknown ODE solver: type %d", ode); /* [REF src/par_saturation_solver.cpp:121-123]; RK6 tableau unpinned */
        rc = IMPES_ERR_UNSUPPORTED;
    }
  }
  if (nsub_out) *nsub_out = nt;
  csr_free(&Kt);
  free(b); free(Mphi); free(y); free(z); free(fs); free(k2); free(k3); free(k4); free(tmp);
  return rc;
}

/* guards of the global loop  [REF src/run_serial.cpp:132-142] */
ORC_API void orc_guards(orc *o, const double *S, long *n_nonfinite, double *s_max) {
  long bad = 0;
  double mx = S[0];
  for (long i = 0; i < o->ns; ++i) {
    if (!isfinite(S[i])) ++bad;
    if (S[i] > mx) mx = S[i];
  }
  *n_nonfinite = bad;
  *s_max = mx;
}

#include "orc_mg.inc"
#include "orc_q1.inc"
