// api.cu -- C ABI (include/impes_b200.h) over the sm_100a kernels in kernels.cuh.
//
// Replaces, behind plain-C entry points, the bodies of
//   PressureSolver   [REF src/pressure_solver.cpp:7-125]
//   SaturationSolver [REF src/saturation_solver.cpp:19-112]
// (and their MPI analogues [REF src/par_pressure_solver.cpp] [REF src/par_saturation_solver.cpp]).
// Nothing is assembled: the state a solve needs (coefficients, Krylov vectors, halo buffers,
// the NCCL communicator) lives in the handle.  No CPU fallback exists in this file.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/impes_b200.h"
#include "kernels.cuh"
#include "mg.cuh"
#include "rt1.cuh"
#include "rt1q.cuh"
#include "nccl_shim.h"

using namespace impes;

#define API extern "C" __attribute__((visibility("default")))

namespace {

std::string g_create_error;

enum { KC_DARCY = 0, KC_MINRES = 1, KC_PREC = 2, KC_SAT = 3, KC_COEFF = 4, KC_DOT = 5, KC_MISC = 6, KC_COMM = 7, KC_N = 8 };

// kernel-level accounting (impes_profile_read_kernels): one id per kernel (or family of tiny kernels)
enum {
  KI_DARCY_RT0 = 0, KI_DARCY_RT1, KI_DARCY_Q1A, KI_DARCY_Q1B, KI_MINRES_B, KI_MINRES_C, KI_MINRES_INIT, KI_DOT,
  KI_JACOBI_U, KI_SGS, KI_MGQ_FIRST, KI_MGQ_JACOBI, KI_MGQ_RESTRICT, KI_MGQ_PROLONG, KI_MG_CELL, KI_SAT_P0,
  KI_SAT_Q1, KI_SAT_AUX, KI_COEFF, KI_MISC, KI_COMM, KI_N
};
const char* const KI_NAME[KI_N] = {
  "k_darcy_apply", "k1_darcy_apply", "k1q_apply_a", "k1q_apply_b", "k_minres_b", "k_minres_c", "k_minres_init",
  "k_dot_owned", "k_jacobi_u", "k_sgs", "k1_mg_jacobi(first)", "k1_mg_jacobi", "k1_mg_restrict", "k1_mg_prolong",
  "k_mg_cell_levels", "k_sat_stream", "k1q_sat", "sat_aux", "coeff", "misc", "nccl"};
const int KI_CLASS[KI_N] = {KC_DARCY, KC_DARCY, KC_DARCY, KC_DARCY, KC_MINRES, KC_MINRES, KC_MINRES, KC_DOT,
                            KC_PREC, KC_PREC, KC_PREC, KC_PREC, KC_PREC, KC_PREC, KC_PREC, KC_SAT,
                            KC_SAT, KC_SAT, KC_COEFF, KC_MISC, KC_COMM};

}  // namespace

struct impes_ctx {
  impes_desc d;
  Lay L;
  int ord = 0;           // 0: RT0/P0, 1: RT1/L2-Q1 (rt1.cuh); saturation DG_0 in both
  Lay1 L1;               // dvec1 layout (ord == 1); L then only describes the cell arrays and the flux vector
  Rt1Const rc;
  int grid_dofs1 = 0;    // blocks of 256 L2 dofs (ord == 1: K1, dot, Q1 multigrid level run one thread per dof)
  int vec_n = 0, vec_offp = 0;  // length and p-offset of a Darcy-sized device vector in the active layout
  double* flux0 = nullptr;      // ord == 1: face fluxes of x in the RT0 dvec layout (input of K6)
  double* aq = nullptr;         // ord == 0, os == 1: RT0 element-block entries (k0q_coeff)
  double* vhq = nullptr;        // os == 1: per-step velocity records of the DG-Q1 stage kernel (k1q_sat_pre)
  double* fnq = nullptr;
  double* mg1_ya = nullptr;     // ord == 1: second iterate buffer of the Q1 multigrid level (p-part)
  int os = 0, nsl = 1;          // saturation order (DG_os) and dofs per cell (rt1q.cuh)
  QTab qt;
  double* rlam = nullptr;       // os == 1: w_q / lambda(S(x_q)) at the mass-rule points, ncv x NQP
  double* yup = nullptr;        // os == 1: upper-face-node element contributions of K1 pass A, ncv x NUP
  double* accA = nullptr;       // os == 1: per-cell share of <x, Ax> from K1 pass A (reduced by pass B)
  int nqp = 0, nup = 0;
  Phys ph;
  SatParam sp;
  int dev = 0;
  cudaStream_t stream = nullptr;
  // global (all-rank) sizes in native host layout
  long g_ncell = 0, g_nvx = 0, g_nvy = 0, g_nvz = 0, g_nv = 0, g_np = 0;
  double h[3] = {1, 1, 1}, vol = 1, geo[3] = {0, 0, 0};  // internal directions
  // device arrays
  double* K[3] = {nullptr, nullptr, nullptr};
  double* phi = nullptr;
  double* S[2] = {nullptr, nullptr};
  double* F[2] = {nullptr, nullptr};  // f(S) carried beside S (streaming K6)
  int scur = 0;
  double* stage[4] = {nullptr, nullptr, nullptr, nullptr};  // RK stage buffers (lazy)
  double* m[3] = {nullptr, nullptr, nullptr};
  double *Md = nullptr, *T = nullptr, *rhs = nullptr;
  double *x = nullptr, *v0 = nullptr, *v1 = nullptr, *w0 = nullptr, *w1 = nullptr, *q = nullptr, *u1 = nullptr;
  double* tmpv = nullptr;  // scratch dvec for operator-level entry points
  double* partial = nullptr;
  double* partial_u = nullptr;  // per-chunk velocity part of <v0, P v0> (k_minres_b_dot), n_u_chunks entries
  int n_u_chunks = 0;       // chunks that hold velocity entries
  int n_u_chunks_all = 0;   // chunks of the whole vector (grid of k_minres_b_dot)
  bool lazy_u = false;          // DG-Q1 path: u = P v stays unscaled in memory (MinresState::mu)
  bool fold_dot = false;        // RT1/Q1 + multigrid: beta^2 is folded into k_minres_b and the last V-cycle sweep
  double* scal = nullptr;  // [0] alpha, [1] beta^2
  unsigned* ticket = nullptr;
  MinresState* st = nullptr;  // [2]
  unsigned long long* guard = nullptr;
  double* flushbuf = nullptr;
  long flush_n = 0;
  bool coeff_valid = false;
  // launch geometry
  int grid_cells = 0, grid_vec = 0, grid_sgs = 0, num_sms = 0;
  int sgs_mode = 2, sgs_threads = 256;  // 0 one CTA, 1 cluster, 2 cooperative grid
  bool sgs_skew = false;                // stage-major (skewed) sweep available for this shape
  int sgs_ppt = 1;
  bool sgs_tma = false;                 // one-CTA TMA ring version (<= 1024 pencils)
  int sgs_ns = 0;
  size_t sgs_smem = 0;
  SgsPack pack = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
  int sgs_yp = 0;
  double* pack_blk = nullptr;
  // graphs
  bool use_graph = true;
  cudaGraphExec_t minres_exec = nullptr;
  int graph_iters = 0;
  // multigrid Schur preconditioner (IMPES_SCHUR_MG)
  std::vector<MgLevel> mg;
  std::vector<double*> mg_bufs;
  int mg_g = 1 << 30;            // first replicated level (multi-GPU): levels < mg_g are slab-distributed
  std::vector<int> mg_goff, mg_gcnt;  // plane offset / count of every rank's part of level mg_g
  double mg_omega = 0.8, mg_cfac = 2.0;  // damping of the coarsest level's sweeps; coarse-correction factor between cell levels
  // Damping per smoothing sweep of a side: a 2-step (Chebyshev-like) Richardson iteration on D^-1 S with roots 1/1.4 and
  // 1/0.55; pre-smoothing runs the sequence forwards, post-smoothing backwards (the V-cycle stays symmetric).  Against
  // 0.8 throughout: 43 -> 35 MINRES iterations on the C5 family.  mg_cfacq: over-correction of the cell-wise constant
  // prolongation onto the L2-Q1 level.  (The CPU checker restates the same constants.)
  double mg_om[2] = {1.4, 0.55}, mg_cfacq = 2.5;
  double om_pre(int s) const { return mg_om[s & 1]; }
  double om_post(int s, int sweeps) const { return mg_om[(sweeps - 1 - s) & 1]; }
  int mg_nu = 2, mg_coarse_sweeps = 16;
  long long graph_launches = 0;
  long long graph_prof_n[KC_N] = {0};
  long long graph_ki_n[KI_N] = {0};
  // multi-GPU
  NcclShim nccl;
  ncclComm_t comm = nullptr;
  // accounting
  long long launches = 0;
  bool profiling = false;
  double prof_ms[KC_N] = {0};
  long long prof_n[KC_N] = {0};
  double ki_ms[KI_N] = {0};
  long long ki_n[KI_N] = {0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t user_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::string err;
};

namespace {

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512];                                                                                \
      snprintf(b_, sizeof b_, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      h->err = b_;                                                                                 \
      return IMPES_ERR_CUDA;                                                                       \
    }                                                                                              \
  } while (0)

#define NK(call)                                                                  \
  do {                                                                            \
    int r_ = (call);                                                              \
    if (r_ != 0) {                                                                \
      char b_[512];                                                               \
      snprintf(b_, sizeof b_, "%s:%d: %s -> nccl error %d", __FILE__, __LINE__, #call, r_); \
      h->err = b_;                                                                \
      return IMPES_ERR_NCCL;                                                      \
    }                                                                             \
  } while (0)

// ncclGroupStart ... ncclGroupEnd that cannot be left open by an early error return (a dangling group would make the
// next collective hang): the guard closes the group when it goes out of scope.
struct NcclGroup {
  NcclShim& n;
  bool open;
  explicit NcclGroup(NcclShim& n_) : n(n_), open(n_.GroupStart() == 0) {}
  int end() { open = false; return n.GroupEnd(); }
  ~NcclGroup() { if (open) n.GroupEnd(); }
};

#define RET_IF(x)            \
  do {                       \
    int rc_ = (x);           \
    if (rc_) return rc_;     \
  } while (0)

long pad32(long n) { return (n + 31) / 32 * 32; }

// Launch accounting (+ optional per-kernel device timing when profiling is on: every scope is then bracketed by
// an event pair and synchronised, so the profiled step is slower than a timed one but its per-kernel durations
// are those of back-to-back launches on warm data).
struct Scope {
  impes_ctx* h;
  int ki;
  int n;
  Scope(impes_ctx* h_, int ki_, int n_ = 1) : h(h_), ki(ki_), n(n_) {
    h->launches += n;
    h->prof_n[KI_CLASS[ki]] += n;
    h->ki_n[ki] += n;
    if (h->profiling) cudaEventRecord(h->ev0, h->stream);
  }
  ~Scope() {
    if (h->profiling) {
      cudaEventRecord(h->ev1, h->stream);
      cudaEventSynchronize(h->ev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, h->ev0, h->ev1);
      h->prof_ms[KI_CLASS[ki]] += ms;
      h->ki_ms[ki] += ms;
    }
  }
};

int dev_alloc(impes_ctx* h, double** p, long n) {
  CK(cudaMalloc((void**)p, sizeof(double) * (size_t)std::max<long>(n, 1)));
  CK(cudaMemsetAsync(*p, 0, sizeof(double) * (size_t)std::max<long>(n, 1), h->stream));
  return 0;
}

// ---- host (global, native) <-> device (slab local) transfers -------------------------------

// cvec (cell array with ghost planes) from a global cell array
// mult: dofs per cell (1 for cell arrays, nsl for a DG-Q1 saturation vector)
int upload_cells(impes_ctx* h, const double* host, double* dev, int mult = 1) {
  const Lay& L = h->L;
  const int k0 = std::max(L.zlo - 1, 0), k1 = std::min(L.zhi + 1, L.NZ);
  const long pl = (long)L.plane * mult;
  CK(cudaMemcpyAsync(dev + pl * (k0 - L.zlo + 1), host + pl * k0, sizeof(double) * pl * (k1 - k0),
                     cudaMemcpyHostToDevice, h->stream));
  return 0;
}
int download_cells_owned(impes_ctx* h, const double* dev, double* host, int mult = 1) {
  const Lay& L = h->L;
  const long pl = (long)L.plane * mult;
  CK(cudaMemcpyAsync(host + pl * L.zlo, dev + pl, sizeof(double) * pl * L.nzl, cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

// dvec from / to a global native Darcy vector [ux|uy|uz|p]; which: bit0 = velocity, bit1 = pressure.
// Uploads fill the ghost planes from the (global) host array as well.
int upload_dvec1(impes_ctx* h, const double* host_v, const double* host_p, double* dev) {
  const Lay1& L = h->L1;
  if (host_v) {
    const long px = (long)L.gx0 * L.gx1, py = (long)L.gy0 * L.gy1, pz = (long)L.gz0 * L.gz1;
    CK(cudaMemcpyAsync(dev + L.off_x, host_v + px * 2 * L.zlo, sizeof(double) * px * 2 * L.nzl, cudaMemcpyHostToDevice, h->stream));
    if (L.has_y)
      CK(cudaMemcpyAsync(dev + L.off_y, host_v + h->g_nvx + py * 2 * L.zlo, sizeof(double) * py * 2 * L.nzl,
                         cudaMemcpyHostToDevice, h->stream));
    const int g0 = std::max(2 * L.zlo - 2, 0), g1 = 2 * L.zhi + 1;  // node planes [g0, g1)
    CK(cudaMemcpyAsync(dev + L.off_z + pz * (g0 - 2 * L.zlo + 2), host_v + h->g_nvx + h->g_nvy + pz * g0,
                       sizeof(double) * pz * (g1 - g0), cudaMemcpyHostToDevice, h->stream));
  }
  if (host_p) {
    const int k0 = std::max(L.zlo - 1, 0), k1 = std::min(L.zhi + 1, L.NZ);
    const long pp = (long)L.plane * L.npl;
    CK(cudaMemcpyAsync(dev + L.off_p + pp * (k0 - L.zlo + 1), host_p + pp * k0, sizeof(double) * pp * (k1 - k0),
                       cudaMemcpyHostToDevice, h->stream));
  }
  return 0;
}
int download_dvec1_owned(impes_ctx* h, const double* dev, double* host_v, double* host_p) {
  const Lay1& L = h->L1;
  if (host_v) {
    const long px = (long)L.gx0 * L.gx1, py = (long)L.gy0 * L.gy1, pz = (long)L.gz0 * L.gz1;
    CK(cudaMemcpyAsync(host_v + px * 2 * L.zlo, dev + L.off_x, sizeof(double) * px * 2 * L.nzl, cudaMemcpyDeviceToHost, h->stream));
    if (L.has_y)
      CK(cudaMemcpyAsync(host_v + h->g_nvx + py * 2 * L.zlo, dev + L.off_y, sizeof(double) * py * 2 * L.nzl,
                         cudaMemcpyDeviceToHost, h->stream));
    const int nf = 2 * L.nzl + (L.zhi == L.NZ ? 1 : 0);  // owned node planes (+ top boundary on the last rank)
    CK(cudaMemcpyAsync(host_v + h->g_nvx + h->g_nvy + pz * 2 * L.zlo, dev + L.off_z + pz * 2, sizeof(double) * pz * nf,
                       cudaMemcpyDeviceToHost, h->stream));
  }
  if (host_p) {
    const long pp = (long)L.plane * L.npl;
    CK(cudaMemcpyAsync(host_p + pp * L.zlo, dev + L.off_p + pp, sizeof(double) * pp * L.nzl, cudaMemcpyDeviceToHost, h->stream));
  }
  return 0;
}

int upload_dvec(impes_ctx* h, const double* host_v, const double* host_p, double* dev) {
  if (h->ord == 1) return upload_dvec1(h, host_v, host_p, dev);
  const Lay& L = h->L;
  if (host_v) {
    const long px = (long)(L.NX + 1) * L.NY, py = (long)L.NX * (L.NY + 1);
    CK(cudaMemcpyAsync(dev + L.off_x, host_v + px * L.zlo, sizeof(double) * px * L.nzl, cudaMemcpyHostToDevice, h->stream));
    if (L.has_y)
      CK(cudaMemcpyAsync(dev + L.off_y, host_v + h->g_nvx + py * L.zlo, sizeof(double) * py * L.nzl,
                         cudaMemcpyHostToDevice, h->stream));
    const int f0 = std::max(L.zlo - 1, 0), f1 = L.zhi + 1;  // face planes [f0, f1)
    CK(cudaMemcpyAsync(dev + L.off_z + L.plane * (f0 - L.zlo + 1), host_v + h->g_nvx + h->g_nvy + L.plane * f0,
                       sizeof(double) * L.plane * (f1 - f0), cudaMemcpyHostToDevice, h->stream));
  }
  if (host_p) {
    const int k0 = std::max(L.zlo - 1, 0), k1 = std::min(L.zhi + 1, L.NZ);
    CK(cudaMemcpyAsync(dev + L.off_p + L.plane * (k0 - L.zlo + 1), host_p + L.plane * k0,
                       sizeof(double) * L.plane * (k1 - k0), cudaMemcpyHostToDevice, h->stream));
  }
  return 0;
}
int download_dvec_owned(impes_ctx* h, const double* dev, double* host_v, double* host_p) {
  if (h->ord == 1) return download_dvec1_owned(h, dev, host_v, host_p);
  const Lay& L = h->L;
  if (host_v) {
    const long px = (long)(L.NX + 1) * L.NY, py = (long)L.NX * (L.NY + 1);
    CK(cudaMemcpyAsync(host_v + px * L.zlo, dev + L.off_x, sizeof(double) * px * L.nzl, cudaMemcpyDeviceToHost, h->stream));
    if (L.has_y)
      CK(cudaMemcpyAsync(host_v + h->g_nvx + py * L.zlo, dev + L.off_y, sizeof(double) * py * L.nzl,
                         cudaMemcpyDeviceToHost, h->stream));
    const int nf = L.nzl + (L.zhi == L.NZ ? 1 : 0);  // owned face planes (+ top boundary on the last rank)
    CK(cudaMemcpyAsync(host_v + h->g_nvx + h->g_nvy + L.plane * L.zlo, dev + L.off_z + L.plane, sizeof(double) * L.plane * nf,
                       cudaMemcpyDeviceToHost, h->stream));
  }
  if (host_p)
    CK(cudaMemcpyAsync(host_p + L.plane * L.zlo, dev + L.off_p + L.plane, sizeof(double) * L.plane * L.nzl,
                       cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

// ---- slab-local host arrays (impes_*_local): the native layout restricted to what this rank owns, like the
// local dofs of a ParGridFunction [REF src/run_parallel.cpp:97-104].  Components keep the native order
// [ux | uy | uz | p]; ghosts are never read from / written to the host (the device halo exchange fills them).
struct LocalMap {
  long nx, ny, nz, np, ns;      // owned entries per component
  long gx, gy, gz, gp, gs;      // first global index of the owned part inside its global component array
  long dx, dy, dz, dp, ds;      // device offset of the first owned entry (dvec / cvec)
};
LocalMap local_map(const impes_ctx* h) {
  LocalMap m;
  const Lay& L = h->L;
  const bool top = L.zhi == L.NZ;
  if (h->ord == 1) {
    const Lay1& Q = h->L1;
    const long px = (long)Q.gx0 * Q.gx1, py = (long)Q.gy0 * Q.gy1, pz = (long)Q.gz0 * Q.gz1, pp = (long)Q.plane * Q.npl;
    m.nx = px * 2 * Q.nzl; m.gx = px * 2 * Q.zlo; m.dx = Q.off_x;
    m.ny = Q.has_y ? py * 2 * Q.nzl : 0; m.gy = Q.has_y ? py * 2 * Q.zlo : 0; m.dy = Q.off_y;
    m.nz = pz * (2 * Q.nzl + (top ? 1 : 0)); m.gz = pz * 2 * Q.zlo; m.dz = Q.off_z + pz * 2;
    m.np = pp * Q.nzl; m.gp = pp * Q.zlo; m.dp = Q.off_p + pp;
  } else {
    const long px = (long)(L.NX + 1) * L.NY, py = (long)L.NX * (L.NY + 1);
    m.nx = px * L.nzl; m.gx = px * L.zlo; m.dx = L.off_x;
    m.ny = L.has_y ? py * L.nzl : 0; m.gy = L.has_y ? py * L.zlo : 0; m.dy = L.off_y;
    m.nz = (long)L.plane * (L.nzl + (top ? 1 : 0)); m.gz = (long)L.plane * L.zlo; m.dz = L.off_z + L.plane;
    m.np = (long)L.plane * L.nzl; m.gp = (long)L.plane * L.zlo; m.dp = L.off_p + L.plane;
  }
  m.ns = (long)L.plane * h->nsl * L.nzl; m.gs = (long)L.plane * h->nsl * L.zlo; m.ds = (long)L.plane * h->nsl;
  return m;
}
// dir: cudaMemcpyHostToDevice / cudaMemcpyDeviceToHost; host_v / host_p: LOCAL arrays (either may be NULL)
int copy_dvec_local(impes_ctx* h, double* host_v, double* host_p, double* dev, cudaMemcpyKind dir) {
  const LocalMap m = local_map(h);
  auto cp = [&](double* hostp, double* devp, long n) -> cudaError_t {
    if (n <= 0) return cudaSuccess;
    return dir == cudaMemcpyHostToDevice ? cudaMemcpyAsync(devp, hostp, sizeof(double) * n, dir, h->stream)
                                         : cudaMemcpyAsync(hostp, devp, sizeof(double) * n, dir, h->stream);
  };
  if (host_v) {
    CK(cp(host_v, dev + m.dx, m.nx));
    CK(cp(host_v + m.nx, dev + m.dy, m.ny));
    CK(cp(host_v + m.nx + m.ny, dev + m.dz, m.nz));
  }
  if (host_p) CK(cp(host_p, dev + m.dp, m.np));
  return 0;
}
int copy_cells_local(impes_ctx* h, double* host, double* dev, cudaMemcpyKind dir) {
  const LocalMap m = local_map(h);
  if (dir == cudaMemcpyHostToDevice) CK(cudaMemcpyAsync(dev + m.ds, host, sizeof(double) * m.ns, dir, h->stream));
  else CK(cudaMemcpyAsync(host, dev + m.ds, sizeof(double) * m.ns, dir, h->stream));
  return 0;
}

// ---- halo exchange (world_size > 1): grouped ncclSend/Recv of whole planes ---------------------
// up = towards rank+1.  Each (ptr, plane index) pair names a plane inside a (nzl+2)-plane region.
int exchange_planes(impes_ctx* h, double* base, bool send_up_top_owned, bool send_down_bottom_owned, int mult = 1) {
  // send_up_top_owned:   my plane index nzl   (top owned)    -> upper rank's plane 0      (ghost below)
  // send_down_bottom_owned: my plane index 1 (bottom owned)  -> lower rank's plane nzl+1  (ghost above)
  if (h->L.nranks == 1) return 0;
  const Lay& L = h->L;
  const int up = L.rank + 1 < L.nranks ? L.rank + 1 : -1, dn = L.rank > 0 ? L.rank - 1 : -1;
  const size_t pl = (size_t)L.plane * mult;
  Scope sc(h, KI_COMM, 1);
  NcclGroup grp(h->nccl);
  if (!grp.open) { h->err = "ncclGroupStart failed"; return IMPES_ERR_NCCL; }
  if (send_up_top_owned) {
    if (up >= 0) NK(h->nccl.Send(base + pl * L.nzl, pl, ncclFloat64, up, h->comm, h->stream));
    if (dn >= 0) NK(h->nccl.Recv(base, pl, ncclFloat64, dn, h->comm, h->stream));
  }
  if (send_down_bottom_owned) {
    if (dn >= 0) NK(h->nccl.Send(base + pl, pl, ncclFloat64, dn, h->comm, h->stream));
    if (up >= 0) NK(h->nccl.Recv(base + pl * (L.nzl + 1), pl, ncclFloat64, up, h->comm, h->stream));
  }
  NK(grp.end());
  return 0;
}
// both ghost planes of a dof-major cell array a[m ncv + c] (f(S) of the DG-Q1 stage kernel): nsl planes per side
int exchange_planes_dofmajor(impes_ctx* h, double* base, int nsl) {
  if (h->L.nranks == 1) return 0;
  const Lay& L = h->L;
  const int up = L.rank + 1 < L.nranks ? L.rank + 1 : -1, dn = L.rank > 0 ? L.rank - 1 : -1;
  const size_t pl = (size_t)L.plane;
  Scope sc(h, KI_COMM, 1);
  NcclGroup grp(h->nccl);
  if (!grp.open) { h->err = "ncclGroupStart failed"; return IMPES_ERR_NCCL; }
  for (int m = 0; m < nsl; ++m) {
    double* b = base + (size_t)m * L.ncv;
    if (up >= 0) {
      NK(h->nccl.Send(b + pl * L.nzl, pl, ncclFloat64, up, h->comm, h->stream));
      NK(h->nccl.Recv(b + pl * (L.nzl + 1), pl, ncclFloat64, up, h->comm, h->stream));
    }
    if (dn >= 0) {
      NK(h->nccl.Recv(b, pl, ncclFloat64, dn, h->comm, h->stream));
      NK(h->nccl.Send(b + pl, pl, ncclFloat64, dn, h->comm, h->stream));
    }
  }
  NK(grp.end());
  return 0;
}
// halo of a Darcy vector before K1: p ghost-below, uz ghost-below (from the lower rank's top owned
// face plane) and uz ghost-above (the upper rank's bottom owned face plane).
int halo_darcy(impes_ctx* h, double* v) {
  if (h->L.nranks == 1) return 0;
  const Lay& L = h->L;
  const int up = L.rank + 1 < L.nranks ? L.rank + 1 : -1, dn = L.rank > 0 ? L.rank - 1 : -1;
  Scope sc(h, KI_COMM, 1);
  NcclGroup grp(h->nccl);
  if (!grp.open) { h->err = "ncclGroupStart failed"; return IMPES_ERR_NCCL; }
  if (up >= 0) {
    NK(h->nccl.Send(v + L.off_p + L.plane * L.nzl, (size_t)L.plane, ncclFloat64, up, h->comm, h->stream));
    NK(h->nccl.Send(v + L.off_z + L.plane * L.nzl, (size_t)L.plane, ncclFloat64, up, h->comm, h->stream));
    NK(h->nccl.Recv(v + L.off_z + L.plane * (L.nzl + 1), (size_t)L.plane, ncclFloat64, up, h->comm, h->stream));
  }
  if (dn >= 0) {
    NK(h->nccl.Recv(v + L.off_p, (size_t)L.plane, ncclFloat64, dn, h->comm, h->stream));
    NK(h->nccl.Recv(v + L.off_z, (size_t)L.plane, ncclFloat64, dn, h->comm, h->stream));
    NK(h->nccl.Send(v + L.off_z + L.plane, (size_t)L.plane, ncclFloat64, dn, h->comm, h->stream));
  }
  NK(grp.end());
  return 0;
}
// dvec1 halo before K1 (rt1.cuh): p ghost-below cell plane; uz ghost node planes below (2: the lower cell's
// interior and lower-face nodes) and above (1: the shared face nodes owned by the upper rank).
int halo_darcy1(impes_ctx* h, double* v) {
  if (h->L1.nranks == 1) return 0;
  const Lay1& L = h->L1;
  const int up = L.rank + 1 < L.nranks ? L.rank + 1 : -1, dn = L.rank > 0 ? L.rank - 1 : -1;
  const size_t pz = (size_t)L.gz0 * L.gz1, pp = (size_t)L.plane * L.npl;
  Scope sc(h, KI_COMM, 1);
  NcclGroup grp(h->nccl);
  if (!grp.open) { h->err = "ncclGroupStart failed"; return IMPES_ERR_NCCL; }
  if (up >= 0) {
    NK(h->nccl.Send(v + L.off_p + pp * L.nzl, pp, ncclFloat64, up, h->comm, h->stream));
    NK(h->nccl.Send(v + L.off_z + pz * (2 * L.nzl), 2 * pz, ncclFloat64, up, h->comm, h->stream));
    NK(h->nccl.Recv(v + L.off_z + pz * (2 * L.nzl + 2), pz, ncclFloat64, up, h->comm, h->stream));
  }
  if (dn >= 0) {
    NK(h->nccl.Recv(v + L.off_p, pp, ncclFloat64, dn, h->comm, h->stream));
    NK(h->nccl.Recv(v + L.off_z, 2 * pz, ncclFloat64, dn, h->comm, h->stream));
    NK(h->nccl.Send(v + L.off_z + pz * 2, pz, ncclFloat64, dn, h->comm, h->stream));
  }
  NK(grp.end());
  return 0;
}
int allreduce_scalar(impes_ctx* h, double* p) {
  if (h->L.nranks == 1) return 0;
  Scope sc(h, KI_COMM, 1);
  NK(h->nccl.AllReduce(p, p, 1, ncclFloat64, ncclSum, h->comm, h->stream));
  return 0;
}

// ---- coefficient update (K4) -------------------------------------------------------------------
int mg_gather(impes_ctx* h, double* full, size_t pe);
MgLevel mg_view(const impes_ctx* h, const MgLevel& v);
#ifndef IMPES_APPLYB_RESIDENT
#define IMPES_APPLYB_RESIDENT 0   // 1: K1 pass B with the fused dot as a resident grid (measured: 2 % slower than one CTA per block)
#endif
int update_coeff(impes_ctx* h) {
  const Lay& L = h->L;
  RET_IF(exchange_planes(h, h->S[h->scur], true, true, h->nsl));
  if (h->os == 1 && h->ord == 0) {  // RT0 with S in DG-Q1: element-block entries by quadrature, then the RT0 face coefficients
    Scope sc(h, KI_COEFF, 2);
    const int g = (int)((L.ncv + BLK - 1) / BLK);
    double* a = h->aq;
    if (L.has_y) k0q_coeff<true><<<g, BLK, 0, h->stream>>>(L, h->ph, h->qt, h->S[h->scur], h->K[0], h->K[1], h->K[2], h->geo[0], h->geo[1], h->geo[2], a);
    else k0q_coeff<false><<<g, BLK, 0, h->stream>>>(L, h->ph, h->qt, h->S[h->scur], h->K[0], h->K[1], h->K[2], h->geo[0], h->geo[1], h->geo[2], a);
    k_face_coeff<true><<<h->grid_cells, BLK, 0, h->stream>>>(L, a, a + 3L * L.ncv, a + 6L * L.ncv, h->Md, h->T);
  } else
  if (h->os == 1) {  // S in DG-Q1: 1/lambda at the mass-rule points, diag(M) by quadrature, Schur diagonal
    const bool slab = L.nranks > 1;
    Scope sc(h, KI_COEFF, slab ? 5 : 4);
    const int g = (int)(((long)L.ncv * h->nqp + BLK - 1) / BLK);   // one thread per (cell, mass-rule point)
    const int ga = (L.ncl + K1Q_TC - 1) / K1Q_TC, gs = (2 * L.plane + BLK - 1) / BLK;
    if (L.has_y) {
      k1q_rlam<true><<<g, BLK, 0, h->stream>>>(L, h->ph, h->qt, h->S[h->scur], h->rlam);
      k1q_md_a<true><<<ga, K1QS<true>::THREADS, sizeof(double) * K1Q_TC * K1QS<true>::RS, h->stream>>>(h->L1, h->qt, h->rlam, h->K[0], h->K[1], h->K[2], h->geo[0], h->geo[1], h->geo[2], h->Md, h->T, h->yup);
      k1q_md_b<true><<<h->grid_cells, BLK, 0, h->stream>>>(h->L1, h->Md, h->T, h->yup);
      if (slab) k1q_md_slab<true><<<gs, BLK, 0, h->stream>>>(h->L1, h->qt, h->rlam, h->K[2], h->geo[2], h->Md, h->T);
      k1_schur_diag<true><<<h->grid_dofs1, BLK, 0, h->stream>>>(h->L1, h->rc, h->T, h->Md);
    } else {
      k1q_rlam<false><<<g, BLK, 0, h->stream>>>(L, h->ph, h->qt, h->S[h->scur], h->rlam);
      k1q_md_a<false><<<ga, K1QS<false>::THREADS, sizeof(double) * K1Q_TC * K1QS<false>::RS, h->stream>>>(h->L1, h->qt, h->rlam, h->K[0], h->K[1], h->K[2], h->geo[0], h->geo[1], h->geo[2], h->Md, h->T, h->yup);
      k1q_md_b<false><<<h->grid_cells, BLK, 0, h->stream>>>(h->L1, h->Md, h->T, h->yup);
      if (slab) k1q_md_slab<false><<<gs, BLK, 0, h->stream>>>(h->L1, h->qt, h->rlam, h->K[2], h->geo[2], h->Md, h->T);
      k1_schur_diag<false><<<h->grid_dofs1, BLK, 0, h->stream>>>(h->L1, h->rc, h->T, h->Md);
    }
  } else
  {
    Scope sc(h, KI_COEFF, 2);
    const int g = (int)((L.ncv + BLK - 1) / BLK);
    k_cell_coeff<<<g, BLK, 0, h->stream>>>(L, h->ph, h->S[h->scur], h->K[0], h->K[1], h->K[2], h->geo[0], h->geo[1],
                                           h->geo[2], h->m[0], h->m[1], h->m[2]);
    if (h->ord == 1) {
      if (L.has_y) k1_node_coeff<true><<<h->grid_cells, BLK, 0, h->stream>>>(h->L1, h->rc, h->m[0], h->m[1], h->m[2], h->Md, h->T);
      else k1_node_coeff<false><<<h->grid_cells, BLK, 0, h->stream>>>(h->L1, h->rc, h->m[0], h->m[1], h->m[2], h->Md, h->T);
    } else {
      k_face_coeff<false><<<h->grid_cells, BLK, 0, h->stream>>>(L, h->m[0], h->m[1], h->m[2], h->Md, h->T);
    }
  }
  if (h->sgs_skew) {
    Scope sc(h, KI_COEFF, 1);
    k_sgs_pack<<<h->grid_cells, BLK, 0, h->stream>>>(L, h->pack, h->T, h->Md);
  }
  if (!h->mg.empty()) {  // multigrid hierarchy of the Schur stencil (S depends on the saturation)
    auto grid_of = [](const MgLevel& v) { return (int)(((long)v.nx * v.ny * v.nz + MG_B - 1) / MG_B); };
    const bool multi = L.nranks > 1;
    {
      Scope sc(h, KI_COEFF, 1 + 2 * ((int)h->mg.size() - 1) + (h->ord == 1 ? 1 : 0));
      if (h->ord == 1) {
        if (L.has_y) k1_mg_copy_fine<true><<<h->grid_cells, BLK, 0, h->stream>>>(h->L1, h->rc, h->mg[0], h->T);
        else k1_mg_copy_fine<false><<<h->grid_cells, BLK, 0, h->stream>>>(h->L1, h->rc, h->mg[0], h->T);
        k_mg_diag<<<grid_of(h->mg[0]), MG_B, 0, h->stream>>>(h->mg[0]);
      } else
        k_mg_copy_fine<<<grid_of(h->mg[0]), MG_B, 0, h->stream>>>(L, h->mg[0], h->T, h->Md);
    }
    for (size_t l = 1; l < h->mg.size(); ++l) {
      if (multi && (int)l == h->mg_g) {  // coarsen my planes, assemble the level on every rank
        const MgLevel cv = mg_view(h, h->mg[l]);
        k_mg_coarsen<<<grid_of(cv), MG_B, 0, h->stream>>>(h->mg[l - 1], cv);
        RET_IF(mg_gather(h, h->mg[l].Tx, (size_t)(cv.nx + 1) * cv.ny));
        if (cv.ny > 1) RET_IF(mg_gather(h, h->mg[l].Ty, (size_t)cv.nx * (cv.ny + 1)));
        RET_IF(mg_gather(h, h->mg[l].Tz, (size_t)cv.nx * cv.ny));
      } else {
        k_mg_coarsen<<<grid_of(h->mg[l]), MG_B, 0, h->stream>>>(h->mg[l - 1], h->mg[l]);
      }
      k_mg_diag<<<grid_of(h->mg[l]), MG_B, 0, h->stream>>>(h->mg[l]);
    }
  }
  CK(cudaGetLastError());
  h->coeff_valid = true;
  return 0;
}

// ---- K1 ----------------------------------------------------------------------------------------
int darcy_apply(impes_ctx* h, double* x, double* y, bool dot, const MinresState* st) {
  if (h->os == 1 && h->ord == 1) {  // variable coefficient inside the cells: element products (A), then face-node completion (B)
    RET_IF(halo_darcy1(h, x));
    {
      Scope sc(h, KI_DARCY_Q1A, 1);
#define KQA(HY, DOT)                                                                                              \
  k1q_apply_a<HY, DOT><<<(h->L.ncl + K1Q_TC - 1) / K1Q_TC, K1QS<HY>::THREADS, sizeof(double) * K1QS<HY>::SMEM_DOUBLES, \
                         h->stream>>>(h->L1, h->rc, h->qt, h->rlam, h->K[0], h->K[1], h->K[2], h->geo[0], h->geo[1],  \
                                      h->geo[2], x, y, h->yup, h->accA, st)
      if (h->L.has_y) { if (dot) KQA(true, true); else KQA(true, false); }
      else { if (dot) KQA(false, true); else KQA(false, false); }
#undef KQA
    }
    RET_IF(exchange_planes(h, h->yup, true, false, h->nup));  // the lower rank's top-plane contributions
    {
      Scope sc(h, KI_DARCY_Q1B, 1);
#define KQB(HY, DOT)                                                                                               \
  k1q_apply_b<HY, DOT><<<(DOT && IMPES_APPLYB_RESIDENT) ? std::min(h->grid_cells, h->num_sms * 8) : h->grid_cells, BLK, 0, h->stream>>>( \
      h->L1, x, y, h->yup, h->accA, h->partial, h->scal + 0, h->ticket, st, (unsigned)h->grid_cells)
      if (h->L.has_y) { if (dot) KQB(true, true); else KQB(true, false); }
      else { if (dot) KQB(false, true); else KQB(false, false); }
#undef KQB
    }
    if (dot) RET_IF(allreduce_scalar(h, h->scal + 0));
    return 0;
  }
  if (h->ord == 1) {
    RET_IF(halo_darcy1(h, x));
    {
      Scope sc(h, KI_DARCY_RT1, 1);
      const int g = h->grid_dofs1;
#define K1_LAUNCH(HY, DOT)                                                                                        \
  k1_darcy_apply<HY, DOT><<<g, BLK, 0, h->stream>>>(h->L1, h->rc, h->m[0], h->m[1], h->m[2], x, y, h->partial, \
                                                    h->scal + 0, h->ticket, st)
      if (h->L.has_y) { if (dot) K1_LAUNCH(true, true); else K1_LAUNCH(true, false); }
      else { if (dot) K1_LAUNCH(false, true); else K1_LAUNCH(false, false); }
#undef K1_LAUNCH
    }
    if (dot) RET_IF(allreduce_scalar(h, h->scal + 0));
    return 0;
  }
  RET_IF(halo_darcy(h, x));
  {
    Scope sc(h, KI_DARCY_RT0, 1);
    if (h->os == 1) {  // DG-Q1 coefficient: the same kernel with the element-block entries of k0q_coeff
      const double *ax = h->aq, *ay = h->aq + 3L * h->L.ncv, *az = h->aq + 6L * h->L.ncv;
      if (dot) k_darcy_apply<true, true><<<h->grid_cells, BLK, 0, h->stream>>>(h->L, ax, ay, az, x, y, h->partial, h->scal + 0, h->ticket, st);
      else k_darcy_apply<false, true><<<h->grid_cells, BLK, 0, h->stream>>>(h->L, ax, ay, az, x, y, h->partial, h->scal + 0, h->ticket, st);
    } else if (dot)
      k_darcy_apply<true><<<h->grid_cells, BLK, 0, h->stream>>>(h->L, h->m[0], h->m[1], h->m[2], x, y, h->partial,
                                                                h->scal + 0, h->ticket, st);
    else
      k_darcy_apply<false><<<h->grid_cells, BLK, 0, h->stream>>>(h->L, h->m[0], h->m[1], h->m[2], x, y, h->partial,
                                                                 h->scal + 0, h->ticket, st);
  }
  if (dot) RET_IF(allreduce_scalar(h, h->scal + 0));
  return 0;
}

// ---- K3: Schur block --------------------------------------------------------------------------
template <int MODE>
int launch_skew_sweep(impes_ctx* h, double* y, int fwd, int bwd, const MinresState* st) {
  const Lay& L = h->L;
  const double* T = h->T;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(h->grid_sgs);
  cfg.blockDim = dim3(h->sgs_threads);
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = h->grid_sgs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = MODE == 1 ? 1 : 0;
  switch (h->sgs_ppt) {
    case 1: CK(cudaLaunchKernelEx(&cfg, k_sgs_skew<MODE, 1>, L, h->pack, T, y, fwd, bwd, st)); break;
    case 2: CK(cudaLaunchKernelEx(&cfg, k_sgs_skew<MODE, 2>, L, h->pack, T, y, fwd, bwd, st)); break;
    default: CK(cudaLaunchKernelEx(&cfg, k_sgs_skew<MODE, 4>, L, h->pack, T, y, fwd, bwd, st)); break;
  }
  return 0;
}

template <int NPS, bool HY>
void launch_sgs_tma_t(impes_ctx* h, double* y, int fwd, int bwd, const MinresState* st) {
  k_sgs_tma<NPS, HY><<<1, h->sgs_threads, h->sgs_smem, h->stream>>>(h->L, h->pack, h->T, y, h->sgs_ns, h->sgs_yp, fwd, bwd, st);
}
void launch_sgs_tma(impes_ctx* h, double* y, int fwd, int bwd, const MinresState* st) {
  const bool hy = h->L.has_y != 0;
  switch (h->pack.NPs) {
    case 64: hy ? launch_sgs_tma_t<64, true>(h, y, fwd, bwd, st) : launch_sgs_tma_t<64, false>(h, y, fwd, bwd, st); break;
    case 128: hy ? launch_sgs_tma_t<128, true>(h, y, fwd, bwd, st) : launch_sgs_tma_t<128, false>(h, y, fwd, bwd, st); break;
    case 256: hy ? launch_sgs_tma_t<256, true>(h, y, fwd, bwd, st) : launch_sgs_tma_t<256, false>(h, y, fwd, bwd, st); break;
    case 512: hy ? launch_sgs_tma_t<512, true>(h, y, fwd, bwd, st) : launch_sgs_tma_t<512, false>(h, y, fwd, bwd, st); break;
    default: hy ? launch_sgs_tma_t<0, true>(h, y, fwd, bwd, st) : launch_sgs_tma_t<0, false>(h, y, fwd, bwd, st); break;
  }
}

template <bool HY>
int launch_sgs1(impes_ctx* h, const double* r, double* y, int fwd, int bwd, const MinresState* st) {
  Lay1 L = h->L1;
  Rt1Const C = h->rc;
  const double* T = h->T;
  const double* Md = h->Md;
  Scope sc(h, KI_SGS, 1);
  if (h->sgs_mode == 2) {
    void* args[] = {(void*)&L, (void*)&C, (void*)&T, (void*)&Md, (void*)&r, (void*)&y, (void*)&fwd, (void*)&bwd, (void*)&st};
    CK(cudaLaunchCooperativeKernel((void*)k1_sgs_wavefront<2, HY>, dim3(h->grid_sgs), dim3(h->sgs_threads), args, 0, h->stream));
  } else if (h->sgs_mode == 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h->grid_sgs);
    cfg.blockDim = dim3(h->sgs_threads);
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = h->grid_sgs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k1_sgs_wavefront<1, HY>, L, C, T, Md, r, y, fwd, bwd, st));
  } else {
    k1_sgs_wavefront<0, HY><<<1, h->sgs_threads, 0, h->stream>>>(L, C, T, Md, r, y, fwd, bwd, st);
  }
  return 0;
}

int launch_sgs(impes_ctx* h, const double* r, double* y, int fwd, int bwd, const MinresState* st) {
  if (h->ord == 1) return h->L.has_y ? launch_sgs1<true>(h, r, y, fwd, bwd, st) : launch_sgs1<false>(h, r, y, fwd, bwd, st);
  Lay L = h->L;
  const double* T = h->T;
  const double* Md = h->Md;
  if (h->sgs_skew) {
    Scope sc(h, KI_SGS, 2 + (fwd ? 1 : 0));
    if (fwd) k_sgs_skew_in<<<h->grid_cells, BLK, 0, h->stream>>>(L, h->pack, r, st);
    if (h->sgs_tma)
      launch_sgs_tma(h, y, fwd, bwd, st);
    else if (h->sgs_mode == 0) RET_IF(launch_skew_sweep<0>(h, y, fwd, bwd, st));
    else RET_IF(launch_skew_sweep<1>(h, y, fwd, bwd, st));
    k_sgs_skew_out<<<h->grid_cells, BLK, 0, h->stream>>>(L, h->pack, bwd, y, st);
    return 0;
  }
  Scope sc(h, KI_SGS, 1);
  if (h->sgs_mode == 2) {
    void* args[] = {(void*)&L, (void*)&T, (void*)&Md, (void*)&r, (void*)&y, (void*)&fwd, (void*)&bwd, (void*)&st};
    CK(cudaLaunchCooperativeKernel((void*)k_sgs_wavefront<2>, dim3(h->grid_sgs), dim3(h->sgs_threads), args, 0, h->stream));
  } else if (h->sgs_mode == 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h->grid_sgs);
    cfg.blockDim = dim3(h->sgs_threads);
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = h->grid_sgs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k_sgs_wavefront<1>, L, T, Md, r, y, fwd, bwd, st));
  } else {
    k_sgs_wavefront<0><<<1, h->sgs_threads, 0, h->stream>>>(L, T, Md, r, y, fwd, bwd, st);
  }
  return 0;
}

// ---- multigrid plumbing for slabs ------------------------------------------------------------
// ghost planes of a slab-distributed level buffer (buf = local plane 0; ghosts at buf - pe and buf + nz*pe)
int mg_halo(impes_ctx* h, double* buf, size_t pe, int nz) {
  const Lay& L = h->L;
  if (L.nranks == 1) return 0;
  const int up = L.rank + 1 < L.nranks ? L.rank + 1 : -1, dn = L.rank > 0 ? L.rank - 1 : -1;
  Scope sc(h, KI_COMM, 1);
  NcclGroup grp(h->nccl);
  if (!grp.open) { h->err = "ncclGroupStart failed"; return IMPES_ERR_NCCL; }
  if (up >= 0) {
    NK(h->nccl.Send(buf + pe * (size_t)(nz - 1), pe, ncclFloat64, up, h->comm, h->stream));
    NK(h->nccl.Recv(buf + pe * (size_t)nz, pe, ncclFloat64, up, h->comm, h->stream));
  }
  if (dn >= 0) {
    NK(h->nccl.Send(buf, pe, ncclFloat64, dn, h->comm, h->stream));
    NK(h->nccl.Recv(buf - pe, pe, ncclFloat64, dn, h->comm, h->stream));
  }
  NK(grp.end());
  return 0;
}
// every rank contributes planes [off[r], off[r]+cnt[r]) of `full` (pe doubles per plane); afterwards all
// ranks hold the whole array
int mg_gather(impes_ctx* h, double* full, size_t pe) {
  const Lay& L = h->L;
  if (L.nranks == 1) return 0;
  Scope sc(h, KI_COMM, 1);
  NcclGroup grp(h->nccl);
  if (!grp.open) { h->err = "ncclGroupStart failed"; return IMPES_ERR_NCCL; }
  for (int r = 0; r < L.nranks; ++r) {
    if (r == L.rank) continue;
    NK(h->nccl.Send(full + pe * (size_t)h->mg_goff[L.rank], pe * (size_t)h->mg_gcnt[L.rank], ncclFloat64, r, h->comm, h->stream));
    NK(h->nccl.Recv(full + pe * (size_t)h->mg_goff[r], pe * (size_t)h->mg_gcnt[r], ncclFloat64, r, h->comm, h->stream));
  }
  NK(grp.end());
  return 0;
}
// this rank's part of the first replicated level, addressed like a distributed level
MgLevel mg_view(const impes_ctx* h, const MgLevel& v) {
  MgLevel w = v;
  const int zo = h->mg_goff[h->L.rank];
  const long pl = (long)v.nx * v.ny;
  w.nz = h->mg_gcnt[h->L.rank];
  w.zoff = zo;
  w.Tx += (long)(v.nx + 1) * v.ny * zo;
  w.Ty += (long)v.nx * (v.ny + 1) * zo;
  w.Tz += pl * zo;
  w.D += pl * zo;
  if (w.r) w.r += pl * zo;
  if (w.ya) w.ya += pl * zo;
  if (w.yb) w.yb += pl * zo;
  return w;
}

// One V(nu,nu) cycle, zero initial guess: y_p = MG(r_p)  (mg.cuh).  Iterates alternate between ya and yb of
// each level; all sweep counts are even, so every level's result ends in yb (level 0: the p-part of y).
// Multi-GPU: levels < mg_g live on slabs (ghost planes refreshed before every sweep that reads neighbours),
// level mg_g is assembled on every rank by a gather and the levels below it run replicated -- the same
// arithmetic as on one GPU, so the result does not depend on the number of GPUs.
int mg_vcycle(impes_ctx* h, const double* r, double* y, const MinresState* st, bool dot = false) {
  const Lay& L = h->L;
  std::vector<MgLevel>& mg = h->mg;
  const int nl = (int)mg.size();
  const bool q1 = h->ord == 1;
  const bool multi = L.nranks > 1;
  // Q1 level (ord == 1): p-parts of the Krylov vectors; iterates alternate between qa and qb like every level
  const size_t qpe = q1 ? (size_t)L.plane * h->L1.npl : 0;
  const double* qr = q1 ? r + h->L1.off_p : nullptr;
  double* qb = q1 ? y + h->L1.off_p : nullptr;
  double* qa = h->mg1_ya;
  const double* qD = q1 ? h->Md + h->L1.off_p : nullptr;
  const int gq = q1 ? (int)(((long)L.ncl * h->L1.npl + BLK - 1) / BLK) : 0;  // 256-dof blocks (one thread per pair of L2 dofs)
  const int gdot = std::min(gq, h->num_sms * IMPES_MGQ_MINB2);               // resident grid of the sweep with the fused dot
  auto q1_smooth = [&](int sweeps, bool from_zero) -> int {
    for (int s = 0; s < sweeps; ++s) {
      double* in = (s & 1) ? qa : qb;
      double* out = (s & 1) ? qb : qa;
      const int first = from_zero && s == 0;
      const double om = from_zero ? h->om_pre(s) : h->om_post(s, sweeps);
      if (!first && multi) RET_IF(mg_halo(h, in + qpe, qpe, L.nzl));
      Scope sc(h, first ? KI_MGQ_FIRST : KI_MGQ_JACOBI, 1);
      if (first) {  // owned L2 dofs = one contiguous range of the p-part (16-byte aligned: npl is even, parts are padded)
        const long n2 = (long)L.ncl * h->L1.npl / 2;
        const size_t o2 = qpe / 2;
        k1_mg_first<<<h->num_sms * 8, BLK, 0, h->stream>>>(n2, reinterpret_cast<const double2*>(qD) + o2, reinterpret_cast<const double2*>(qr) + o2,
                                                           reinterpret_cast<double2*>(out) + o2, om, st);
      } else if (L.has_y) k1_mg_jacobi<true><<<gq, MGQ_T, 0, h->stream>>>(h->L1, h->rc, h->T, qD, qr, in, out, om, 0, st, (unsigned)gq);
      else k1_mg_jacobi<false><<<gq, MGQ_T, 0, h->stream>>>(h->L1, h->rc, h->T, qD, qr, in, out, om, 0, st, (unsigned)gq);
    }
    return 0;
  };
  if (!q1) {
    mg[0].r = const_cast<double*>(r) + L.off_p + L.plane;
    mg[0].yb = y + L.off_p + L.plane;
  }
  auto grid_of = [](const MgLevel& v) { return (int)(((long)v.nx * v.ny * v.nz + MG_B - 1) / MG_B); };
  auto dist = [&](int l) { return multi && l < h->mg_g; };
  auto smooth = [&](int l, int sweeps, bool from_zero) -> int {
    MgLevel& v = mg[l];
    for (int s = 0; s < sweeps; ++s) {
      // sweep s reads (s even ? yb : ya) and writes (s even ? ya : yb); with from_zero the first read is skipped
      double* in = (s & 1) ? v.ya : v.yb;
      double* out = (s & 1) ? v.yb : v.ya;
      const bool first = from_zero && s == 0;
      if (!first && dist(l)) RET_IF(mg_halo(h, in, (size_t)v.nx * v.ny, v.nz));
      Scope sc(h, KI_MG_CELL, 1);
      const double om = sweeps > 8 ? h->mg_omega : (from_zero ? h->om_pre(s) : h->om_post(s, sweeps));  // > 8: the coarsest level
      k_mg_jacobi<<<grid_of(v), MG_B, 0, h->stream>>>(v, v.r, in, out, om, first, st);
    }
    return 0;
  };
  if (q1) {
    RET_IF(q1_smooth(h->mg_nu, true));
    if (multi) RET_IF(mg_halo(h, qb + qpe, qpe, L.nzl));
    Scope sc(h, KI_MGQ_RESTRICT, 1);
    if (L.has_y) k1_mg_restrict<true><<<gq, MGQ_T, 0, h->stream>>>(h->L1, h->rc, h->T, qD, qr, qb, mg[0].r, st);
    else k1_mg_restrict<false><<<gq, MGQ_T, 0, h->stream>>>(h->L1, h->rc, h->T, qD, qr, qb, mg[0].r, st);
  }
  for (int l = 0; l + 1 < nl; ++l) {
    RET_IF(smooth(l, h->mg_nu, true));
    if (dist(l)) RET_IF(mg_halo(h, mg[l].yb, (size_t)mg[l].nx * mg[l].ny, mg[l].nz));
    if (multi && l + 1 == h->mg_g) {  // my planes of the first replicated level, then assemble it everywhere
      const MgLevel cv = mg_view(h, mg[l + 1]);
      {
        Scope sc(h, KI_MG_CELL, 1);
        k_mg_restrict<<<grid_of(cv), MG_B, 0, h->stream>>>(mg[l], cv, mg[l].r, mg[l].yb, cv.r, st);
      }
      RET_IF(mg_gather(h, mg[l + 1].r, (size_t)cv.nx * cv.ny));
    } else {
      Scope sc(h, KI_MG_CELL, 1);
      k_mg_restrict<<<grid_of(mg[l + 1]), MG_B, 0, h->stream>>>(mg[l], mg[l + 1], mg[l].r, mg[l].yb, mg[l + 1].r, st);
    }
  }
  RET_IF(smooth(nl - 1, h->mg_coarse_sweeps, true));
  for (int l = nl - 2; l >= 0; --l) {
    const MgLevel cv = (multi && l + 1 == h->mg_g) ? mg_view(h, mg[l + 1]) : mg[l + 1];
    {
      Scope sc(h, KI_MG_CELL, 1);
      k_mg_prolong<<<grid_of(mg[l]), MG_B, 0, h->stream>>>(mg[l], cv, cv.yb, mg[l].yb, h->mg_cfac, st);
    }
    RET_IF(smooth(l, h->mg_nu, false));
  }
  if (q1) {
    // prolongation fused into the first post-smoothing sweep (k1_mg_jacobi<.., PRO>): reads qb (before the
    // correction) and the level-0 iterate, writes qa; the remaining sweeps alternate as before.  Needs the ghost
    // planes of both: qb's are still valid from the exchange before the restriction, mg[0].yb's are fetched here.
    if (multi) RET_IF(mg_halo(h, mg[0].yb, (size_t)L.plane, L.nzl));
    {
      Scope sc(h, KI_MGQ_JACOBI, 1);
      if (L.has_y) k1_mg_jacobi<true, true><<<gq, MGQ_T, 0, h->stream>>>(h->L1, h->rc, h->T, qD, qr, qb, qa, h->om_post(0, h->mg_nu), 0, st, (unsigned)gq, mg[0].yb, h->mg_cfacq);
      else k1_mg_jacobi<false, true><<<gq, MGQ_T, 0, h->stream>>>(h->L1, h->rc, h->T, qD, qr, qb, qa, h->om_post(0, h->mg_nu), 0, st, (unsigned)gq, mg[0].yb, h->mg_cfacq);
    }
    for (int s = 1; s < h->mg_nu; ++s) {
      double* in = (s & 1) ? qa : qb;
      double* out = (s & 1) ? qb : qa;
      if (multi) RET_IF(mg_halo(h, in + qpe, qpe, L.nzl));
      Scope sc(h, KI_MGQ_JACOBI, 1);
      const double om = h->om_post(s, h->mg_nu);
      if (dot && s + 1 == h->mg_nu) {  // last sweep: finishes beta^2 = <v0, P v0> (see k1_mg_jacobi<.., DOT>)
        if (L.has_y) k1_mg_jacobi<true, false, true><<<gdot, MGQ_T, 0, h->stream>>>(h->L1, h->rc, h->T, qD, qr, in, out, om, 0, st, (unsigned)gq, nullptr, 0.0, h->partial, h->partial_u, (unsigned)h->n_u_chunks, h->scal + 1, h->ticket);
        else k1_mg_jacobi<false, false, true><<<gdot, MGQ_T, 0, h->stream>>>(h->L1, h->rc, h->T, qD, qr, in, out, om, 0, st, (unsigned)gq, nullptr, 0.0, h->partial, h->partial_u, (unsigned)h->n_u_chunks, h->scal + 1, h->ticket);
      } else {
        if (L.has_y) k1_mg_jacobi<true><<<gq, MGQ_T, 0, h->stream>>>(h->L1, h->rc, h->T, qD, qr, in, out, om, 0, st, (unsigned)gq);
        else k1_mg_jacobi<false><<<gq, MGQ_T, 0, h->stream>>>(h->L1, h->rc, h->T, qD, qr, in, out, om, 0, st, (unsigned)gq);
      }
    }
  }
  return 0;
}

// y_p = SchurPrec(r_p).  Parity mode: lexicographic symmetric GS; across slabs the sweep is pipelined
// rank by rank (exact global ordering, so results do not depend on the number of GPUs).
int schur_prec(impes_ctx* h, const double* r, double* y, const MinresState* st, bool dot = false) {
  const Lay& L = h->L;
  if (h->d.schur_prec == IMPES_SCHUR_MG) return mg_vcycle(h, r, y, st, dot);
  if (L.nranks == 1) return launch_sgs(h, r, y, 1, 1, st);
  const int up = L.rank + 1 < L.nranks ? L.rank + 1 : -1, dn = L.rank > 0 ? L.rank - 1 : -1;
  double* yp = y + h->vec_offp;
  const size_t pp = (size_t)L.plane * (h->ord == 1 ? h->L1.npl : 1);  // one cell plane of pressure dofs
  // forward: wait for the plane below, sweep, hand my top plane up
  if (dn >= 0) { Scope sc(h, KI_COMM, 1); NK(h->nccl.Recv(yp, pp, ncclFloat64, dn, h->comm, h->stream)); }
  RET_IF(launch_sgs(h, r, y, 1, 0, st));
  if (up >= 0) { Scope sc(h, KI_COMM, 1); NK(h->nccl.Send(yp + pp * L.nzl, pp, ncclFloat64, up, h->comm, h->stream)); }
  // backward: wait for the plane above, sweep, hand my bottom plane down
  if (up >= 0) { Scope sc(h, KI_COMM, 1); NK(h->nccl.Recv(yp + pp * (L.nzl + 1), pp, ncclFloat64, up, h->comm, h->stream)); }
  RET_IF(launch_sgs(h, r, y, 0, 1, st));
  if (dn >= 0) { Scope sc(h, KI_COMM, 1); NK(h->nccl.Send(yp + pp, pp, ncclFloat64, dn, h->comm, h->stream)); }
  return 0;
}

int dot_owned(impes_ctx* h, const double* a, const double* b, double* out, const MinresState* st) {
  {
    Scope sc(h, KI_DOT, 1);
    if (h->ord == 1) {
      if (h->L.has_y) k1_dot_owned<true><<<h->grid_dofs1, BLK, 0, h->stream>>>(h->L1, a, b, h->partial, out, h->ticket, st);
      else k1_dot_owned<false><<<h->grid_dofs1, BLK, 0, h->stream>>>(h->L1, a, b, h->partial, out, h->ticket, st);
    } else
      k_dot_owned<<<h->grid_cells, BLK, 0, h->stream>>>(h->L, a, b, h->partial, out, h->ticket, st);
  }
  return allreduce_scalar(h, out);
}

// one MINRES iteration with the given vector roles; reads st_cur, writes st_next
int minres_iteration(impes_ctx* h, double* u1, double* q, double* v0, double* v1, double* w0, double* w1,
                     const MinresState* st_cur, MinresState* st_next) {
  RET_IF(darcy_apply(h, u1, q, true, st_cur));
  if (h->fold_dot) {
    // beta^2 = <v0, P v0> without a dot kernel: velocity part inside pass B, pressure part + final sum inside the last
    // smoothing sweep of the V-cycle
    {
      Scope sc(h, KI_MINRES_B, 1);
      const Lay1& Q = h->L1;
      const long pz = (long)Q.gz0 * Q.gz1;
      const long own_z0 = Q.off_z + 2 * pz, own_z1 = own_z0 + pz * (2 * Q.nzl + (Q.zhi == Q.NZ ? 1 : 0));
      k_minres_b_dot<<<h->n_u_chunks_all, MINRES_B_T, 0, h->stream>>>(h->vec_n, h->vec_offp, q, v0, v1, h->Md, q, h->scal + 0, st_cur,
                                                               (long)Q.off_z, own_z0, own_z1, (long)Q.off_p, h->partial_u);
    }
    RET_IF(schur_prec(h, v0, q, st_cur, true));
    RET_IF(allreduce_scalar(h, h->scal + 1));
  } else {
    {
      Scope sc(h, KI_MINRES_B, 1);
      k_minres_b<<<h->grid_vec, BLK, 0, h->stream>>>(h->vec_n, h->vec_offp, q, v0, v1, h->Md, q, h->scal + 0, st_cur);
    }
    RET_IF(schur_prec(h, v0, q, st_cur));
    RET_IF(dot_owned(h, v0, q, h->scal + 1, st_cur));
  }
  {
    Scope sc(h, KI_MINRES_C, 1);
    if (h->lazy_u)
      k_minres_c<true><<<h->grid_vec, BLK, 0, h->stream>>>(h->vec_n, w0, w1, u1, h->x, v0, q, h->scal + 0, h->scal + 1, st_cur,
                                                           st_next, h->d.darcy_maxiter);
    else
      k_minres_c<false><<<h->grid_vec, BLK, 0, h->stream>>>(h->vec_n, w0, w1, u1, h->x, v0, q, h->scal + 0, h->scal + 1, st_cur,
                                                            st_next, h->d.darcy_maxiter);
  }
  return 0;
}

int minres_pair(impes_ctx* h) {
  RET_IF(minres_iteration(h, h->u1, h->q, h->v0, h->v1, h->w0, h->w1, h->st + 0, h->st + 1));
  RET_IF(minres_iteration(h, h->q, h->u1, h->v1, h->v0, h->w1, h->w0, h->st + 1, h->st + 0));
  return 0;
}

int build_minres_graph(impes_ctx* h) {
  // multi-GPU: NCCL send/recv/all-reduce are capturable (validated on 2 and 8 B200: results unchanged, ~2 % faster);
  // IMPES_GRAPH_MULTI=0 turns it off.  Every rank replays the same graph: `done` derives from all-reduced scalars.
  static const bool graph_multi = !(getenv("IMPES_GRAPH_MULTI") && atoi(getenv("IMPES_GRAPH_MULTI")) == 0);
  if (h->minres_exec || !h->use_graph || (h->L.nranks > 1 && !graph_multi) || h->profiling) return 0;
  const int pairs = 8;
  cudaGraph_t g = nullptr;
  const long long l0 = h->launches;
  long long n0[KC_N], k0[KI_N];
  memcpy(n0, h->prof_n, sizeof n0);
  memcpy(k0, h->ki_n, sizeof k0);
  if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    h->use_graph = false;
    return 0;
  }
  int rc = 0;
  for (int p = 0; p < pairs && !rc; ++p) rc = minres_pair(h);
  cudaError_t e = cudaStreamEndCapture(h->stream, &g);
  h->graph_launches = h->launches - l0;  // what one replay of the graph launches
  for (int c = 0; c < KC_N; ++c) h->graph_prof_n[c] = h->prof_n[c] - n0[c];
  for (int c = 0; c < KI_N; ++c) h->graph_ki_n[c] = h->ki_n[c] - k0[c];
  h->launches = l0;  // capture is not execution
  memcpy(h->prof_n, n0, sizeof n0);
  memcpy(h->ki_n, k0, sizeof k0);
  if (rc || e != cudaSuccess || !g) {
    cudaGetLastError();
    if (g) cudaGraphDestroy(g);
    h->use_graph = false;  // e.g. cooperative launch not capturable: fall back to direct launches
    return 0;
  }
  if (cudaGraphInstantiate(&h->minres_exec, g, 0) != cudaSuccess) {
    cudaGetLastError();
    h->minres_exec = nullptr;
    h->use_graph = false;
  }
  cudaGraphDestroy(g);
  h->graph_iters = 2 * pairs;
  return 0;
}

int pressure_solve_device(impes_ctx* h, impes_solve_stats* stats) {
  auto t0 = std::chrono::steady_clock::now();
  struct EventPair {  // destroyed on every return path
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  } ev;
  CK(cudaEventCreate(&ev.a));
  CK(cudaEventCreate(&ev.b));
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  CK(cudaEventRecord(e0, h->stream));
  RET_IF(update_coeff(h));
  const size_t nb = sizeof(double) * (size_t)h->vec_n;
  CK(cudaMemsetAsync(h->x, 0, nb, h->stream));   // x = 0  [REF src/pressure_solver.cpp:101]
  CK(cudaMemsetAsync(h->v0, 0, nb, h->stream));
  CK(cudaMemsetAsync(h->w0, 0, nb, h->stream));
  CK(cudaMemsetAsync(h->w1, 0, nb, h->stream));
  CK(cudaMemsetAsync(h->q, 0, nb, h->stream));
  CK(cudaMemsetAsync(h->u1, 0, nb, h->stream));
  CK(cudaMemcpyAsync(h->v1, h->rhs, nb, cudaMemcpyDeviceToDevice, h->stream));  // v1 = b - A*0
  {
    Scope sc(h, KI_JACOBI_U, 1);
    k_jacobi_u<<<h->grid_vec, BLK, 0, h->stream>>>(h->vec_offp, h->v1, h->Md, h->u1);
  }
  RET_IF(schur_prec(h, h->v1, h->u1, nullptr));
  RET_IF(dot_owned(h, h->u1, h->v1, h->scal + 1, nullptr));
  {
    Scope sc(h, KI_MINRES_INIT, 1);
    k_minres_init<<<h->grid_vec, BLK, 0, h->stream>>>(h->vec_n, h->v1, h->u1, h->scal + 1, h->st + 0, h->d.darcy_rtol,
                                                      h->d.darcy_atol, h->lazy_u ? 1 : 0);
  }
  CK(cudaGetLastError());
  RET_IF(build_minres_graph(h));
  MinresState hs;
  for (;;) {
    CK(cudaMemcpyAsync(&hs, h->st + 0, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (hs.done) break;
    if (h->d.darcy_print_level == 1) {
      // MFEM prints the residual of EVERY iteration [MFEM-recall: MINRESSolver::Mult, print_level 1]: no batching --
      // one pair of iterations per host check, both states read back
      if (h->L.rank == 0 && hs.it == 0) printf("MINRES: iteration %3d: ||r||_B = %g\n", 0, fabs(hs.eta));
      RET_IF(minres_pair(h));
      MinresState h1;
      CK(cudaMemcpyAsync(&h1, h->st + 1, sizeof h1, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(&hs, h->st + 0, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      if (h->L.rank == 0) {
        printf("MINRES: iteration %3d: ||r||_B = %g\n", h1.it, fabs(h1.eta));
        if (hs.it > h1.it) printf("MINRES: iteration %3d: ||r||_B = %g\n", hs.it, fabs(hs.eta));
      }
      continue;
    }
    if (h->minres_exec && !h->profiling) {
      CK(cudaGraphLaunch(h->minres_exec, h->stream));
      h->launches += h->graph_launches;
      for (int c = 0; c < KC_N; ++c) h->prof_n[c] += h->graph_prof_n[c];
      for (int c = 0; c < KI_N; ++c) h->ki_n[c] += h->graph_ki_n[c];
    } else {
      for (int p = 0; p < 4; ++p) RET_IF(minres_pair(h));
    }
    CK(cudaGetLastError());
  }
  RET_IF(h->ord == 1 ? halo_darcy1(h, h->x) : halo_darcy(h, h->x));  // ghosts of the result (ghost faces feed the saturation step)
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  if (stats) {
    stats->iterations = hs.it;
    stats->converged = hs.done == 1;
    stats->final_norm = fabs(hs.eta);
    stats->initial_norm = hs.eta0;
    stats->device_ms = ms;
    stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  if (hs.done != 1) {
    char b[256];
    snprintf(b, sizeof b, "Darcy solver DIDN'T converge. Iter = %d. Final norm = %g", hs.it, fabs(hs.eta));
    h->err = b;
    return IMPES_ERR_NOT_CONVERGED;
  }
  return 0;
}

// ---- K6 driver ---------------------------------------------------------------------------------
int n_substeps(const impes_ctx* h, double global_dt, double dt) {
  if (h->d.substep_rule == IMPES_SUBSTEPS_PARALLEL) {
    int n = 0;
    double t = 0.0;
    while (!(t >= global_dt - dt / 2)) { t += dt; ++n; }  // [REF src/par_saturation_solver.cpp:149-155]
    return n;
  }
  return (int)(global_dt / dt);  // [REF src/saturation_solver.cpp:74-75]
}

// f(S) over the whole cvec (after the ghost planes of S are in place)
int launch_frac(impes_ctx* h, const double* S, double* F) {
  Scope sc(h, KI_SAT_AUX, 1);
  if (h->os == 1) {
    k_frac_flat<<<h->grid_vec, BLK, 0, h->stream>>>((long)h->L.ncv, h->nsl, h->ph, h->sp.two_phase, S, F);
    return 0;
  }
  const int g = (int)((h->L.ncv + BLK - 1) / BLK);
  k_frac_flow<<<g, BLK, 0, h->stream>>>(h->L, h->ph, h->sp.two_phase, S, F);
  return 0;
}

// os == 1: velocity records of the stage kernel from the RT1 dofs in h->x (once per step: the velocity is frozen)
int launch_sat_pre(impes_ctx* h) {
  Scope sc(h, KI_SAT_AUX, 1);
  if (h->ord == 0) {  // records from the RT0 face fluxes
    if (h->L.has_y) k0q_sat_pre<true><<<h->grid_cells, BLK, 0, h->stream>>>(h->L, h->qt, h->x, (long)h->L.ncv, h->vhq, h->fnq);
    else k0q_sat_pre<false><<<h->grid_cells, BLK, 0, h->stream>>>(h->L, h->qt, h->x, (long)h->L.ncv, h->vhq, h->fnq);
    return 0;
  }
  const int g = (h->L.ncl + K1Q_TC - 1) / K1Q_TC;
  if (h->L.has_y) k1q_sat_pre<true><<<g, SatQ<true>::THREADS, 0, h->stream>>>(h->L1, h->qt, h->x, (long)h->L.ncv, h->vhq, h->fnq);
  else k1q_sat_pre<false><<<g, SatQ<false>::THREADS, 0, h->stream>>>(h->L1, h->qt, h->x, (long)h->L.ncv, h->vhq, h->fnq);
  return 0;
}

// one K6 launch: rate = false: Sout = Sin + dt*F(Sin);  rate = true: Sout = F(Sin)
// Fin / Fout: f(S) arrays of the streaming form (Fin must hold f(Sin) incl. ghost planes); ignored otherwise.
int launch_sat(impes_ctx* h, bool rate, const double* Sin, double* Sout, const double* Fin = nullptr,
               double* Fout = nullptr) {
  Scope sc(h, h->os == 1 ? KI_SAT_Q1 : KI_SAT_P0, 1);
  const Lay& L = h->L;
  const double* satv = h->ord == 1 ? h->flux0 : h->x;  // face fluxes in the RT0 dvec layout
  if (h->os == 1) {  // DG-Q1: volume + face terms straight from the RT1 dofs (rt1q.cuh)
    const int g = (L.ncl + K1Q_TC - 1) / K1Q_TC;
    if (L.has_y) {
      if (rate) k1q_sat<true, true><<<g, K1Q_TC * 3, 0, h->stream>>>(h->L1, h->qt, h->ph, h->sp, (long)L.ncv, h->vhq, h->fnq, h->phi, Sin, Fin, Sout, Fout);
      else k1q_sat<true, false><<<g, K1Q_TC * 3, 0, h->stream>>>(h->L1, h->qt, h->ph, h->sp, (long)L.ncv, h->vhq, h->fnq, h->phi, Sin, Fin, Sout, Fout);
    } else {
      if (rate) k1q_sat<false, true><<<g, K1Q_TC * 2, 0, h->stream>>>(h->L1, h->qt, h->ph, h->sp, (long)L.ncv, h->vhq, h->fnq, h->phi, Sin, Fin, Sout, Fout);
      else k1q_sat<false, false><<<g, K1Q_TC * 2, 0, h->stream>>>(h->L1, h->qt, h->ph, h->sp, (long)L.ncv, h->vhq, h->fnq, h->phi, Sin, Fin, Sout, Fout);
    }
    return 0;
  }
  // DG_0: streaming form, f(S) carried as a second resident array (round 1's tiled / untiled alternatives measured slower
  // and were removed)
  if (rate) k_sat_stream<true><<<h->grid_cells, BLK, 0, h->stream>>>(L, h->ph, h->sp, satv, h->phi, Sin, Fin, Sout, Fout);
  else k_sat_stream<false><<<h->grid_cells, BLK, 0, h->stream>>>(L, h->ph, h->sp, satv, h->phi, Sin, Fin, Sout, Fout);
  return 0;
}

int sat_rate(impes_ctx* h, double* Sin, double* out) {  // out = M_phi^-1 (Kt f(Sin) + b)
  RET_IF(exchange_planes(h, Sin, true, true, h->nsl));
  RET_IF(launch_frac(h, Sin, h->F[0]));
  return launch_sat(h, true, Sin, out, h->F[0], nullptr);
}
int axpby(impes_ctx* h, double a, const double* x, double b, const double* y, double* out) {
  Scope sc(h, KI_SAT_AUX, 1);
  k_axpby<<<h->grid_vec, BLK, 0, h->stream>>>((long)h->L.ncv * h->nsl, a, x, b, y, out);
  return 0;
}
int xpby(impes_ctx* h, const double* x, double b, const double* y, double* out) {
  Scope sc(h, KI_SAT_AUX, 1);
  k_xpby<<<h->grid_vec, BLK, 0, h->stream>>>((long)h->L.ncv * h->nsl, x, b, y, out);
  return 0;
}

int saturation_device(impes_ctx* h, double global_dt, double dt, int* nsub) {
  const int nt = n_substeps(h, global_dt, dt);
  if (nsub) *nsub = nt;
  h->sp.dt = dt;
  if (h->ord == 1 && h->os == 0) {  // RT1 velocity -> face fluxes once per call (the velocity is frozen during the sub-steps)
    Scope sc(h, KI_SAT_AUX, 1);
    if (h->L.has_y) k1_face_flux<true><<<h->grid_cells, BLK, 0, h->stream>>>(h->L1, h->L, h->rc, h->x, h->flux0);
    else k1_face_flux<false><<<h->grid_cells, BLK, 0, h->stream>>>(h->L1, h->L, h->rc, h->x, h->flux0);
  }
  if (h->os == 1) RET_IF(launch_sat_pre(h));
  const int ode = h->d.substep_rule == IMPES_SUBSTEPS_SERIAL ? 1 : h->d.ode_solver_type;  // serial ignores -ode
  if (ode != 1 && ode != 2 && ode != 3 && ode != 4) {
    char b[128];
    snprintf(b, sizeof b, "Uknown ODE solver: type %d", ode);  // [REF src/par_saturation_solver.cpp:121-123]
    h->err = b;
    return IMPES_ERR_UNSUPPORTED;
  }
  if (ode != 1)
    for (int s = 0; s < 4; ++s)
      if (!h->stage[s]) RET_IF(dev_alloc(h, &h->stage[s], (long)h->L.ncv * h->nsl));
  for (int ti = 1; ti <= nt; ++ti) {
    double* S = h->S[h->scur];
    double* Sn = h->S[h->scur ^ 1];
    if (ode == 1) {
      // f(S) travels with S: F[scur] = f(S[scur]); only its ghost planes are exchanged per sub-step
      if (ti == 1) {
        RET_IF(exchange_planes(h, S, true, true, h->nsl));
        RET_IF(launch_frac(h, S, h->F[h->scur]));
      } else {
        if (h->os == 1) RET_IF(exchange_planes_dofmajor(h, h->F[h->scur], h->nsl));   // F[m ncv + c]
        else RET_IF(exchange_planes(h, h->F[h->scur], true, true, h->nsl));
      }
      RET_IF(launch_sat(h, false, S, Sn, h->F[h->scur], h->F[h->scur ^ 1]));
      h->scur ^= 1;
    } else if (ode == 2) {  // RK2Solver(a = 1)
      const double a = 1.0, bb = 0.5 / a;
      double *k = h->stage[0], *x1 = h->stage[1];
      RET_IF(sat_rate(h, S, k));
      RET_IF(xpby(h, S, (1. - bb) * dt, k, x1));
      RET_IF(xpby(h, S, a * dt, k, Sn));
      RET_IF(sat_rate(h, Sn, k));
      RET_IF(xpby(h, x1, bb * dt, k, S));
    } else if (ode == 3) {  // RK3SSPSolver
      double *k = h->stage[0], *y = h->stage[1];
      RET_IF(sat_rate(h, S, k));
      RET_IF(xpby(h, S, dt, k, y));
      RET_IF(sat_rate(h, y, k));
      RET_IF(xpby(h, y, dt, k, y));
      RET_IF(axpby(h, 0.75, S, 0.25, y, y));
      RET_IF(sat_rate(h, y, k));
      RET_IF(xpby(h, y, dt, k, y));
      RET_IF(axpby(h, 1. / 3, S, 2. / 3, y, S));
    } else {  // RK4Solver
      double *k = h->stage[0], *y = h->stage[1], *z = h->stage[2];
      RET_IF(sat_rate(h, S, k));
      RET_IF(xpby(h, S, dt / 2, k, y));
      RET_IF(xpby(h, S, dt / 6, k, z));
      RET_IF(sat_rate(h, y, k));
      RET_IF(xpby(h, S, dt / 2, k, y));
      RET_IF(xpby(h, z, dt / 3, k, z));
      RET_IF(sat_rate(h, y, k));
      RET_IF(xpby(h, S, dt, k, y));
      RET_IF(xpby(h, z, dt / 3, k, z));
      RET_IF(sat_rate(h, y, k));
      RET_IF(xpby(h, z, dt / 6, k, S));
    }
  }
  CK(cudaGetLastError());
  h->coeff_valid = false;
  return 0;
}

// RHS: block 0 = 0, block 1 = -(Q, psi) with the cell-centre rule (P0)  [REF src/pressure_solver.cpp:35-45]
// WellFunctionCoefficient::eval2D/eval3D [REF src/param.hpp:306-345].  Static: built once on the host.
double well_Q(const impes_desc& d, const double* p) {
  const impes_well &in = d.injection, &pr = d.production;
  if (d.dim == 2) {
    if (fabs(p[0] - in.center[0]) <= in.radius && fabs(p[1] - in.center[1]) <= in.radius) return d.inflow;
    if (fabs(p[0] - pr.center[0]) <= pr.radius && fabs(p[1] - pr.center[1]) <= pr.radius) return d.outflow;
    return 0.;
  }
  if (fabs(p[0] - in.center[0]) <= in.radius && fabs(p[2] - in.center[2]) <= in.radius && p[1] >= in.center[1] &&
      p[1] <= in.center[1] + in.height)
    return d.inflow;
  if (fabs(p[0] - pr.center[0]) <= pr.radius && fabs(p[2] - pr.center[2]) <= pr.radius && p[1] >= pr.center[1] &&
      p[1] <= pr.center[1] + pr.height)
    return d.outflow;
  return 0.;
}

// dt_local > 0, both finite, and a sub-step count that fits an int (the serial rule casts global_dt/dt to int, the
// parallel rule loops until t >= global_dt - dt/2: a NaN/inf global_dt would never terminate)
int check_dt(impes_ctx* h, double global_dt, double dt_local) {
  if (!(dt_local > 0) || !std::isfinite(dt_local)) { h->err = "dt_local must be finite and > 0"; return IMPES_ERR_BAD_ARG; }
  if (!std::isfinite(global_dt)) { h->err = "global_dt must be finite"; return IMPES_ERR_BAD_ARG; }
  if (global_dt / dt_local > 2.0e9) { h->err = "global_dt / dt_local exceeds the int range of the sub-step counter"; return IMPES_ERR_BAD_ARG; }
  return 0;
}

int check_handle(impes_ctx* h) {
  if (!h) return IMPES_ERR_BAD_ARG;
  if (cudaSetDevice(h->dev) != cudaSuccess) {
    h->err = "cudaSetDevice failed";
    return IMPES_ERR_CUDA;
  }
  return 0;
}

}  // namespace

// =================================================================================================
// exported C ABI
// =================================================================================================

API int impes_abi_version(void) { return IMPES_ABI_VERSION; }

API void impes_desc_defaults(impes_desc* d) {
  // Param::Param() / DarcySolverParam() / Well()  [REF src/param.cpp:8-13,26-30,35-68]
  memset(d, 0, sizeof *d);
  d->struct_size = (int32_t)sizeof(impes_desc);
  d->dim = 2;
  d->nx = d->ny = d->nz = 100;
  d->sx = d->sy = d->sz = 100.;
  d->order_v = 1; d->order_p = 1; d->order_s = 0;
  d->two_phase = 0;
  d->ode_solver_type = 1;
  d->mu_w = 1.0; d->mu_o = 1.0;  // MU_W, MU_O [REF config.hpp.in:37,43]
  d->injection.radius = d->production.radius = 1.;
  d->injection.height = d->production.height = 1.;
  d->inflow = 1.; d->outflow = -1.; d->saturation_source = 1.;
  d->darcy_maxiter = 1000; d->darcy_rtol = 1e-6; d->darcy_atol = 1e-10; d->darcy_print_level = 0;
  d->schur_prec = IMPES_SCHUR_SGS_LEX;
  d->substep_rule = IMPES_SUBSTEPS_SERIAL;
  d->rank = 0; d->world_size = 1; d->device = 0;
}

API const char* impes_last_error(impes_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

API void impes_destroy(impes_handle h) {
  if (!h) return;
  cudaSetDevice(h->dev);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->minres_exec) cudaGraphExecDestroy(h->minres_exec);
  if (h->comm) h->nccl.CommDestroy(h->comm);
  double* ptrs[] = {h->K[0], h->K[1], h->K[2], h->phi, h->S[0], h->S[1], h->F[0], h->F[1], h->stage[0], h->stage[1], h->stage[2],
                    h->stage[3], h->m[0], h->m[1], h->m[2], h->Md, h->T, h->rhs, h->x, h->v0, h->v1, h->w0, h->w1,
                    h->q, h->u1, h->tmpv, h->partial, h->scal, h->flushbuf, h->pack_blk, h->pack.Yb, h->flux0, h->mg1_ya, h->rlam, h->yup, h->accA, h->partial_u, h->vhq, h->fnq, h->aq};
  for (double* p : ptrs)
    if (p) cudaFree(p);
  for (double* b : h->mg_bufs)
    if (b) cudaFree(b);
  if (h->ticket) cudaFree(h->ticket);
  if (h->st) cudaFree(h->st);
  if (h->guard) cudaFree(h->guard);
  for (cudaEvent_t e : h->user_ev)
    if (e) cudaEventDestroy(e);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

static int create_impl(impes_ctx* h, const double* Kx, const double* Ky, const double* Kz, const double* phi) {
  const impes_desc& d = h->d;
  if (d.dim != 2 && d.dim != 3) { h->err = "dim must be 2 or 3"; return IMPES_ERR_BAD_ARG; }
  if (d.nx < 1 || d.ny < 1 || (d.dim == 3 && d.nz < 1)) { h->err = "grid sizes must be >= 1"; return IMPES_ERR_BAD_ARG; }
  if (!Kx || !Ky || (d.dim == 3 && !Kz)) { h->err = "permeability arrays missing"; return IMPES_ERR_BAD_ARG; }
  if (d.order_v == 0 && d.order_p == 0 && d.order_s == 0) h->ord = 0;
  else if (d.order_v == 1 && d.order_p == 1 && d.order_s == 0) h->ord = 1;
  else if (d.order_v == 1 && d.order_p == 1 && d.order_s == 1) { h->ord = 1; h->os = 1; }
  else if (d.order_v == 0 && d.order_p == 0 && d.order_s == 1) { h->ord = 0; h->os = 1; }
  else {
    h->err = "GPU path implements -ov 0 -op 0 -os 0 (RT0/P0/P0), -ov 0 -op 0 -os 1 (RT0/P0/DG-Q1), -ov 1 -op 1 -os 0 "
             "(RT1/Q1/P0, the reference default) and -ov 1 -op 1 -os 1 (RT1/Q1/DG-Q1)";
    return IMPES_ERR_UNSUPPORTED;
  }
  h->nsl = h->os == 1 ? (d.dim == 3 ? 8 : 4) : 1;
  if (d.world_size < 1 || d.rank < 0 || d.rank >= d.world_size) { h->err = "bad rank/world_size"; return IMPES_ERR_BAD_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    h->err = "no CUDA device available (the IMPES hot path has no CPU fallback)";
    return IMPES_ERR_CUDA;
  }
  h->dev = d.device;
  CK(cudaSetDevice(h->dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->dev));
  h->num_sms = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  if (const char* e = getenv("IMPES_NO_GRAPH")) h->use_graph = !(e[0] == '1');

  Lay& L = h->L;
  L.NX = d.nx;
  L.has_y = d.dim == 3;
  L.NY = d.dim == 3 ? d.ny : 1;
  L.NZ = d.dim == 3 ? d.nz : d.ny;
  const double H[3] = {d.sx / d.nx, d.sy / d.ny, d.dim == 3 ? d.sz / d.nz : 1.0};
  h->vol = H[0] * H[1] * (d.dim == 3 ? H[2] : 1.0);
  if (d.dim == 3) { h->h[0] = H[0]; h->h[1] = H[1]; h->h[2] = H[2]; }
  else { h->h[0] = H[0]; h->h[1] = 1.0; h->h[2] = H[1]; }
  for (int c = 0; c < 3; ++c) h->geo[c] = h->h[c] * h->h[c] / h->vol;
  if (d.world_size > L.NZ) { h->err = "more ranks than cell planes"; return IMPES_ERR_BAD_ARG; }
  L.rank = d.rank;
  L.nranks = d.world_size;
  L.fdx = make_fastdiv((unsigned)L.NX);
  L.fdy = make_fastdiv((unsigned)L.NY);
  int mg_m = 0;  // multigrid mode on several GPUs: slabs are cut in units of 2^mg_m planes
  if (d.schur_prec == IMPES_SCHUR_MG && L.nranks > 1) {
    // Levels 0 .. g-1 stay distributed over the slabs (halo exchange per sweep), level g and below run replicated on
    // every GPU after one gather.  g = the first level small enough to be cheaper replicated than exchanged
    // (<= mg_repl_cells unknowns: a sweep over it costs less than one NCCL hand-shake), at most the largest g whose
    // 2^g-plane units still give every GPU a unit.  (Round 1 distributed as many levels as possible: ~12 more NCCL
    // operations per V-cycle at 8 GPUs and slabs of 16,16,16,8,8,8,8,5 planes for the 85-plane SPE10 shape.)
    long repl = 1L << 19;
    if (const char* e = getenv("IMPES_MG_REPL_CELLS")) repl = std::max(1L, atol(e));
    int gmax = 0;
    while (gmax < 6 && (L.NZ >> (gmax + 1)) >= L.nranks) ++gmax;
    if (gmax < 1) { h->err = "IMPES_SCHUR_MG: needs at least 2 cell planes per GPU"; return IMPES_ERR_BAD_ARG; }
    auto level_cells = [&](int l) {
      long nx = L.NX, ny = L.NY, nz = L.NZ;
      for (int k = 0; k < l; ++k) { nx = (nx + 1) / 2; ny = ny > 1 ? (ny + 1) / 2 : 1; nz = (nz + 1) / 2; }
      return nx * ny * nz;
    };
    mg_m = 1;
    while (mg_m < gmax && level_cells(mg_m) > repl) ++mg_m;
  }
  auto slab_of = [&](int rank, int& lo, int& hi) {
    // balanced slabs of plane units: the first (n % P) ranks get one extra unit (unit = 1 plane
    // outside the multi-GPU multigrid mode)
    const int unit = 1 << mg_m, nu = (L.NZ + unit - 1) / unit;
    const int base = nu / L.nranks, rem = nu % L.nranks;
    const int ulo = rank * base + std::min(rank, rem), uhi = ulo + base + (rank < rem ? 1 : 0);
    lo = ulo * unit;
    hi = std::min(uhi * unit, L.NZ);
  };
  slab_of(d.rank, L.zlo, L.zhi);
  L.nzl = L.zhi - L.zlo;
  {  // all per-rank element offsets must fit 32-bit ints (layout.h)
    const long plane = (long)L.NX * L.NY;
    const long off_y = pad32((long)(L.NX + 1) * L.NY * L.nzl);
    const long off_z = L.has_y ? pad32(off_y + (long)L.NX * (L.NY + 1) * L.nzl) : off_y;
    const long off_p = pad32(off_z + plane * (L.nzl + 2));
    const long ntot = pad32(off_p + plane * (L.nzl + 2));
    if (ntot >= (1L << 31) - 64) {
      h->err = "slab too large: more than 2^31 Darcy dofs per GPU (use more GPUs)";
      return IMPES_ERR_UNSUPPORTED;
    }
    L.plane = (int)plane;
    L.ncl = (int)(plane * L.nzl);
    L.ncv = (int)(plane * (L.nzl + 2));
    L.off_x = 0; L.off_y = (int)off_y; L.off_z = (int)off_z; L.off_p = (int)off_p; L.ntot = (int)ntot;
  }
  L.nblk = (int)((L.ncl + BLK - 1) / BLK);
  h->g_ncell = L.plane * L.NZ;
  h->g_nvx = (long)(L.NX + 1) * L.NY * L.NZ;
  h->g_nvy = L.has_y ? (long)L.NX * (L.NY + 1) * L.NZ : 0;
  h->g_nvz = L.plane * (L.NZ + 1);
  h->g_nv = h->g_nvx + h->g_nvy + h->g_nvz;
  h->g_np = h->g_ncell;
  h->vec_n = L.ntot;
  h->vec_offp = L.off_p;
  if (h->ord == 1) {  // dvec1 layout (rt1.cuh)
    Lay1& Q = h->L1;
    Q.fdx = L.fdx; Q.fdy = L.fdy;
    Q.NX = L.NX; Q.NY = L.NY; Q.NZ = L.NZ; Q.has_y = L.has_y; Q.oy = L.has_y ? 2 : 1; Q.npl = 4 * Q.oy;
    Q.zlo = L.zlo; Q.zhi = L.zhi; Q.nzl = L.nzl; Q.rank = L.rank; Q.nranks = L.nranks;
    Q.plane = L.plane; Q.ncl = L.ncl; Q.nblk = L.nblk;
    Q.gx0 = 2 * L.NX + 1; Q.gx1 = Q.oy * L.NY;
    Q.gy0 = 2 * L.NX; Q.gy1 = 2 * L.NY + 1;
    Q.gz0 = 2 * L.NX; Q.gz1 = Q.oy * L.NY;
    const long nx_ = (long)Q.gx0 * Q.gx1 * 2 * L.nzl;
    const long ny_ = L.has_y ? (long)Q.gy0 * Q.gy1 * 2 * L.nzl : 0;
    const long nz_ = (long)Q.gz0 * Q.gz1 * (2 * L.nzl + 3);
    const long np_ = (long)L.plane * (L.nzl + 2) * Q.npl;
    const long off_y = pad32(nx_), off_z = pad32(off_y + ny_), off_p = pad32(off_z + nz_), ntot = pad32(off_p + np_);
    if (ntot >= (1L << 31) - 64) {
      h->err = "slab too large: more than 2^31 Darcy dofs per GPU (use more GPUs)";
      return IMPES_ERR_UNSUPPORTED;
    }
    Q.off_x = 0; Q.off_y = (int)off_y; Q.off_z = (int)off_z; Q.off_p = (int)off_p; Q.ntot = (int)ntot;
    h->g_nvx = (long)Q.gx0 * Q.gx1 * 2 * L.NZ;
    h->g_nvy = L.has_y ? (long)Q.gy0 * Q.gy1 * 2 * L.NZ : 0;
    h->g_nvz = (long)Q.gz0 * Q.gz1 * (2 * L.NZ + 1);
    h->g_nv = h->g_nvx + h->g_nvy + h->g_nvz;
    h->g_np = h->g_ncell * Q.npl;
    h->vec_n = Q.ntot;
    h->vec_offp = Q.off_p;
    // reference-element constants, same expressions as the oracle's canonical arithmetic
    static const int MCI[3][3] = {{4, 2, -1}, {2, 16, 2}, {-1, 2, 4}};
    const double r3 = sqrt(3.0) / 3.0;
    const double G[2][3] = {{-0.5 - r3, 2.0 * r3, 0.5 - r3}, {-0.5 + r3, -2.0 * r3, 0.5 + r3}};
    const double wt = d.dim == 3 ? 0.25 : 0.5;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) h->rc.CW[i][j] = (MCI[i][j] / 30.0) * wt;
    for (int a = 0; a < 2; ++a)
      for (int i = 0; i < 3; ++i) h->rc.Bv[a][i] = -(G[a][i] * wt);
    h->rc.wflux = wt;
    h->grid_dofs1 = (int)(((long)L.ncl * Q.npl + BLK - 1) / BLK);
  }
  if (h->os == 1 && h->ord == 0) {  // k1q_sat walks the cells through a Lay1: only its cell-index fields are used
    Lay1& Q = h->L1;
    memset(&Q, 0, sizeof Q);
    Q.fdx = L.fdx; Q.fdy = L.fdy;
    Q.NX = L.NX; Q.NY = L.NY; Q.NZ = L.NZ; Q.has_y = L.has_y; Q.oy = L.has_y ? 2 : 1; Q.npl = 1;
    Q.zlo = L.zlo; Q.zhi = L.zhi; Q.nzl = L.nzl; Q.rank = L.rank; Q.nranks = L.nranks;
    Q.plane = L.plane; Q.ncl = L.ncl; Q.nblk = L.nblk;
  }
  if (h->os == 1) {  // quadrature and basis tables of rt1q.cuh
    QTab& t = h->qt;
    memset(&t, 0, sizeof t);
    const int nq = d.dim == 3 ? 4 : 3, nf = d.dim == 3 ? 3 : 2;
    double gx[4], gwt[4], fx[3];
    auto gauss = [](int n, double* x, double* w) {  // Gauss-Legendre on [0,1], ascending
      if (n == 2) { const double a = 0.5 / sqrt(3.0); x[0] = 0.5 - a; x[1] = 0.5 + a; w[0] = w[1] = 0.5; }
      else if (n == 3) { const double a = 0.5 * sqrt(0.6); x[0] = 0.5 - a; x[1] = 0.5; x[2] = 0.5 + a; w[0] = w[2] = 5.0 / 18.0; w[1] = 8.0 / 18.0; }
      else {
        const double r = sqrt(1.2), a = 0.5 * sqrt((3.0 - 2.0 * r) / 7.0), b = 0.5 * sqrt((3.0 + 2.0 * r) / 7.0);
        const double wa = (18.0 + sqrt(30.0)) / 72.0, wb = (18.0 - sqrt(30.0)) / 72.0;
        x[0] = 0.5 - b; x[1] = 0.5 - a; x[2] = 0.5 + a; x[3] = 0.5 + b; w[0] = w[3] = wb; w[1] = w[2] = wa;
      }
    };
    gauss(nq, gx, gwt);
    gauss(nf, fx, t.fw);
    gauss(d.dim == 3 ? 3 : 2, t.x0, t.gw0);   // RT0 mass rule (-ov 0 -os 1)
    const double g1 = 0.5 / sqrt(3.0), on[2] = {0.5 - g1, 0.5 + g1}, cn[3] = {0.0, 0.5, 1.0};
    auto lagr = [](const double* nodes, int nn, int a, double x) {
      double v = 1.0;
      for (int m = 0; m < nn; ++m) if (m != a) v *= (x - nodes[m]) / (nodes[a] - nodes[m]);
      return v;
    };
    for (int q = 0; q < nq; ++q) {
      t.gw[q] = gwt[q];
      for (int i = 0; i < 3; ++i) t.C[q][i] = lagr(cn, 3, i, gx[q]);
      for (int a = 0; a < 2; ++a) t.O[q][a] = lagr(on, 2, a, gx[q]);
    }
    for (int q = 0; q < nf; ++q)
      for (int a = 0; a < 2; ++a) t.OF[q][a] = lagr(on, 2, a, fx[q]);
    for (int q = 0; q < (d.dim == 3 ? 3 : 2); ++q)
      for (int a = 0; a < 2; ++a) t.O0[q][a] = lagr(on, 2, a, t.x0[q]);
    for (int a = 0; a < 2; ++a) {
      for (int i = 0; i < 3; ++i) t.LC[a][i] = lagr(cn, 3, i, on[a]);
      t.E0[a] = lagr(on, 2, a, 0.0);
      t.E1[a] = lagr(on, 2, a, 1.0);
      t.sn[a] = on[a];
    }
    t.dl[0] = 1.0 / (on[0] - on[1]); t.dl[1] = 1.0 / (on[1] - on[0]);
    t.wv = d.dim == 3 ? 0.125 : 0.25;
    for (int qe = 0; qe < 4; ++qe)
      for (int qf = 0; qf < 4; ++qf)
        for (int tf = 0; tf < 2; ++tf)
          for (int te = 0; te < 2; ++te) t.OO[qe][qf][te + 2 * tf] = t.O[qe][te] * t.O[qf][tf];
    for (int qe = 0; qe < 3; ++qe)
      for (int qf = 0; qf < 3; ++qf)
        for (int tf = 0; tf < 2; ++tf)
          for (int te = 0; te < 2; ++te) t.OOF[qe][qf][te + 2 * tf] = t.OF[qe][te] * t.OF[qf][tf];
    if (d.dim == 3) {  // the rolled 3-D quadrature loops read these from constant memory (same values for every handle)
      double wf[3][3];
      for (int qe = 0; qe < 3; ++qe)
        for (int qf = 0; qf < 3; ++qf) wf[qe][qf] = t.fw[qe] * t.fw[qf];
      CK(cudaMemcpyToSymbol(c_OO3, t.OO, sizeof t.OO));
      CK(cudaMemcpyToSymbol(c_OOF3, t.OOF, sizeof t.OOF));
      CK(cudaMemcpyToSymbol(c_WF3, wf, sizeof wf));
      double oosq[4][4][4];
      for (int qe = 0; qe < 4; ++qe)
        for (int qf = 0; qf < 4; ++qf)
          for (int tf = 0; tf < 2; ++tf)
            for (int te = 0; te < 2; ++te)
              oosq[qe][qf][te + 2 * tf] = (t.O[qe][te] * t.O[qe][te]) * (t.O[qf][tf] * t.O[qf][tf]);
      CK(cudaMemcpyToSymbol(c_OO3SQ, oosq, sizeof oosq));
    }
    h->nqp = d.dim == 3 ? 64 : 9;
    h->nup = d.dim == 3 ? 12 : 4;
  }
  h->ph.two_phase = d.two_phase; h->ph.Sor = d.Sor; h->ph.Swc = d.Swc; h->ph.mu_w = d.mu_w; h->ph.mu_o = d.mu_o;
  h->ph.den = 1. - d.Sor - d.Swc; h->ph.rden = 1.0 / h->ph.den; h->ph.rmu_w = 1.0 / d.mu_w; h->ph.rmu_o = 1.0 / d.mu_o;
  h->ph.unit_den = (d.Swc == 0.0 && h->ph.den == 1.0) ? 1 : 0;
  h->ph.unit_mu = (d.mu_w == 1.0 && d.mu_o == 1.0) ? 1 : 0;
  SatParam& sp = h->sp;
  sp.two_phase = d.two_phase; sp.dt = 0; sp.vol = h->vol;
  sp.hx = h->h[0]; sp.hy = h->h[1]; sp.hz = h->h[2];
  sp.well_2d = d.dim == 2;
  sp.inj_c0 = d.injection.center[0]; sp.inj_c1 = d.injection.center[1]; sp.inj_c2 = d.injection.center[2];
  sp.inj_r = d.injection.radius; sp.inj_h = d.injection.height; sp.src = d.saturation_source;
  {  // injection box as inclusive internal index ranges: the per-axis predicates of
     // SaturationSourceCoefficient::eval2D/3D [REF src/param.hpp:382-406] evaluated at the cell centres
    auto range = [](int n, auto pred, int& lo, int& hi) {
      lo = n; hi = -1;
      for (int q = 0; q < n; ++q)
        if (pred(q)) { if (lo == n) lo = q; hi = q; }
      if (hi < 0) { lo = 1; hi = 0; }
    };
    range(L.NX, [&](int q) { return fabs((q + 0.5) * sp.hx - sp.inj_c0) <= sp.inj_r; }, sp.bi0, sp.bi1);
    if (d.dim == 2) {
      sp.bj0 = 0; sp.bj1 = 0;
      range(L.NZ, [&](int q) { return fabs((q + 0.5) * sp.hz - sp.inj_c1) <= sp.inj_r; }, sp.bk0, sp.bk1);
    } else {
      range(L.NY, [&](int q) { const double py = (q + 0.5) * sp.hy; return py >= sp.inj_c1 && py <= sp.inj_c1 + sp.inj_h; },
            sp.bj0, sp.bj1);
      range(L.NZ, [&](int q) { return fabs((q + 0.5) * sp.hz - sp.inj_c2) <= sp.inj_r; }, sp.bk0, sp.bk1);
    }
    // DG-Q1 nodes (1/2 -+ 1/(2 sqrt 3) inside the cell): cells with a node inside, per axis (k1q_sat skips the source
    // evaluation everywhere else)
    const double g1 = 0.5 / sqrt(3.0), sn[2] = {0.5 - g1, 0.5 + g1};
    auto any = [&](auto pred) { return [=](int q) { return pred((q + sn[0])) || pred((q + sn[1])); }; };
    range(L.NX, any([&](double t) { return fabs(t * sp.hx - sp.inj_c0) <= sp.inj_r; }), sp.ni0, sp.ni1);
    if (d.dim == 2) {
      sp.nj0 = 0; sp.nj1 = 0;
      range(L.NZ, any([&](double t) { return fabs(t * sp.hz - sp.inj_c1) <= sp.inj_r; }), sp.nk0, sp.nk1);
    } else {
      range(L.NY, any([&](double t) { const double py = t * sp.hy; return py >= sp.inj_c1 && py <= sp.inj_c1 + sp.inj_h; }), sp.nj0, sp.nj1);
      range(L.NZ, any([&](double t) { return fabs(t * sp.hz - sp.inj_c2) <= sp.inj_r; }), sp.nk0, sp.nk1);
    }
  }

  h->grid_cells = L.nblk;
  h->grid_vec = (int)std::min<long>(((long)h->vec_n + BLK - 1) / BLK, (long)h->num_sms * 8);
  {  // Schur sweep launch shape: the fewest SMs whose barrier is cheap enough
    const long npair = (long)L.NY * L.nzl;
    int csz = 8;
    if (const char* e = getenv("IMPES_SGS_CLUSTER")) csz = std::max(1, std::min(16, atoi(e)));
    bool skew = h->ord == 0;  // the stage-major packed sweep exists for the RT0/P0 stencil only
    if (const char* e = getenv("IMPES_SGS_SKEW")) skew = skew && atoi(e) != 0;
    // one CTA only when the TMA ring kernel applies (<= 992 pencils); otherwise a cluster (hardware barrier)
    int mode = (npair + 31) / 32 * 32 + 32 <= 1024 ? 0 : 1;
    if (h->ord == 1 && npair > 8192) mode = 2;
    if (const char* e = getenv("IMPES_SGS_MODE")) mode = atoi(e);
    const long nstage = (long)L.NX + L.NY + L.nzl - 2;
    const long cap = (long)(mode == 0 ? 1 : csz) * SGS_T * 4;
    if (skew && mode != 2 && npair <= cap && nstage * ((npair + 31) / 32 * 32) * 7 < (1L << 33)) {
      h->sgs_skew = true;
      h->sgs_mode = mode;
      h->grid_sgs = mode == 0 ? 1 : csz;
      const long per_cta = (npair + h->grid_sgs - 1) / h->grid_sgs;
      h->sgs_ppt = per_cta <= SGS_T ? 1 : (per_cta <= 2 * SGS_T ? 2 : 4);
      const long thr = (per_cta + h->sgs_ppt - 1) / h->sgs_ppt;
      h->sgs_threads = (int)std::min<long>(SGS_T, std::max<long>(32, (thr + 31) / 32 * 32));
      if (mode == 1 && csz > 8) {
        CK(cudaFuncSetAttribute(k_sgs_skew<1, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        CK(cudaFuncSetAttribute(k_sgs_skew<1, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        CK(cudaFuncSetAttribute(k_sgs_skew<1, 4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      }
      h->pack.NPs = (int)((npair + 31) / 32 * 32);
      for (int c : {64, 128, 256, 512})
        if (npair <= c) { h->pack.NPs = c; break; }
      h->pack.SS = 7 * h->pack.NPs;
      bool tma = mode == 0 && h->pack.NPs + 32 <= 1024;
      if (const char* e = getenv("IMPES_SGS_TMA")) tma = tma && atoi(e) != 0;
      if (tma) {
        const size_t row = (size_t)h->pack.NPs * 8;
        h->sgs_yp = (L.NY + 1 + 15) / 16 * 16;
        const size_t ybytes = 2 * (row + 2 * (size_t)h->sgs_yp * 8);
        int ns = (int)((200 * 1024 - ybytes) / (7 * row));
        ns = std::min(ns, 8);
        if (const char* e = getenv("IMPES_SGS_NS")) ns = std::min(ns, std::max(4, atoi(e)));
        if (ns >= 4) {
          h->sgs_tma = true;
          h->sgs_ns = ns;
          h->sgs_smem = (size_t)ns * 7 * row + ybytes + (size_t)ns * 8;
          h->sgs_threads = h->pack.NPs + 32;
          h->sgs_ppt = 1;
          const int sm = (int)h->sgs_smem;
          CK(cudaFuncSetAttribute(k_sgs_tma<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
          CK(cudaFuncSetAttribute(k_sgs_tma<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
          CK(cudaFuncSetAttribute(k_sgs_tma<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
          CK(cudaFuncSetAttribute(k_sgs_tma<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
          CK(cudaFuncSetAttribute(k_sgs_tma<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
          CK(cudaFuncSetAttribute(k_sgs_tma<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
          CK(cudaFuncSetAttribute(k_sgs_tma<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
          CK(cudaFuncSetAttribute(k_sgs_tma<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
          CK(cudaFuncSetAttribute(k_sgs_tma<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
          CK(cudaFuncSetAttribute(k_sgs_tma<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        }
      }
      RET_IF(dev_alloc(h, &h->pack_blk, nstage * h->pack.SS));
      RET_IF(dev_alloc(h, &h->pack.Yb, nstage * h->pack.NPs));
      h->pack.Rs = h->pack_blk;
      h->pack.D = h->pack_blk + 1L * h->pack.NPs;
      h->pack.rD = h->pack_blk + 2L * h->pack.NPs;
      h->pack.nTx = h->pack_blk + 3L * h->pack.NPs;
      h->pack.nTz = h->pack_blk + 4L * h->pack.NPs;
      h->pack.Yf = h->pack_blk + 5L * h->pack.NPs;
      h->pack.nTy = h->pack_blk + 6L * h->pack.NPs;
    } else {
      if (mode == 0 && npair > 1024) mode = 1;
      h->sgs_mode = mode;
      if (mode == 0) {
        h->grid_sgs = 1;
        h->sgs_threads = (int)std::min<long>(1024, (npair + 31) / 32 * 32);
      } else if (mode == 1) {
        if (csz > 8) {
          CK(cudaFuncSetAttribute(k_sgs_wavefront<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
          CK(cudaFuncSetAttribute(k1_sgs_wavefront<1, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
          CK(cudaFuncSetAttribute(k1_sgs_wavefront<1, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
        h->grid_sgs = csz;
        const long per = (npair + csz - 1) / csz;
        h->sgs_threads = (int)std::min<long>(1024, std::max<long>(32, (per + 31) / 32 * 32));
      } else {
        int per_sm = 0;
        h->sgs_threads = 256;
        if (h->ord == 1) {
          if (L.has_y) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_sgs_wavefront<2, true>, 256, 0));
          else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_sgs_wavefront<2, false>, 256, 0));
        } else
          CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sgs_wavefront<2>, 256, 0));
        const long want = (npair + 255) / 256;
        h->grid_sgs = (int)std::max<long>(1, std::min<long>(want, (long)per_sm * h->num_sms));
      }
    }
  }

  // device arrays
  const double* Kin[3] = {Kx, d.dim == 3 ? Ky : nullptr, d.dim == 3 ? Kz : Ky};
  for (int c = 0; c < 3; ++c) {
    RET_IF(dev_alloc(h, &h->K[c], L.ncv));
    RET_IF(dev_alloc(h, &h->m[c], L.ncv));
    if (Kin[c]) RET_IF(upload_cells(h, Kin[c], h->K[c]));
  }
  RET_IF(dev_alloc(h, &h->phi, L.ncv));
  {
    std::vector<double> ph(h->g_ncell, 1.0);
    if (phi) memcpy(ph.data(), phi, sizeof(double) * h->g_ncell);
    RET_IF(upload_cells(h, ph.data(), h->phi));
    CK(cudaStreamSynchronize(h->stream));
  }
  RET_IF(dev_alloc(h, &h->S[0], (long)L.ncv * h->nsl));
  RET_IF(dev_alloc(h, &h->S[1], (long)L.ncv * h->nsl));
  RET_IF(dev_alloc(h, &h->F[0], (long)L.ncv * h->nsl));
  RET_IF(dev_alloc(h, &h->F[1], (long)L.ncv * h->nsl));
  if (h->os == 1) {
    if (h->ord == 1) {
      RET_IF(dev_alloc(h, &h->rlam, (long)L.ncv * h->nqp));
      RET_IF(dev_alloc(h, &h->yup, (long)L.ncv * h->nup));
      RET_IF(dev_alloc(h, &h->accA, (long)L.nblk * BLK));
    } else {
      RET_IF(dev_alloc(h, &h->aq, 9L * L.ncv));   // RT0: entries of the 2x2 element blocks, [3 dir + {00,01,11}][ncv]
    }
    if (L.has_y) {
      RET_IF(dev_alloc(h, &h->vhq, (long)L.ncv * SatQ<true>::ND * SatQ<true>::NVH));
      RET_IF(dev_alloc(h, &h->fnq, (long)L.ncv * SatQ<true>::ND * SatQ<true>::NFP));
    } else {
      RET_IF(dev_alloc(h, &h->vhq, (long)L.ncv * SatQ<false>::ND * SatQ<false>::NVH));
      RET_IF(dev_alloc(h, &h->fnq, (long)L.ncv * SatQ<false>::ND * SatQ<false>::NFP));
    }
    CK(cudaFuncSetAttribute(k1q_apply_a<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * K1QS<true>::SMEM_DOUBLES)));
    CK(cudaFuncSetAttribute(k1q_apply_a<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * K1QS<true>::SMEM_DOUBLES)));
    h->lazy_u = h->ord == 1 && !getenv("IMPES_NO_LAZY_U");   // K1 of the RT1 / DG-Q1 path applies the lazy scaling of u
  }
  double** vecs[] = {&h->Md, &h->T, &h->rhs, &h->x, &h->v0, &h->v1, &h->w0, &h->w1, &h->q, &h->u1, &h->tmpv};
  for (double** v : vecs) RET_IF(dev_alloc(h, v, h->vec_n));
  if (h->ord == 1) RET_IF(dev_alloc(h, &h->flux0, L.ntot));
  {  // Md = 1 everywhere first (padding and ghost entries must stay harmless under division)
    k_fill<<<h->grid_vec, BLK, 0, h->stream>>>(h->Md, h->vec_n, 1.0);
    h->launches++;
  }
  if (d.schur_prec == IMPES_SCHUR_MG) {
    if (d.mg_smooth > 0) h->mg_nu = std::min(8, (d.mg_smooth + 1) / 2 * 2);  // even sweep counts (see mg_vcycle), at most 8
    const bool multi = L.nranks > 1;
    h->mg_g = multi ? mg_m : (1 << 30);
    if (h->ord == 1) {  // second iterate buffer of the Q1 level, p-part layout with ghost planes
      RET_IF(dev_alloc(h, &h->mg1_ya, (long)L.plane * (L.nzl + 2) * h->L1.npl));
    }
    int nx = L.NX, ny = L.NY, nzg = L.NZ;
    for (int l = 0; l < 24; ++l) {
      MgLevel v;
      memset(&v, 0, sizeof v);
      v.nx = nx; v.ny = ny; v.nzg = nzg;
      if (multi && l < h->mg_g) {  // slab part: planes [zlo >> l, ceil(zhi / 2^l))
        v.zoff = L.zlo >> l;
        v.nz = ((L.zhi + (1 << l) - 1) >> l) - v.zoff;
      } else {
        v.zoff = 0; v.nz = nzg;
      }
      if (multi && l == h->mg_g) {
        h->mg_goff.resize(L.nranks); h->mg_gcnt.resize(L.nranks);
        for (int r = 0; r < L.nranks; ++r) {
          int lo, hi;
          slab_of(r, lo, hi);
          h->mg_goff[r] = lo >> l;
          h->mg_gcnt[r] = ((hi + (1 << l) - 1) >> l) - h->mg_goff[r];
        }
      }
      const long pl = (long)nx * ny, nc = pl * v.nz;
      double** arrs[] = {&v.Tx, &v.Ty, &v.Tz, &v.D};
      const long sizes[] = {(long)(nx + 1) * ny * v.nz, (long)nx * (ny + 1) * v.nz, pl * (v.nz + 1), nc};
      for (int a = 0; a < 4; ++a) { RET_IF(dev_alloc(h, arrs[a], sizes[a])); h->mg_bufs.push_back(*arrs[a]); }
      // iterate / rhs buffers carry one ghost plane below and above (used by slab-distributed levels)
      double* b = nullptr;
      RET_IF(dev_alloc(h, &b, nc + 2 * pl)); h->mg_bufs.push_back(b); v.ya = b + pl;
      if (l > 0 || h->ord == 1) {
        RET_IF(dev_alloc(h, &b, nc + 2 * pl)); h->mg_bufs.push_back(b); v.r = b + pl;
        RET_IF(dev_alloc(h, &b, nc + 2 * pl)); h->mg_bufs.push_back(b); v.yb = b + pl;
      }
      h->mg.push_back(v);
      if (nx <= 4 && ny <= 4 && nzg <= 4) break;
      nx = (nx + 1) / 2; ny = ny > 1 ? (ny + 1) / 2 : 1; nzg = (nzg + 1) / 2;
    }
  }
  RET_IF(dev_alloc(h, &h->partial, std::max(2 * L.nblk, h->grid_dofs1) + 8));
  RET_IF(dev_alloc(h, &h->scal, 8));
  if (h->ord == 1 && d.schur_prec == IMPES_SCHUR_MG && !getenv("IMPES_NO_FOLD_DOT")) {
    h->fold_dot = true;
    h->n_u_chunks_all = (int)(((long)h->vec_n + DOT_CHUNK - 1) / DOT_CHUNK);
    h->n_u_chunks = (int)(((long)h->vec_offp + DOT_CHUNK - 1) / DOT_CHUNK);
    RET_IF(dev_alloc(h, &h->partial_u, h->n_u_chunks_all + 8));
  }
  CK(cudaMalloc((void**)&h->ticket, sizeof(unsigned) * 4));
  CK(cudaMemsetAsync(h->ticket, 0, sizeof(unsigned) * 4, h->stream));
  CK(cudaMalloc((void**)&h->st, sizeof(MinresState) * 4));  // [0,1] solver, [2] timing state, [3] sink
  CK(cudaMemsetAsync(h->st, 0, sizeof(MinresState) * 4, h->stream));
  {
    MinresState bs;
    memset(&bs, 0, sizeof bs);
    bs.beta = 1.0; bs.eta = 1.0; bs.gamma0 = bs.gamma1 = 1.0; bs.m0 = bs.m1 = 1.0; bs.mu = 1.0;
    const double sc[4] = {0.0, 0.0, 0.5, 1.0};
    CK(cudaMemcpyAsync(h->st + 2, &bs, sizeof bs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->scal, sc, sizeof sc, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  CK(cudaMalloc((void**)&h->guard, sizeof(unsigned long long) * 2));
  {  // rhs (global, native) built on the host, uploaded slab-wise
    std::vector<double> g(h->g_np);
    if (h->ord == 1) {
      // Q1: DomainLFIntegrator's 2-point Gauss rule coincides with the L2 nodes => g_m = -Q(x_m) |cell| / 2^dim
      // (SURVEY A.8); well membership is tested at the Gauss points [REF src/param.hpp:285-292]
      const double g1 = 0.5 / sqrt(3.0);
      const double node[2] = {0.5 - g1, 0.5 + g1};
      const double wv = h->vol * (d.dim == 3 ? 0.125 : 0.25);
      const int npl = h->L1.npl, nb2 = d.dim == 3 ? 2 : 1;
      for (int kz = 0; kz < (d.dim == 3 ? d.nz : 1); ++kz)
        for (int jy = 0; jy < d.ny; ++jy)
          for (int ix = 0; ix < d.nx; ++ix)
            for (int b2 = 0; b2 < nb2; ++b2)
              for (int b1 = 0; b1 < 2; ++b1)
                for (int b0 = 0; b0 < 2; ++b0) {
                  const double p[3] = {(ix + node[b0]) * H[0], (jy + node[b1]) * H[1], d.dim == 3 ? (kz + node[b2]) * H[2] : 0.0};
                  g[(ix + (long)d.nx * (jy + (long)d.ny * kz)) * npl + b0 + 2 * (b1 + 2 * b2)] = -(well_Q(d, p) * wv);
                }
    } else
    for (int kz = 0; kz < (d.dim == 3 ? d.nz : 1); ++kz)
      for (int jy = 0; jy < d.ny; ++jy)
        for (int ix = 0; ix < d.nx; ++ix) {
          const double p[3] = {(ix + 0.5) * H[0], (jy + 0.5) * H[1], (kz + 0.5) * H[2]};
          const double s = h->vol * well_Q(d, p);
          g[ix + (long)d.nx * (jy + (long)d.ny * kz)] = -s;
        }
    RET_IF(upload_dvec(h, nullptr, g.data(), h->rhs));
    CK(cudaStreamSynchronize(h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return 0;
}

API int impes_create(const impes_desc* d, const double* Kx, const double* Ky, const double* Kz, const double* phi,
                     impes_handle* out) {
  if (!out) return IMPES_ERR_BAD_ARG;
  *out = nullptr;
  if (!d || d->struct_size != (int32_t)sizeof(impes_desc)) {
    g_create_error = "impes_create: null descriptor or struct_size mismatch";
    return IMPES_ERR_BAD_ARG;
  }
  impes_ctx* h = new impes_ctx();
  h->d = *d;
  const int rc = create_impl(h, Kx, Ky, Kz, phi);
  if (rc) {
    g_create_error = h->err;
    impes_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

API int impes_get_dims(impes_handle h, impes_dims* out) {
  if (!h || !out) return IMPES_ERR_BAD_ARG;
  memset(out, 0, sizeof *out);
  out->n_cells = h->g_ncell;
  out->n_v = h->g_nv; out->n_p = h->g_np; out->n_s = h->g_ncell * h->nsl;
  out->n_v_comp[0] = h->g_nvx;
  if (h->d.dim == 3) { out->n_v_comp[1] = h->g_nvy; out->n_v_comp[2] = h->g_nvz; }
  else { out->n_v_comp[1] = h->g_nvz; out->n_v_comp[2] = 0; }
  out->slab_lo = h->L.zlo; out->slab_hi = h->L.zhi;
  return 0;
}

API int impes_comm_unique_id(void* id128) {
  NcclShim sh;
  if (!sh.load()) return IMPES_ERR_NCCL;
  ncclUniqueId id;
  if (sh.GetUniqueId(&id) != 0) return IMPES_ERR_NCCL;
  memcpy(id128, &id, 128);
  return 0;
}

API int impes_comm_init(impes_handle h, const void* id128) {
  RET_IF(check_handle(h));
  if (h->L.nranks == 1) return 0;
  if (!h->nccl.load()) { h->err = "cannot dlopen libnccl.so.2"; return IMPES_ERR_NCCL; }
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NK(h->nccl.CommInitRank(&h->comm, h->L.nranks, id, h->L.rank));
  return 0;
}

API int impes_set_field(impes_handle h, int which, const double* host) {
  RET_IF(check_handle(h));
  if (!host) return IMPES_ERR_BAD_ARG;
  switch (which) {
    case IMPES_FIELD_SATURATION: RET_IF(upload_cells(h, host, h->S[h->scur], h->nsl)); h->coeff_valid = false; break;
    case IMPES_FIELD_VELOCITY: RET_IF(upload_dvec(h, host, nullptr, h->x)); break;
    case IMPES_FIELD_PRESSURE: RET_IF(upload_dvec(h, nullptr, host, h->x)); break;
    case IMPES_FIELD_DARCY_X: RET_IF(upload_dvec(h, host, host + h->g_nv, h->x)); break;
    default: h->err = "unknown field"; return IMPES_ERR_BAD_ARG;
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

API int impes_get_field(impes_handle h, int which, double* host) {
  RET_IF(check_handle(h));
  if (!host) return IMPES_ERR_BAD_ARG;
  switch (which) {
    case IMPES_FIELD_SATURATION: RET_IF(download_cells_owned(h, h->S[h->scur], host, h->nsl)); break;
    case IMPES_FIELD_VELOCITY: RET_IF(download_dvec_owned(h, h->x, host, nullptr)); break;
    case IMPES_FIELD_PRESSURE: RET_IF(download_dvec_owned(h, h->x, nullptr, host)); break;
    case IMPES_FIELD_DARCY_X: RET_IF(download_dvec_owned(h, h->x, host, host + h->g_nv)); break;
    default: h->err = "unknown field"; return IMPES_ERR_BAD_ARG;
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

API int impes_update_coefficients(impes_handle h) {
  RET_IF(check_handle(h));
  RET_IF(update_coeff(h));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

API int impes_pressure_solve(impes_handle h, const double* S_host, double* x_host, impes_solve_stats* stats) {
  RET_IF(check_handle(h));
  impes_solve_stats loc;
  if (!stats) stats = &loc;
  memset(stats, 0, sizeof *stats);
  auto t0 = std::chrono::steady_clock::now();
  if (S_host) RET_IF(upload_cells(h, S_host, h->S[h->scur], h->nsl));
  const int rc = pressure_solve_device(h, stats);
  if (rc && rc != IMPES_ERR_NOT_CONVERGED) return rc;
  if (x_host) {
    RET_IF(download_dvec_owned(h, h->x, x_host, x_host + h->g_nv));
    CK(cudaStreamSynchronize(h->stream));
  }
  stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

API int impes_saturation_advance(impes_handle h, const double* V_host, double* S_inout, int global_ti, double global_dt,
                                 double dt_local, int* n_substeps_out) {
  (void)global_ti;  // only names output files in the reference [REF src/saturation_solver.cpp:93-108]
  RET_IF(check_handle(h));
  RET_IF(check_dt(h, global_dt, dt_local));
  if (V_host) RET_IF(upload_dvec(h, V_host, nullptr, h->x));
  if (S_inout) RET_IF(upload_cells(h, S_inout, h->S[h->scur], h->nsl));
  RET_IF(saturation_device(h, global_dt, dt_local, n_substeps_out));
  if (S_inout) RET_IF(download_cells_owned(h, h->S[h->scur], S_inout, h->nsl));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// ---- slab-local variants: every rank passes only what it owns (ParGridFunction-style local data) ----------
API int impes_get_local_dims(impes_handle h, impes_local_dims* out) {
  if (!h || !out) return IMPES_ERR_BAD_ARG;
  memset(out, 0, sizeof *out);
  const LocalMap m = local_map(h);
  out->n_s = m.ns; out->n_p = m.np; out->n_v = m.nx + m.ny + m.nz;
  out->s_start = m.gs; out->p_start = m.gp;
  if (h->d.dim == 3) {
    out->n_v_comp[0] = m.nx; out->n_v_comp[1] = m.ny; out->n_v_comp[2] = m.nz;
    out->v_comp_start[0] = m.gx; out->v_comp_start[1] = m.gy; out->v_comp_start[2] = m.gz;
  } else {
    out->n_v_comp[0] = m.nx; out->n_v_comp[1] = m.nz;
    out->v_comp_start[0] = m.gx; out->v_comp_start[1] = m.gz;
  }
  return 0;
}

API int impes_pressure_solve_local(impes_handle h, const double* S_local, double* x_local, impes_solve_stats* stats) {
  RET_IF(check_handle(h));
  impes_solve_stats loc;
  if (!stats) stats = &loc;
  memset(stats, 0, sizeof *stats);
  auto t0 = std::chrono::steady_clock::now();
  if (S_local) RET_IF(copy_cells_local(h, const_cast<double*>(S_local), h->S[h->scur], cudaMemcpyHostToDevice));
  const int rc = pressure_solve_device(h, stats);
  if (rc && rc != IMPES_ERR_NOT_CONVERGED) return rc;
  if (x_local) {
    const LocalMap m = local_map(h);
    RET_IF(copy_dvec_local(h, x_local, x_local + m.nx + m.ny + m.nz, h->x, cudaMemcpyDeviceToHost));
    CK(cudaStreamSynchronize(h->stream));
  }
  stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

API int impes_saturation_advance_local(impes_handle h, const double* V_local, double* S_local, int global_ti,
                                       double global_dt, double dt_local, int* n_substeps_out) {
  (void)global_ti;
  RET_IF(check_handle(h));
  RET_IF(check_dt(h, global_dt, dt_local));
  if (V_local) {  // owned dofs from the host, ghost planes from the neighbours (V.Distribute [REF src/run_parallel.cpp:186])
    RET_IF(copy_dvec_local(h, const_cast<double*>(V_local), nullptr, h->x, cudaMemcpyHostToDevice));
    RET_IF(h->ord == 1 ? halo_darcy1(h, h->x) : halo_darcy(h, h->x));
  }
  if (S_local) RET_IF(copy_cells_local(h, S_local, h->S[h->scur], cudaMemcpyHostToDevice));
  RET_IF(saturation_device(h, global_dt, dt_local, n_substeps_out));
  if (S_local) RET_IF(copy_cells_local(h, S_local, h->S[h->scur], cudaMemcpyDeviceToHost));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

API int impes_set_field_local(impes_handle h, int which, const double* host) {
  RET_IF(check_handle(h));
  if (!host) return IMPES_ERR_BAD_ARG;
  double* hp = const_cast<double*>(host);
  const LocalMap m = local_map(h);
  switch (which) {
    case IMPES_FIELD_SATURATION: RET_IF(copy_cells_local(h, hp, h->S[h->scur], cudaMemcpyHostToDevice)); h->coeff_valid = false; break;
    case IMPES_FIELD_VELOCITY: RET_IF(copy_dvec_local(h, hp, nullptr, h->x, cudaMemcpyHostToDevice)); break;
    case IMPES_FIELD_PRESSURE: RET_IF(copy_dvec_local(h, nullptr, hp, h->x, cudaMemcpyHostToDevice)); break;
    case IMPES_FIELD_DARCY_X: RET_IF(copy_dvec_local(h, hp, hp + m.nx + m.ny + m.nz, h->x, cudaMemcpyHostToDevice)); break;
    default: h->err = "unknown field"; return IMPES_ERR_BAD_ARG;
  }
  if (which != IMPES_FIELD_SATURATION) RET_IF(h->ord == 1 ? halo_darcy1(h, h->x) : halo_darcy(h, h->x));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

API int impes_get_field_local(impes_handle h, int which, double* host) {
  RET_IF(check_handle(h));
  if (!host) return IMPES_ERR_BAD_ARG;
  const LocalMap m = local_map(h);
  switch (which) {
    case IMPES_FIELD_SATURATION: RET_IF(copy_cells_local(h, host, h->S[h->scur], cudaMemcpyDeviceToHost)); break;
    case IMPES_FIELD_VELOCITY: RET_IF(copy_dvec_local(h, host, nullptr, h->x, cudaMemcpyDeviceToHost)); break;
    case IMPES_FIELD_PRESSURE: RET_IF(copy_dvec_local(h, nullptr, host, h->x, cudaMemcpyDeviceToHost)); break;
    case IMPES_FIELD_DARCY_X: RET_IF(copy_dvec_local(h, host, host + m.nx + m.ny + m.nz, h->x, cudaMemcpyDeviceToHost)); break;
    default: h->err = "unknown field"; return IMPES_ERR_BAD_ARG;
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

API int impes_guards(impes_handle h, int64_t* n_nonfinite, double* s_max) {
  RET_IF(check_handle(h));
  CK(cudaMemsetAsync(h->guard, 0, sizeof(unsigned long long) * 2, h->stream));
  {
    Scope sc(h, KI_MISC, 1);
    if (h->os == 1)
      k_guards_flat<<<h->grid_vec, BLK, 0, h->stream>>>((long)h->L.ncl * h->nsl, h->S[h->scur] + (long)h->L.plane * h->nsl,
                                                        h->guard, h->guard + 1);
    else
      k_guards<<<h->grid_cells, BLK, 0, h->stream>>>(h->L, h->S[h->scur], h->guard, h->guard + 1);
  }
  unsigned long long hg[2];
  CK(cudaMemcpyAsync(hg, h->guard, sizeof hg, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  long long bad = (long long)hg[0];
  unsigned long long b = hg[1];
  double mx;
  if (b == 0) mx = -INFINITY;
  else {
    b = (b & 0x8000000000000000ull) ? (b & 0x7fffffffffffffffull) : ~b;
    memcpy(&mx, &b, 8);
  }
  if (h->L.nranks > 1) {  // MPI_Allreduce(SUM) / (MAX) of the reference [REF src/run_parallel.cpp:220-230]
    double* tmp = h->scal + 4;
    double hv[2] = {(double)bad, mx};
    CK(cudaMemcpyAsync(tmp, hv, sizeof hv, cudaMemcpyHostToDevice, h->stream));
    NK(h->nccl.AllReduce(tmp, tmp, 1, ncclFloat64, ncclSum, h->comm, h->stream));
    NK(h->nccl.AllReduce(tmp + 1, tmp + 1, 1, ncclFloat64, ncclMax, h->comm, h->stream));
    CK(cudaMemcpyAsync(hv, tmp, sizeof hv, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    bad = (long long)hv[0];
    mx = hv[1];
  }
  if (n_nonfinite) *n_nonfinite = bad;
  if (s_max) *s_max = mx;
  return 0;
}

API int impes_darcy_apply(impes_handle h, const double* x_host, double* y_host) {
  RET_IF(check_handle(h));
  if (!x_host || !y_host) return IMPES_ERR_BAD_ARG;
  if (!h->coeff_valid) RET_IF(update_coeff(h));
  RET_IF(upload_dvec(h, x_host, x_host + h->g_nv, h->tmpv));
  CK(cudaMemsetAsync(h->q, 0, sizeof(double) * h->vec_n, h->stream));
  RET_IF(darcy_apply(h, h->tmpv, h->q, false, nullptr));
  RET_IF(download_dvec_owned(h, h->q, y_host, y_host + h->g_nv));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

API int impes_prec_apply(impes_handle h, const double* r_host, double* y_host) {
  RET_IF(check_handle(h));
  if (!r_host || !y_host) return IMPES_ERR_BAD_ARG;
  if (!h->coeff_valid) RET_IF(update_coeff(h));
  RET_IF(upload_dvec(h, r_host, r_host + h->g_nv, h->tmpv));
  CK(cudaMemsetAsync(h->q, 0, sizeof(double) * h->vec_n, h->stream));
  {
    Scope sc(h, KI_JACOBI_U, 1);
    k_jacobi_u<<<h->grid_vec, BLK, 0, h->stream>>>(h->vec_offp, h->tmpv, h->Md, h->q);
  }
  RET_IF(schur_prec(h, h->tmpv, h->q, nullptr));
  RET_IF(download_dvec_owned(h, h->q, y_host, y_host + h->g_nv));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

API int impes_get_rhs(impes_handle h, double* rhs_host) {
  RET_IF(check_handle(h));
  if (!rhs_host) return IMPES_ERR_BAD_ARG;
  RET_IF(download_dvec_owned(h, h->rhs, rhs_host, rhs_host + h->g_nv));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

API int impes_time_kernel(impes_handle h, int kernel, int reps, int flush_l2, double* ms_avg, double* ms_min) {
  RET_IF(check_handle(h));
  if (reps < 1) return IMPES_ERR_BAD_ARG;
  if (!h->coeff_valid) RET_IF(update_coeff(h));
  if (flush_l2 && !h->flushbuf) {
    h->flush_n = (long)(256u << 20) / 8;  // 256 MiB > 126 MB L2
    CK(cudaMalloc((void**)&h->flushbuf, sizeof(double) * h->flush_n));
  }
  double sum = 0, mn = 1e300;
  struct Guard {  // events destroyed and the profiling flag restored on every return path
    impes_ctx* h; bool prof; cudaEvent_t a = nullptr, b = nullptr;
    ~Guard() { h->profiling = prof; if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  } guard{h, h->profiling};
  CK(cudaEventCreate(&guard.a));
  CK(cudaEventCreate(&guard.b));
  cudaEvent_t e0 = guard.a, e1 = guard.b;
  h->profiling = false;
  if (kernel == IMPES_KERNEL_SAT_SUBSTEP)  // f(S) is carried between sub-steps
    RET_IF(launch_frac(h, h->S[h->scur], h->F[h->scur]));
  if (kernel == IMPES_KERNEL_SAT_SUBSTEP && h->os == 1) RET_IF(launch_sat_pre(h));
  for (int r = -2; r < reps; ++r) {  // 2 untimed warm-ups
    if (flush_l2) k_fill<<<h->num_sms * 4, 512, 0, h->stream>>>(h->flushbuf, h->flush_n, (double)r);
    CK(cudaEventRecord(e0, h->stream));
    switch (kernel) {
      case IMPES_KERNEL_DARCY_APPLY: RET_IF(darcy_apply(h, h->u1, h->q, true, nullptr)); break;
      case IMPES_KERNEL_SAT_SUBSTEP:
        h->sp.dt = 0.0;  // dt = 0: the state is not disturbed by repeated timing launches
        RET_IF(launch_sat(h, false, h->S[h->scur], h->S[h->scur ^ 1], h->F[h->scur], h->F[h->scur ^ 1]));
        break;
      case IMPES_KERNEL_MINRES_UPDATE: {
        { Scope sc(h, KI_MINRES_B, 1);
        if (h->fold_dot)
          k_minres_b_dot<<<h->n_u_chunks_all, MINRES_B_T, 0, h->stream>>>(h->vec_n, h->vec_offp, h->q, h->v0, h->v1, h->Md, h->q, h->scal + 2,
                                                                   h->st + 2, 0L, 0L, 0L, 0L, h->partial_u);
        else
        k_minres_b<<<h->grid_vec, BLK, 0, h->stream>>>(h->vec_n, h->vec_offp, h->q, h->v0, h->v1, h->Md, h->q, h->scal + 2,
                                                       h->st + 2); }
        Scope sc(h, KI_MINRES_C, 1);
        if (h->lazy_u)
          k_minres_c<true><<<h->grid_vec, BLK, 0, h->stream>>>(h->vec_n, h->w0, h->w1, h->u1, h->tmpv, h->v0, h->q, h->scal + 2,
                                                               h->scal + 3, h->st + 2, h->st + 3, 1 << 30);
        else
          k_minres_c<false><<<h->grid_vec, BLK, 0, h->stream>>>(h->vec_n, h->w0, h->w1, h->u1, h->tmpv, h->v0, h->q, h->scal + 2,
                                                                h->scal + 3, h->st + 2, h->st + 3, 1 << 30);
        break;
      }
      case IMPES_KERNEL_PREC_APPLY: {
        {
          Scope sc(h, KI_JACOBI_U, 1);
          k_jacobi_u<<<h->grid_vec, BLK, 0, h->stream>>>(h->vec_offp, h->v1, h->Md, h->tmpv);
        }
        RET_IF(schur_prec(h, h->v1, h->tmpv, nullptr));
        break;
      }
      case IMPES_KERNEL_COEFF: RET_IF(update_coeff(h)); break;
      default: h->err = "unknown kernel id"; return IMPES_ERR_BAD_ARG;
    }
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 0) { sum += ms; mn = std::min<double>(mn, ms); }
  }
  CK(cudaGetLastError());
  if (ms_avg) *ms_avg = sum / reps;
  if (ms_min) *ms_min = mn;
  return 0;
}

API int impes_event_record(impes_handle h, int slot) {
  RET_IF(check_handle(h));
  if (slot < 0 || slot >= 8) return IMPES_ERR_BAD_ARG;
  if (!h->user_ev[slot]) CK(cudaEventCreate(&h->user_ev[slot]));
  CK(cudaEventRecord(h->user_ev[slot], h->stream));
  return 0;
}

API int impes_event_elapsed_ms(impes_handle h, int a, int b, double* ms) {
  RET_IF(check_handle(h));
  if (a < 0 || a >= 8 || b < 0 || b >= 8 || !h->user_ev[a] || !h->user_ev[b] || !ms) return IMPES_ERR_BAD_ARG;
  CK(cudaEventSynchronize(h->user_ev[b]));
  float f = 0;
  CK(cudaEventElapsedTime(&f, h->user_ev[a], h->user_ev[b]));
  *ms = f;
  return 0;
}

API int64_t impes_launch_count(impes_handle h) { return h ? h->launches : 0; }

API int impes_describe(impes_handle h, char* buf, int buflen) {
  if (!h) return 0;
  char b[512];
  std::string s;
  snprintf(b, sizeof b, "spaces RT%d/L2_%d/DG_%d, %d-D %dx%dx%d, slab %d of %d (planes %d..%d)", h->ord, h->ord, h->os, h->d.dim,
           h->L.NX, h->L.NY, h->L.NZ, h->L.rank, h->L.nranks, h->L.zlo, h->L.zhi);
  s = b;
  if (h->d.schur_prec == IMPES_SCHUR_MG) {
    snprintf(b, sizeof b, "; Schur: multigrid V(%d,%d), %d cell level(s)%s, replicated from level %d", h->mg_nu, h->mg_nu,
             (int)h->mg.size(), h->ord == 1 ? " + L2-Q1 level" : "", h->mg_g > 100 ? -1 : h->mg_g);
    s += b;
  } else if (h->d.schur_prec == IMPES_SCHUR_SGS_LEX) {
    const char* v = h->ord == 1 ? (h->sgs_mode == 2 ? "k1_sgs_wavefront<cooperative grid>" : h->sgs_mode == 1 ? "k1_sgs_wavefront<cluster>" : "k1_sgs_wavefront<one CTA>")
                    : h->sgs_skew ? (h->sgs_tma ? "k_sgs_tma (one-CTA TMA ring)" : h->sgs_mode == 1 ? "k_sgs_skew<cluster>" : "k_sgs_skew<one CTA>")
                    : (h->sgs_mode == 2 ? "k_sgs_wavefront<cooperative grid>" : h->sgs_mode == 1 ? "k_sgs_wavefront<cluster>" : "k_sgs_wavefront<one CTA>");
    snprintf(b, sizeof b, "; Schur: lexicographic symmetric Gauss-Seidel, %s, %d CTA(s) x %d threads", v, h->grid_sgs, h->sgs_threads);
    s += b;
  } else {
    s += "; Schur: Jacobi";
  }
  s += h->os == 1 ? (h->ord == 1 ? "; K1 k1q_apply_a/_b; K6 k1q_sat" : "; K1 k_darcy_apply<VAR>; K6 k1q_sat (records from RT0 fluxes)")
                  : (h->ord == 1 ? "; K1 k1_darcy_apply; K6 k_sat_stream" : "; K1 k_darcy_apply; K6 k_sat_stream");
  s += h->use_graph ? "; MINRES iterations in CUDA graphs" : "; no CUDA graphs";
  if (buf && buflen > 0) {
    strncpy(buf, s.c_str(), (size_t)buflen - 1);
    buf[buflen - 1] = 0;
  }
  return (int)s.size();
}

API int impes_profile_enable(impes_handle h, int on) {
  if (!h) return IMPES_ERR_BAD_ARG;
  h->profiling = on != 0;
  memset(h->prof_ms, 0, sizeof h->prof_ms);
  memset(h->prof_n, 0, sizeof h->prof_n);
  memset(h->ki_ms, 0, sizeof h->ki_ms);
  memset(h->ki_n, 0, sizeof h->ki_n);
  return 0;
}

API int impes_profile_read(impes_handle h, double* ms_by_kernel, int64_t* launches_by_kernel) {
  if (!h) return IMPES_ERR_BAD_ARG;
  for (int i = 0; i < KC_N; ++i) {
    if (ms_by_kernel) ms_by_kernel[i] = h->prof_ms[i];
    if (launches_by_kernel) launches_by_kernel[i] = h->prof_n[i];
  }
  return 0;
}

API int impes_profile_read_kernels(impes_handle h, int cap, const char** names, double* ms, int64_t* launches) {
  if (!h) return -1;
  for (int i = 0; i < KI_N && i < cap; ++i) {
    if (names) names[i] = KI_NAME[i];
    if (ms) ms[i] = h->ki_ms[i];
    if (launches) launches[i] = h->ki_n[i];
  }
  return KI_N;
}
