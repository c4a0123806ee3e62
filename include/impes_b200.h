/*
 * impes_b200.h -- C ABI of the B200-native IMPES hot path.
 *
 * This is the drop-in boundary for the two per-time-step solver calls of
 * martemyev/incompressible-flow:
 *
 *   PressureSolver   [REF src/pressure_solver.hpp:9-15,  src/pressure_solver.cpp:7-125]
 *   SaturationSolver [REF src/saturation_solver.hpp:9-11, src/saturation_solver.cpp:19-112]
 *   (MPI analogues   [REF src/pressure_solver.hpp:18-25] [REF src/saturation_solver.hpp:14-16])
 *
 * Plain C: pointers, sizes and PODs only.  No MFEM, torch or CUDA types cross
 * this boundary.  All field data are fp64.  A handle owns every device
 * allocation (coefficients, Krylov work vectors, halo buffers, NCCL
 * communicator); it is used by one host thread at a time and issues all of
 * its work on one CUDA stream.
 *
 * There is NO CPU fallback: every compute entry point returns IMPES_ERR_CUDA
 * if no sm_100-class device can be used.
 *
 * Native dof layout (what "native" means for impes_get_field/impes_set_field
 * and for the vectors of impes_pressure_solve / impes_saturation_advance):
 *
 *   cells   c = ix + nx*(iy + ny*iz)                   (x fastest; the order of
 *           the reference's K/phi files [REF src/param.hpp:141-144]
 *           [REF scripts/random_field.py:10-18])
 *   RT_k velocity: one node grid per component d, concatenated [ux | uy | uz]:
 *           size along d        n_d*(k+1)+1   ("closed" nodes, shared by cells)
 *           size along e != d   n_e*(k+1)     ("open" nodes)
 *           x fastest inside each grid.  Every dof is oriented along +d.
 *           RT0: the dof is the total flux through the face.
 *   L2_k pressure / DG_k saturation: cell*(k+1)^dim + local (local x fastest),
 *           i.e. MFEM's L2 ordering; for k=0 simply the cell index.
 *   Darcy vector x = [velocity dofs | pressure dofs]
 *           [REF src/run_serial.cpp:66-81].
 *
 * Multi-GPU (one process per GPU): the grid is cut into slabs along the
 * slowest direction (z in 3-D, y in 2-D): balanced, the first (n % P) ranks one
 * layer more; with IMPES_SCHUR_MG in units of 2^m layers so that multigrid
 * aggregates never straddle GPUs.  impes_get_dims reports the slab
 * (slab_lo, slab_hi).  Host-side arrays passed through this ABI are always GLOBAL
 * (full-mesh) arrays in native layout, like the reference where every rank holds
 * the full K/phi arrays [REF src/run_parallel.cpp:83-95]; each rank reads/writes
 * only the part it owns, see INTEGRATION.md.
 *
 * Supported spaces (anything else: impes_create -> IMPES_ERR_UNSUPPORTED):
 *   order_v/order_p/order_s = 0/0/0   RT0 / P0 / P0
 *                             0/0/1   RT0 / P0 / DG-Q1
 *                             1/1/0   RT1 / L2-Q1 / DG_0   (the reference's defaults [REF src/param.cpp:43])
 *                             1/1/1   RT1 / L2-Q1 / DG-Q1  (BASELINE config C5)
 */
#ifndef IMPES_B200_H
#define IMPES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMPES_ABI_VERSION 1

/* status codes (SURVEY 8b "Errors") */
#define IMPES_OK                0
#define IMPES_ERR_BAD_ARG       1
#define IMPES_ERR_NOT_CONVERGED 2 /* reference: MFEM_ABORT [REF src/pressure_solver.cpp:110-117] */
#define IMPES_ERR_NONFINITE     3
#define IMPES_ERR_CUDA          4
#define IMPES_ERR_NCCL          5
#define IMPES_ERR_UNSUPPORTED   6

/* Schur-block preconditioner of the pressure solve */
#define IMPES_SCHUR_SGS_LEX 0 /* 1 symmetric Gauss-Seidel sweep, lexicographic order:
                                 GSSmoother(S) of the serial reference
                                 [REF src/pressure_solver.cpp:83-91] -- parity mode */
#define IMPES_SCHUR_MG      1 /* geometric multigrid V-cycle on the same S -- throughput
                                 mode (stands where the MPI reference uses BoomerAMG
                                 [REF src/par_pressure_solver.cpp:89-90]); distributed over
                                 the slabs, result independent of the number of GPUs */

/* sub-step count rule of the saturation loop */
#define IMPES_SUBSTEPS_SERIAL   0 /* nt = int(global_dt/dt_local)   [REF src/saturation_solver.cpp:74-76] */
#define IMPES_SUBSTEPS_PARALLEL 1 /* while t < global_dt - dt/2     [REF src/par_saturation_solver.cpp:149-155] */

/* fields for impes_get_field / impes_set_field */
#define IMPES_FIELD_SATURATION 0 /* S dofs                            */
#define IMPES_FIELD_VELOCITY   1 /* V dofs  (block 0 of x)            */
#define IMPES_FIELD_PRESSURE   2 /* P dofs  (block 1 of x)            */
#define IMPES_FIELD_DARCY_X    3 /* [V | P]                           */

/* kernels that impes_time_kernel can time in isolation */
#define IMPES_KERNEL_DARCY_APPLY   1 /* K1: y = [M(S) B^T; B 0] x       */
#define IMPES_KERNEL_SAT_SUBSTEP   2 /* K6: one forward-Euler sub-step  */
#define IMPES_KERNEL_MINRES_UPDATE 3 /* K2: fused vector update passes  */
#define IMPES_KERNEL_PREC_APPLY    4 /* K3: block preconditioner        */
#define IMPES_KERNEL_COEFF         5 /* K4: mobility / diag(M) / Schur stencil */

/* A well: box in 2-D, box in (x,z) extruded along y in 3-D
 * [REF src/param.hpp:310-345] [REF src/param.cpp:26-30]. */
typedef struct impes_well {
  double center[3];
  double radius;
  double height;
} impes_well;

/* Everything the two solver calls read from `Param`
 * [REF src/param.hpp:40-122] plus the compile-time viscosities
 * [REF config.hpp.in:37,43] and the device/rank placement. */
typedef struct impes_desc {
  int32_t struct_size;       /* = sizeof(impes_desc); checked by impes_create */
  int32_t dim;               /* -d   : 2 or 3                                  */
  int32_t nx, ny, nz;        /* -nx -ny -nz (nz ignored when dim == 2)         */
  int32_t order_v;           /* -ov  : RT_k                                    */
  int32_t order_p;           /* -op  : L2_k                                    */
  int32_t order_s;           /* -os  : DG_k                                    */
  double  sx, sy, sz;        /* -sx -sy -sz                                    */
  int32_t two_phase;         /* -two-phase / -single-phase                     */
  int32_t ode_solver_type;   /* -ode : 1 FE, 2 RK2, 3 RK3-SSP, 4 RK4 (6 unsupported) */
  double  Sor, Swc;          /* -Sor -Swc                                      */
  double  mu_w, mu_o;        /* MU_W, MU_O (reference: 1.0, 1.0)               */
  impes_well injection;      /* -inj-*                                         */
  impes_well production;     /* -pro-*                                         */
  double  inflow, outflow;   /* -inflow -outflow                               */
  double  saturation_source; /* -sat-source                                    */
  double  darcy_rtol;        /* -darcy-rtol                                    */
  double  darcy_atol;        /* -darcy-atol                                    */
  int32_t darcy_maxiter;     /* -darcy-maxiter                                 */
  int32_t darcy_print_level; /* -darcy-print-level                             */
  int32_t schur_prec;        /* IMPES_SCHUR_*                                  */
  int32_t substep_rule;      /* IMPES_SUBSTEPS_*                               */
  int32_t rank;              /* this process' slab index, 0 .. world_size-1    */
  int32_t world_size;        /* number of slabs / GPUs (1 = single GPU)        */
  int32_t device;            /* CUDA device ordinal to use                     */
  int32_t mg_smooth;         /* MG: pre-/post-smoothing sweeps, rounded up to even (0 -> default 2) */
} impes_desc;

/* What the reference prints after a solve
 * [REF src/pressure_solver.cpp:103-117]. */
typedef struct impes_solve_stats {
  int32_t iterations;   /* solver.GetNumIterations() */
  int32_t converged;    /* solver.GetConverged()     */
  double  final_norm;   /* solver.GetFinalNorm() = |eta| */
  double  initial_norm; /* eta_0 = sqrt(<P r0, r0>)  */
  double  seconds;      /* wall time of the call     */
  double  device_ms;    /* CUDA-event time of the device work */
} impes_solve_stats;

typedef struct impes_dims {
  int64_t n_cells;      /* global */
  int64_t n_v, n_p, n_s;/* global dof counts (block_offsets of the reference) */
  int64_t n_v_comp[3];  /* global size of each velocity component grid */
  int32_t slab_lo, slab_hi; /* owned cell-layer range [lo, hi) along the slab direction */
  int32_t pad_;
} impes_dims;

typedef struct impes_ctx* impes_handle;

/* --- lifetime ----------------------------------------------------------- */

int  impes_abi_version(void);

/* Fill a descriptor with the reference defaults [REF src/param.cpp:8-13,35-68]. */
void impes_desc_defaults(impes_desc* d);

/* Create a solver context.  Kx/Ky/Kz/phi: host, fp64, GLOBAL n_cells entries,
 * x fastest (Kz may be NULL when dim == 2) -- the arrays Param::init_arrays
 * builds [REF src/param.cpp:78-113].  They are copied; the caller keeps
 * ownership (own_array=false in the reference [REF src/pressure_solver.cpp:19-29]). */
int  impes_create(const impes_desc* d, const double* Kx, const double* Ky,
                  const double* Kz, const double* phi, impes_handle* out);
void impes_destroy(impes_handle h);
const char* impes_last_error(impes_handle h); /* h may be NULL: error of the last failed impes_create */

int  impes_get_dims(impes_handle h, impes_dims* out);

/* --- multi-GPU plumbing (world_size > 1) -------------------------------- */

/* NCCL bootstrap: rank 0 calls impes_comm_unique_id, ships the 128 bytes to the
 * other ranks by any means (torch.distributed / MPI_Bcast), then every rank
 * calls impes_comm_init.  Replaces MPI_COMM_WORLD of the reference
 * [REF src/par_pressure_solver.cpp:99]. */
int  impes_comm_unique_id(void* id128);
int  impes_comm_init(impes_handle h, const void* id128);

/* --- the two solver calls ------------------------------------------------ */

/* PressureSolver [REF src/pressure_solver.cpp:7-125]: one mixed-Darcy solve with
 * the saturation frozen.  S_host: saturation dofs (NULL = use the resident S).
 * x_host: receives [V | P] (NULL = leave the result on the device).  x is
 * zeroed before the solve like the reference [REF :101].  Returns
 * IMPES_ERR_NOT_CONVERGED where the reference aborts; stats are filled either way. */
int  impes_pressure_solve(impes_handle h, const double* S_host, double* x_host,
                          impes_solve_stats* stats);

/* SaturationSolver [REF src/saturation_solver.cpp:19-112]: explicit sub-steps
 * S <- S + dt*M_phi^-1 (K f(S) + b) with the velocity frozen.  V_host: velocity
 * dofs (NULL = resident result of the last pressure solve).  S_inout: in/out
 * saturation (NULL = resident).  n_substeps_out (may be NULL) receives the count. */
int  impes_saturation_advance(impes_handle h, const double* V_host, double* S_inout,
                              int global_ti, double global_dt, double dt_local,
                              int* n_substeps_out);

/* --- slab-local variants (one process per GPU, world_size > 1) --------------
 * The MPI reference keeps only local dofs on a rank (ParGridFunction / true-dof
 * vectors [REF src/run_parallel.cpp:97-104,118-125]) and moves shared dofs only
 * (V.Distribute [REF src/run_parallel.cpp:186]).  These entry points are the
 * same two solver calls with LOCAL host arrays: the native layout restricted to
 * what the rank owns -- its cell layers [slab_lo, slab_hi) of every cell-based
 * array and of the x / y velocity grids, and the slab-axis node/face layers it
 * owns (a cell owns its lower faces; the last rank also owns the top boundary
 * layer).  x_local = [ux_loc | uy_loc | uz_loc | p_loc].  Ghost layers never
 * cross the host boundary: the library exchanges them between GPUs (NCCL).
 * With world_size == 1 local == global. */
typedef struct impes_local_dims {
  int64_t n_v, n_p, n_s;       /* local dof counts                                      */
  int64_t n_v_comp[3];         /* local size of each velocity component grid            */
  int64_t v_comp_start[3];     /* first global index of the local part inside its global component grid */
  int64_t p_start, s_start;    /* first global pressure / saturation dof of the local part */
} impes_local_dims;
int  impes_get_local_dims(impes_handle h, impes_local_dims* out);
/* ParPressureSolver [REF src/par_pressure_solver.cpp:8-136] with local arrays (x_local plays trueX). */
int  impes_pressure_solve_local(impes_handle h, const double* S_local, double* x_local,
                                impes_solve_stats* stats);
/* ParSaturationSolver [REF src/par_saturation_solver.cpp:64-177] with local arrays. */
int  impes_saturation_advance_local(impes_handle h, const double* V_local, double* S_local_inout,
                                    int global_ti, double global_dt, double dt_local,
                                    int* n_substeps_out);
int  impes_set_field_local(impes_handle h, int which, const double* host_local);
int  impes_get_field_local(impes_handle h, int which, double* host_local);

/* Guards of the global loop [REF src/run_serial.cpp:132-142]
 * [REF src/run_parallel.cpp:220-230]: number of non-finite S dofs, max S
 * (all-reduced over ranks). */
int  impes_guards(impes_handle h, int64_t* n_nonfinite, double* s_max);

/* --- resident state access (checkpoint / parity dumps) ------------------- */
int  impes_set_field(impes_handle h, int which, const double* host);
int  impes_get_field(impes_handle h, int which, double* host);

/* --- operator-level entry points (used by the parity tests) -------------- */

/* Rebuild the S-dependent coefficients (mobility, diag(M), Schur stencil) from the
 * resident saturation; impes_pressure_solve does this itself. */
int  impes_update_coefficients(impes_handle h);
/* y = [M(S) B^T; B 0] x with the essential-BC rows/columns eliminated
 * [REF src/pressure_solver.cpp:56-74]; x, y host vectors of size n_v+n_p (global). */
int  impes_darcy_apply(impes_handle h, const double* x_host, double* y_host);
/* y = diag(diag(M)^-1, SchurPrec) r  [REF src/pressure_solver.cpp:83-91]. */
int  impes_prec_apply(impes_handle h, const double* r_host, double* y_host);
/* rhs = [0 | -(Q, psi)]  [REF src/pressure_solver.cpp:35-45]. */
int  impes_get_rhs(impes_handle h, double* rhs_host);

/* --- measurement --------------------------------------------------------- */

/* Time `reps` launches of one kernel in isolation with CUDA events on the
 * handle's stream; when flush_l2 != 0 a >L2-sized buffer is rewritten before
 * every timed launch (outside the event pair).  ms_avg / ms_min receive the
 * per-launch duration. */
int  impes_time_kernel(impes_handle h, int kernel, int reps, int flush_l2,
                       double* ms_avg, double* ms_min);
/* CUDA events on the handle's stream (torch.cuda.Event cannot see it): record slot 0..7, then read
 * the elapsed device milliseconds between two recorded slots (synchronises on the later one). */
int  impes_event_record(impes_handle h, int slot);
int  impes_event_elapsed_ms(impes_handle h, int slot_from, int slot_to, double* ms);
/* Number of kernels this handle has launched since creation. */
int64_t impes_launch_count(impes_handle h);
/* One line naming the kernel variants this handle selected at creation for its mesh and options (spaces, Schur
 * preconditioner and which of its sweep kernels, multigrid levels, saturation kernel, CUDA graphs).  Returns the
 * length of the full description; at most buflen - 1 characters are written. */
int  impes_describe(impes_handle h, char* buf, int buflen);
/* Device-side milliseconds spent per kernel class since the last reset when
 * profiling is on (impes_profile_enable(h,1)); used for the step breakdown. */
int  impes_profile_enable(impes_handle h, int on);
int  impes_profile_read(impes_handle h, double* ms_by_kernel /* [8] */, int64_t* launches_by_kernel /* [8] */);
/* The same per kernel: fills up to `cap` entries (static name strings, milliseconds, launches) and returns the
 * number of kernel ids the library distinguishes. */
int  impes_profile_read_kernels(impes_handle h, int cap, const char** names, double* ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* IMPES_B200_H */
