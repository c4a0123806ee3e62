// solver_context.hpp -- the state the reference re-creates on every call (assembled matrices,
// preconditioners) lives, in the GPU build, in one impes_handle per Param.  The reference's solver
// functions are stateless and re-read Param on every call [REF src/pressure_solver.cpp:19-54,119-124]; the adapter
// keeps that contract: the handle is looked up from the Param address and RE-VALIDATED on every call against the
// descriptor and the K / phi arrays it was created from (any changed option, a reloaded permeability field or a new
// Param at a recycled address gets a fresh handle).  Destroyed at exit or by impes_host_release.
#pragma once
#include "../../include/impes_b200.h"
#include "dofmap.hpp"

struct Param;

// rule: IMPES_SUBSTEPS_SERIAL for run_serial's solvers, IMPES_SUBSTEPS_PARALLEL (+ -ode) for run_parallel's;
// rank / world_size: the slab this process owns (one process per GPU)
impes_handle impes_host_context(const Param& param, int substep_rule = IMPES_SUBSTEPS_SERIAL, int rank = 0,
                                int world_size = 1);
void impes_host_release(const Param& param);
// NCCL bootstrap of a multi-process run (one process per GPU): rank 0 creates the unique id, the others receive it.
// Transport: the file named by IMPES_NCCL_ID_FILE (written atomically by rank 0, polled by the others) -- with real
// MFEM + MPI an MPI_Bcast does the same [REF src/par_pressure_solver.cpp:99 uses MPI_COMM_WORLD].  Idempotent.
void impes_host_comm_bootstrap(const Param& param, int rank, int world_size);

// MFEM numbering <-> native layout of a space, built on first use (see dofmap.hpp)
const DofMap& impes_host_dofmap(const Param& param, mfem::FiniteElementSpace& fes, bool rt, int order);
