#include "pressure_solver.hpp"

#include <iostream>
#include <vector>

#include "param.hpp"
#include "solver_context.hpp"

using namespace std;
using namespace mfem;

namespace {
// shared body of the serial and the MPI-signature entry points.  x is in the caller's (MFEM) numbering; the
// library works in its native layout: the dof maps (dofmap.hpp) translate, and vanish when they are the identity.
void pressure_solve(impes_handle h, const Param& param, FiniteElementSpace& V_space, FiniteElementSpace& P_space,
                    GridFunctionCoefficient& saturation, const Array<int>& block_offsets, BlockVector& x, bool print) {
  impes_dims dims;
  impes_get_dims(h, &dims);
  MFEM_VERIFY(block_offsets[1] - block_offsets[0] == dims.n_v && V_space.GetVSize() == dims.n_v,
              "Velocity space size mismatch");
  MFEM_VERIFY(block_offsets[2] - block_offsets[1] == dims.n_p && P_space.GetVSize() == dims.n_p,
              "Pressure space size mismatch");
  const GridFunction* S = saturation.GetGridFunction();
  MFEM_VERIFY(S && S->Size() == dims.n_s, "Saturation space size mismatch");
  const DofMap& mv = impes_host_dofmap(param, V_space, true, param.order_v);
  const DofMap& mp = impes_host_dofmap(param, P_space, false, param.order_p);
  const DofMap& ms = impes_host_dofmap(param, *S->FESpace(), false, param.order_s);
  const bool direct = mv.identity && mp.identity && ms.identity;

  impes_solve_stats st;
  int rc;
  if (direct) {
    rc = impes_pressure_solve(h, S->GetData(), x.GetData(), &st);
  } else {
    vector<double> Sn(dims.n_s), xn(dims.n_v + dims.n_p);
    ms.to_native(S->GetData(), Sn.data());
    rc = impes_pressure_solve(h, Sn.data(), xn.data(), &st);
    mv.from_native(xn.data(), x.GetData() + block_offsets[0]);
    mp.from_native(xn.data() + dims.n_v, x.GetData() + block_offsets[1]);
  }
  if (rc == IMPES_OK) {
    if (print)
      cout << "Darcy solver converged. Iter = " << st.iterations << ". Final norm = " << st.final_norm
           << ". Time = " << st.seconds << " sec" << endl;
  } else if (rc == IMPES_ERR_NOT_CONVERGED) {
    if (print)
      cerr << "Darcy solver DIDN'T converge. Iter = " << st.iterations << ". Final norm = " << st.final_norm
           << ". Time = " << st.seconds << " sec" << endl;
    MFEM_ABORT("The solver didn't converge => terminating");
  } else {
    MFEM_ABORT("impes_pressure_solve failed: " + string(impes_last_error(h)));
  }
}
}  // namespace

// One mixed-Darcy solve with the saturation frozen [REF src/pressure_solver.cpp:7-125].
// Where the reference assembles M(S), B, the Schur approximation and runs MINRES on the CPU, this uploads
// the saturation dofs, runs the matrix-free preconditioned MINRES on the GPU and downloads x = [V | P].
// Output lines and the abort-on-non-convergence policy are the reference's [REF :103-117].
void PressureSolver(const Array<int>& block_offsets, const Mesh& mesh, const Param& param, FiniteElementSpace& V_space,
                    FiniteElementSpace& P_space, GridFunctionCoefficient& saturation, BlockVector& x) {
  (void)mesh;
  pressure_solve(impes_host_context(param), param, V_space, P_space, saturation, block_offsets, x, true);
}

// MPI-signature variant [REF src/pressure_solver.hpp:18-25] [REF src/par_pressure_solver.cpp:8-136]: the GPU library
// takes the place of the rank-local assembly + HYPRE solve.  In this build every process drives ONE GPU holding the
// whole mesh (single-rank "parallel" run: the slab decomposition lives inside the library, see
// impes_*_local and INTEGRATION.md section 5 for the one-process-per-GPU binding); trueX mirrors x like the
// reference's true-dof vector does on one rank [REF :107-108].
void ParPressureSolver(const Array<int>& block_offsets, const Array<int>& block_trueOffsets, const ParMesh& mesh,
                       const Param& param, ParFiniteElementSpace& V_space, ParFiniteElementSpace& P_space,
                       GridFunctionCoefficient& saturation, BlockVector& x, BlockVector& trueX) {
  const int rank = mesh.GetMyRank(), world = mesh.GetNRanks();
  impes_handle h = impes_host_context(param, IMPES_SUBSTEPS_PARALLEL, rank, world);
  if (world == 1) {
    MFEM_VERIFY(block_trueOffsets.Last() == block_offsets.Last(), "ParPressureSolver: one rank owns every dof");
    pressure_solve(h, param, V_space, P_space, saturation, block_offsets, x, true);
    for (int i = 0; i < x.Size(); ++i) trueX(i) = x(i);
    return;
  }
  // one process per GPU: every rank holds the dofs of its slab (native local layout, impes_get_local_dims) -- the
  // local / true dofs of the reference's ParGridFunctions [REF src/run_parallel.cpp:118-125]; ghost layers stay on
  // the GPUs.  Rank 0 prints, like the reference [REF src/par_pressure_solver.cpp:111-127].
  impes_host_comm_bootstrap(param, rank, world);
  impes_local_dims ld;
  impes_get_local_dims(h, &ld);
  MFEM_VERIFY(V_space.GetVSize() == ld.n_v && P_space.GetVSize() == ld.n_p, "ParPressureSolver: local space sizes mismatch");
  const GridFunction* S = saturation.GetGridFunction();
  MFEM_VERIFY(S && S->Size() == ld.n_s, "ParPressureSolver: local saturation size mismatch");
  impes_solve_stats st;
  const int rc = impes_pressure_solve_local(h, S->GetData(), trueX.GetData(), &st);
  for (int i = 0; i < x.Size(); ++i) x(i) = trueX(i);
  if (rc == IMPES_OK) {
    if (rank == 0)
      cout << "Darcy solver converged. Iter = " << st.iterations << ". Final norm = " << st.final_norm
           << ". Time = " << st.seconds << " sec" << endl;
  } else if (rc == IMPES_ERR_NOT_CONVERGED) {
    if (rank == 0)
      cerr << "Darcy solver DIDN'T converge. Iter = " << st.iterations << ". Final norm = " << st.final_norm
           << ". Time = " << st.seconds << " sec" << endl;
    MFEM_ABORT("The solver didn't converge => terminating");
  } else {
    MFEM_ABORT("impes_pressure_solve_local failed: " + string(impes_last_error(h)));
  }
}
