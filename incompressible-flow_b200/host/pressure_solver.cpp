#include "pressure_solver.hpp"

#include <iostream>

#include "param.hpp"
#include "solver_context.hpp"

using namespace std;
using namespace mfem;

// One mixed-Darcy solve with the saturation frozen [REF src/pressure_solver.cpp:7-125].
// Where the reference assembles M(S), B, the Schur approximation and runs MINRES on the CPU, this uploads
// the saturation dofs, runs the matrix-free preconditioned MINRES on the GPU and downloads x = [V | P].
// Output lines and the abort-on-non-convergence policy are the reference's [REF :103-117].
void PressureSolver(const Array<int>& block_offsets, const Mesh& mesh, const Param& param, FiniteElementSpace& V_space,
                    FiniteElementSpace& P_space, GridFunctionCoefficient& saturation, BlockVector& x) {
  (void)mesh;
  impes_handle h = impes_host_context(param);
  impes_dims dims;
  impes_get_dims(h, &dims);
  MFEM_VERIFY(block_offsets[1] - block_offsets[0] == dims.n_v && V_space.GetVSize() == dims.n_v,
              "Velocity space size mismatch");
  MFEM_VERIFY(block_offsets[2] - block_offsets[1] == dims.n_p && P_space.GetVSize() == dims.n_p,
              "Pressure space size mismatch");
  const GridFunction* S = saturation.GetGridFunction();
  MFEM_VERIFY(S && S->Size() == dims.n_s, "Saturation space size mismatch");

  impes_solve_stats st;
  const int rc = impes_pressure_solve(h, S->GetData(), x.GetData(), &st);
  if (rc == IMPES_OK) {
    cout << "Darcy solver converged. Iter = " << st.iterations << ". Final norm = " << st.final_norm
         << ". Time = " << st.seconds << " sec" << endl;
  } else if (rc == IMPES_ERR_NOT_CONVERGED) {
    cerr << "Darcy solver DIDN'T converge. Iter = " << st.iterations << ". Final norm = " << st.final_norm
         << ". Time = " << st.seconds << " sec" << endl;
    MFEM_ABORT("The solver didn't converge => terminating");
  } else {
    MFEM_ABORT("impes_pressure_solve failed: " + string(impes_last_error(h)));
  }
}
