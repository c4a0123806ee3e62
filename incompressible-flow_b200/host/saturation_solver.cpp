#include "saturation_solver.hpp"

#include <string>

#include "param.hpp"
#include "solver_context.hpp"

using namespace std;
using namespace mfem;

// Explicit upwind sub-steps with the velocity frozen [REF src/saturation_solver.cpp:19-112]:
// nt = int(global_dt / dt_local) forward-Euler steps  S <- S + dt M_phi^-1 (K f(S) + b)  [REF :74-91].
// The velocity reaches the reference as a VectorCoefficient wrapping the RT GridFunction
// [REF src/run_serial.cpp:86]; the adapter unwraps it and ships the dofs.
void SaturationSolver(const Param& param, GridFunction& S, VectorCoefficient& velocity, int global_ti,
                      double global_dt) {
  impes_handle h = impes_host_context(param);
  VectorGridFunctionCoefficient* vgf = dynamic_cast<VectorGridFunctionCoefficient*>(&velocity);
  MFEM_VERIFY(vgf && vgf->GetGridFunction(), "SaturationSolver: velocity must wrap a GridFunction");
  impes_dims dims;
  impes_get_dims(h, &dims);
  MFEM_VERIFY(S.Size() == dims.n_s && vgf->GetGridFunction()->Size() == dims.n_v, "Dimensions mismatch");
  const int rc = impes_saturation_advance(h, vgf->GetGridFunction()->GetData(), S.GetData(), global_ti, global_dt,
                                          param.dt_local, nullptr);
  if (rc != IMPES_OK) MFEM_ABORT("impes_saturation_advance failed: " + string(impes_last_error(h)));
}
