#include "saturation_solver.hpp"

#include <stdexcept>
#include <string>
#include <vector>

#include "param.hpp"
#include "solver_context.hpp"

using namespace std;
using namespace mfem;

namespace {
void saturation_advance(impes_handle h, const Param& param, GridFunction& S, VectorCoefficient& velocity, int global_ti,
                        double global_dt) {
  VectorGridFunctionCoefficient* vgf = dynamic_cast<VectorGridFunctionCoefficient*>(&velocity);
  MFEM_VERIFY(vgf && vgf->GetGridFunction(), "SaturationSolver: velocity must wrap a GridFunction");
  GridFunction* V = vgf->GetGridFunction();
  impes_dims dims;
  impes_get_dims(h, &dims);
  MFEM_VERIFY(S.Size() == dims.n_s && V->Size() == dims.n_v, "Dimensions mismatch");
  const DofMap& mv = impes_host_dofmap(param, *V->FESpace(), true, param.order_v);
  const DofMap& ms = impes_host_dofmap(param, *S.FESpace(), false, param.order_s);
  int rc;
  if (mv.identity && ms.identity) {
    rc = impes_saturation_advance(h, V->GetData(), S.GetData(), global_ti, global_dt, param.dt_local, nullptr);
  } else {
    vector<double> Vn(dims.n_v), Sn(dims.n_s);
    mv.to_native(V->GetData(), Vn.data());
    ms.to_native(S.GetData(), Sn.data());
    rc = impes_saturation_advance(h, Vn.data(), Sn.data(), global_ti, global_dt, param.dt_local, nullptr);
    ms.from_native(Sn.data(), S.GetData());
  }
  if (rc == IMPES_ERR_UNSUPPORTED) throw runtime_error(impes_last_error(h));  // "Uknown ODE solver: type N" [REF src/par_saturation_solver.cpp:121-123]
  if (rc != IMPES_OK) MFEM_ABORT("impes_saturation_advance failed: " + string(impes_last_error(h)));
}
}  // namespace

// Explicit upwind sub-steps with the velocity frozen [REF src/saturation_solver.cpp:19-112]:
// nt = int(global_dt / dt_local) forward-Euler steps  S <- S + dt M_phi^-1 (K f(S) + b)  [REF :74-91].
// The velocity reaches the reference as a VectorCoefficient wrapping the RT GridFunction
// [REF src/run_serial.cpp:86]; the adapter unwraps it and ships the dofs.
void SaturationSolver(const Param& param, GridFunction& S, VectorCoefficient& velocity, int global_ti,
                      double global_dt) {
  saturation_advance(impes_host_context(param), param, S, velocity, global_ti, global_dt);
}

// MPI-signature variant [REF src/saturation_solver.hpp:14-16] [REF src/par_saturation_solver.cpp:64-177]: the time loop
// runs while t < global_dt - dt/2 and honours -ode (1 FE, 2 RK2, 3 RK3-SSP, 4 RK4; anything else throws like the
// reference).  The VisIt collection is only written to by the reference when -vsl > 0 [REF :157-166]: off the hot path.
void ParSaturationSolver(const Param& param, ParGridFunction& S, VectorCoefficient& velocity, int global_ti,
                         double global_dt, VisItDataCollection& visit) {
  (void)visit;
  ParFiniteElementSpace* pfes = dynamic_cast<ParFiniteElementSpace*>(S.FESpace());
  const int rank = pfes ? pfes->GetParMesh()->GetMyRank() : 0, world = pfes ? pfes->GetParMesh()->GetNRanks() : 1;
  impes_handle h = impes_host_context(param, IMPES_SUBSTEPS_PARALLEL, rank, world);
  if (world == 1) {
    saturation_advance(h, param, S, velocity, global_ti, global_dt);
    return;
  }
  impes_host_comm_bootstrap(param, rank, world);
  VectorGridFunctionCoefficient* vgf = dynamic_cast<VectorGridFunctionCoefficient*>(&velocity);
  MFEM_VERIFY(vgf && vgf->GetGridFunction(), "ParSaturationSolver: velocity must wrap a GridFunction");
  impes_local_dims ld;
  impes_get_local_dims(h, &ld);
  MFEM_VERIFY(S.Size() == ld.n_s && vgf->GetGridFunction()->Size() == ld.n_v, "ParSaturationSolver: local sizes mismatch");
  const int rc = impes_saturation_advance_local(h, vgf->GetGridFunction()->GetData(), S.GetData(), global_ti, global_dt,
                                                param.dt_local, nullptr);
  if (rc == IMPES_ERR_UNSUPPORTED) throw runtime_error(impes_last_error(h));
  if (rc != IMPES_OK) MFEM_ABORT("impes_saturation_advance_local failed: " + string(impes_last_error(h)));
}
