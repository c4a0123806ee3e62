// saturation_solver.hpp -- SaturationSolver / ParSaturationSolver with the reference's signatures
// [REF src/saturation_solver.hpp:9-11,14-16], executed by the B200 library through the C ABI.
#pragma once
#include "config.hpp"

struct Param;

void SaturationSolver(const Param& param, mfem::GridFunction& S, mfem::VectorCoefficient& velocity, int global_ti,
                      double global_dt);

void ParSaturationSolver(const Param& param, mfem::ParGridFunction& S, mfem::VectorCoefficient& velocity, int global_ti,
                         double global_dt, mfem::VisItDataCollection& visit);
