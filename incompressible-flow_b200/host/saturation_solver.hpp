// saturation_solver.hpp -- SaturationSolver with the reference's signature
// [REF src/saturation_solver.hpp:9-11], executed by the B200 library through the C ABI.
#pragma once
#include "config.hpp"

struct Param;

void SaturationSolver(const Param& param, mfem::GridFunction& S, mfem::VectorCoefficient& velocity, int global_ti,
                      double global_dt);
