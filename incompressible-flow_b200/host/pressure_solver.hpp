// pressure_solver.hpp -- PressureSolver with the reference's signature [REF src/pressure_solver.hpp:9-15],
// executed by the B200 library through the C ABI (include/impes_b200.h).
#pragma once
#include "config.hpp"

struct Param;

void PressureSolver(const mfem::Array<int>& block_offsets, const mfem::Mesh& mesh, const Param& param,
                    mfem::FiniteElementSpace& V_space, mfem::FiniteElementSpace& P_space,
                    mfem::GridFunctionCoefficient& saturation, mfem::BlockVector& x);
