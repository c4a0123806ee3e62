// pressure_solver.hpp -- PressureSolver / ParPressureSolver with the reference's signatures
// [REF src/pressure_solver.hpp:9-15,18-25], executed by the B200 library through the C ABI (include/impes_b200.h).
#pragma once
#include "config.hpp"

struct Param;

void PressureSolver(const mfem::Array<int>& block_offsets, const mfem::Mesh& mesh, const Param& param,
                    mfem::FiniteElementSpace& V_space, mfem::FiniteElementSpace& P_space,
                    mfem::GridFunctionCoefficient& saturation, mfem::BlockVector& x);

void ParPressureSolver(const mfem::Array<int>& block_offsets, const mfem::Array<int>& block_trueOffsets,
                       const mfem::ParMesh& mesh, const Param& param, mfem::ParFiniteElementSpace& V_space,
                       mfem::ParFiniteElementSpace& P_space, mfem::GridFunctionCoefficient& saturation,
                       mfem::BlockVector& x, mfem::BlockVector& trueX);
