// dofmap.hpp -- MFEM dof numbering <-> the library's native layout (include/impes_b200.h).
//
// The reference hands the solvers vectors in MFEM's own numbering: x = [V | P] with V in RT_k and P in L2_k
// [REF src/run_serial.cpp:66-81], S in DG_k [REF src/run_serial.cpp:83-84].  MFEM numbers RT dofs by mesh entity
// (faces, then element interiors) and orients a face dof by the face's first element; the native layout is one
// node grid per velocity component with every dof along +d.  Instead of restating MFEM's numbering rules (which
// differ between versions), the map is IDENTIFIED from MFEM's public element interface at run time: for every element
// and local dof, the node position (FiniteElement::GetNodes + ElementTransformation::Transform), for RT the direction
// and orientation of the basis function at its own node (FiniteElement::CalcVShape), and the signed global index
// (FiniteElementSpace::GetElementDofs, negative = -1-j: flipped).  Works for orders k <= 1 on the Cartesian meshes the
// reference builds.  Values: native = sign * mfem (the dof functionals agree up to orientation).
#pragma once
#include <vector>

#include "config.hpp"

struct DofMap {
  std::vector<int> nat;          // nat[g]: native index of MFEM dof g
  std::vector<signed char> sgn;  // +1 / -1: native value = sgn[g] * MFEM value
  bool identity = true;          // nat[g] == g and sgn[g] == +1 for all g: vectors pass through untouched
  void to_native(const double* mfem, double* native) const {
    for (size_t g = 0; g < nat.size(); ++g) native[nat[g]] = sgn[g] * mfem[g];
  }
  void from_native(const double* native, double* mfem) const {
    for (size_t g = 0; g < nat.size(); ++g) mfem[g] = sgn[g] * native[nat[g]];
  }
};

// order: polynomial order k of the space (RT_k or L2_k); n[], size[]: cells and extents per direction
DofMap build_dofmap(mfem::FiniteElementSpace& fes, bool rt, int order, int dim, const int* n, const double* size);
