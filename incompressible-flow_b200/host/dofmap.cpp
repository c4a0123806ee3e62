#include "dofmap.hpp"

#include <cmath>
#include <string>

using namespace mfem;

DofMap build_dofmap(FiniteElementSpace& fes, bool rt, int order, int dim, const int* n, const double* size) {
  MFEM_VERIFY(order == 0 || order == 1, "dof map: orders 0 and 1 only");
  const int k1 = order + 1;
  DofMap m;
  const int nd = fes.GetVSize();
  m.nat.assign(nd, -1);
  m.sgn.assign(nd, 1);
  double h[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d) h[d] = size[d] / n[d];
  // native component offsets / extents of the RT node grids
  long off[3] = {0, 0, 0}, ext[3][3];
  for (int d = 0; d < dim; ++d) {
    long sz = 1;
    for (int e = 0; e < dim; ++e) { ext[d][e] = (e == d) ? (long)n[e] * k1 + 1 : (long)n[e] * k1; sz *= ext[d][e]; }
    if (d + 1 < dim) off[d + 1] = off[d] + sz;
  }
  long npl = 1;
  for (int d = 0; d < dim; ++d) npl *= k1;
  Array<int> dofs;
  DenseMatrix shape;
  Vector X(dim);
  for (int e = 0; e < fes.GetNE(); ++e) {
    const FiniteElement* fe = fes.GetFE(e);
    ElementTransformation* T = fes.GetElementTransformation(e);
    const IntegrationRule& nodes = fe->GetNodes();
    fes.GetElementDofs(e, dofs);
    MFEM_VERIFY(dofs.Size() == nodes.GetNPoints(), "dof map: element without nodal dofs");
    for (int k = 0; k < dofs.Size(); ++k) {
      const IntegrationPoint& ip = nodes.IntPoint(k);
      T->Transform(ip, X);
      const int raw = dofs[k], g = raw >= 0 ? raw : -1 - raw;
      int dir = -1;
      double s = raw >= 0 ? 1.0 : -1.0;
      if (rt) {  // direction and local orientation of basis function k at its own node
        T->SetIntPoint(&ip);
        fe->CalcVShape(*T, shape);
        double best = 0.0;
        for (int d = 0; d < dim; ++d)
          if (std::fabs(shape(k, d)) > std::fabs(best)) { best = shape(k, d); dir = d; }
        MFEM_VERIFY(dir >= 0, "dof map: vanishing RT basis function at its node");
        if (best < 0) s = -s;
      }
      long gi[3] = {0, 0, 0}, cell[3] = {0, 0, 0}, loc[3] = {0, 0, 0};
      for (int d = 0; d < dim; ++d) {
        const double u = X(d) / h[d];
        if (d == dir) gi[d] = std::lround(u * k1);  // closed nodes: multiples of 1/(k+1)
        else {                                       // open (Gauss) nodes: strictly inside the cell
          long c = (long)std::floor(u);
          c = c < 0 ? 0 : (c >= n[d] ? n[d] - 1 : c);
          cell[d] = c;
          loc[d] = (order == 1 && u - c > 0.5) ? 1 : 0;
          gi[d] = c * k1 + loc[d];
        }
      }
      long nat;
      if (rt) nat = off[dir] + gi[0] + ext[dir][0] * (gi[1] + (dim == 3 ? ext[dir][1] * gi[2] : 0));
      else {
        const long c = cell[0] + (long)n[0] * (cell[1] + (dim == 3 ? (long)n[1] * cell[2] : 0));
        nat = c * npl + loc[0] + k1 * (loc[1] + (dim == 3 ? k1 * loc[2] : 0));
      }
      MFEM_VERIFY(nat >= 0 && nat < nd, "dof map: node outside the native grid");
      MFEM_VERIFY(m.nat[g] < 0 || (m.nat[g] == nat && m.sgn[g] == (s > 0 ? 1 : -1)),
                  "dof map: two elements disagree about a shared dof");
      m.nat[g] = (int)nat;
      m.sgn[g] = s > 0 ? 1 : -1;
    }
  }
  for (int g = 0; g < nd; ++g) {
    MFEM_VERIFY(m.nat[g] >= 0, "dof map: a dof belongs to no element");
    if (m.nat[g] != g || m.sgn[g] != 1) m.identity = false;
  }
  return m;
}
