// run_serial.cpp -- the caller of the hot path: option parsing, K/phi loading, the global IMPES loop with
// its guards and the per-step saturation dumps, line-compatible with the reference's driver
// [REF src/run_serial.cpp:14-155] so that whole runs can be diffed against a real reference build.
// VisIt / .vts visualisation output is not reproduced (off the hot path).
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "config.hpp"
#include "dofmap.hpp"
#include "param.hpp"
#include "pressure_solver.hpp"
#include "saturation_solver.hpp"
#include "solver_context.hpp"
#include "utilities.hpp"

using namespace std;
using namespace mfem;

namespace {
double seconds_since(const chrono::steady_clock::time_point& t0) {
  return chrono::duration<double>(chrono::steady_clock::now() - t0).count();
}
// saturation_<extra>_<dim>D_<phase>_<step>.bin, float32, one value per cell [REF src/utilities.cpp:565-582]
// The reference fills the per-cell array through S.ProjectCoefficient(ValuesInCells(saturation))
// [REF src/run_serial.cpp:123-127] [REF src/param.hpp:413-442]: every nodal evaluation of an element overwrites
// the cell's entry, so the value that survives is the one of the element's LAST local dof [MFEM-recall:
// NodalFiniteElement::Project loops over the dofs in order].  DG_0: the cell value; DG-Q1: local dof 2^dim - 1.
void output_saturation_cells(const Param& p, const Vector& S_own, const DofMap& map, int ti) {
  const string phase = p.two_phase_flow ? "2phase" : "1phase";
  const string fname = string(p.outdir) + "/saturation_" + p.extra + "_" + d2s(p.spacedim) + "D_" + phase + "_" +
                       d2s(ti, 0, 0, 0, 6) + ".bin";
  const int n = p.get_n_cells();
  const int nsl = S_own.Size() / n;
  // cell-major view of the dofs (the space's own numbering need not be: dofmap.hpp)
  std::vector<double> nat;
  const double* S = S_own.GetData();
  if (!map.identity) {
    nat.resize(S_own.Size());
    map.to_native(S_own.GetData(), nat.data());
    S = nat.data();
  }
  Vector in_cells(n);
  for (int c = 0; c < n; ++c) in_cells(c) = S[(long)c * nsl + nsl - 1];
  write_binary(fname.c_str(), n, in_cells.GetData());
  // the .vts twin [REF src/utilities.cpp:565-582]
  const string fvts = fname.substr(0, fname.size() - 4) + ".vts";
  if (p.spacedim == 2) write_vts_scalar_cells(fvts, "saturation", p.sx, p.sy, p.nx, p.ny, in_cells);
  else write_vts_scalar_cells(fvts, "saturation", p.sx, p.sy, p.sz, p.nx, p.ny, p.nz, in_cells);
}
}  // namespace

void run_serial(int argc, char** argv) {
  if (argc == 1) {
    cout << "\nTo see the available options, run\n" << argv[0] << " -h\n" << endl;
    throw 1;
  }
  const auto total_time = chrono::steady_clock::now();

  Param p;
  OptionsParser args(argc, argv);
  p.add_options(args);
  args.Parse();
  if (!args.Good()) {
    args.PrintUsage(cout);
    throw 1;
  }
  if (p.info) {
    cout << p.get_info() << endl;
    throw 1;
  }
  args.PrintOptions(cout);
  MFEM_VERIFY(p.order_s == 0 || p.order_s == 1, "The GPU path implements DG_0 and DG_1 saturation");

  p.init_arrays();
  const string cmd = "mkdir -p " + string(p.outdir);
  MFEM_VERIFY(system(cmd.c_str()) == 0, "Failed to create an output directory");

  const int gen_edges = 1;
  Mesh* mesh = nullptr;
  if (p.spacedim == 2) mesh = new Mesh(p.nx, p.ny, Element::QUADRILATERAL, gen_edges, p.sx, p.sy);
  else if (p.spacedim == 3) mesh = new Mesh(p.nx, p.ny, p.nz, Element::HEXAHEDRON, gen_edges, p.sx, p.sy, p.sz);
  else MFEM_ABORT("Unknown spacedim");
  const int dim = mesh->Dimension();

  FiniteElementCollection* hdiv_coll = new RT_FECollection(p.order_v, dim);
  FiniteElementCollection* l2_coll = new L2_FECollection(p.order_p, dim);
  FiniteElementCollection* dg_coll = new DG_FECollection(p.order_s, dim);
  FiniteElementSpace V_space(mesh, hdiv_coll), P_space(mesh, l2_coll), S_space(mesh, dg_coll);

  Array<int> block_offsets(3);
  block_offsets[0] = 0;
  block_offsets[1] = V_space.GetVSize();
  block_offsets[2] = P_space.GetVSize();
  block_offsets.PartialSum();
  cout << "dim(V)  = " << block_offsets[1] - block_offsets[0] << "\n";
  cout << "dim(P)  = " << block_offsets[2] - block_offsets[1] << "\n";
  cout << "dim(V+P)= " << block_offsets.Last() << "\n";
  cout << "dim(S)  = " << S_space.GetVSize() << "\n" << endl;

  BlockVector x(block_offsets);
  x = 0.0;  // the reference leaves x uninitialised before the first solve; boundary flux is 0
  GridFunction V, P;
  V.Update(&V_space, x.GetBlock(0), 0);
  P.Update(&P_space, x.GetBlock(1), 0);
  GridFunction S(&S_space);
  S = 0.0;
  GridFunctionCoefficient saturation(&S);
  VectorGridFunctionCoefficient velocity(&V);

  const int ncell_dir[3] = {p.nx, p.ny, p.nz};
  const double size_dir[3] = {p.sx, p.sy, p.sz};
  const DofMap s_map = build_dofmap(S_space, false, p.order_s, dim, ncell_dir, size_dir);
  if (p.saturation_steps > 0) output_saturation_cells(p, S, s_map, 0);

  const auto global_time_loop = chrono::steady_clock::now();
  const int nt = (int)ceil(p.t_final / p.dt_global);
  cout << "Number of global time steps: " << nt << endl;
  for (int ti = 1; ti <= nt; ++ti) {
    PressureSolver(block_offsets, *mesh, p, V_space, P_space, saturation, x);
    SaturationSolver(p, S, velocity, ti, p.dt_global);

    if (p.saturation_steps > 0 && ti % p.saturation_steps == 0) output_saturation_cells(p, S, s_map, ti);
    cout << "time step " << ti << " is done" << endl;

    if (S.CheckFinite() > 0) {
      cout << "\n\n\tSaturation has nonfinite values. The process stops.\n" << endl;
      break;
    }
    if (S.Max() > 1.1) {
      cout << "\n\n\tSaturation went up to more than 1.1. The process stops.\n" << endl;
      break;
    }
  }
  cout << "Time of the global time loop: " << seconds_since(global_time_loop) << " sec" << endl;

  impes_host_release(p);
  delete hdiv_coll;
  delete l2_coll;
  delete dg_coll;
  delete mesh;
  cout << "TOTAL TIME: " << seconds_since(total_time) << " sec" << endl;
}
