// run_parallel.cpp -- the MPI-side caller of the hot path [REF src/run_parallel.cpp:18-262], one process per GPU:
// same options, rank-0 log lines, global loop (ParPressureSolver -> V.Distribute -> ParSaturationSolver), guards and
// per-cell dumps.  Where the reference partitions with METIS and talks MPI, the mesh is cut into z-slabs inside the
// library and the ranks talk NCCL; every rank keeps only the dofs of its slab (the reference's local ParGridFunction
// data).  Process placement comes from the environment instead of MPI_Init:
//   IMPES_RANK / IMPES_WORLD_SIZE (or torchrun's RANK / WORLD_SIZE), IMPES_NCCL_ID_FILE (bootstrap, solver_context.hpp);
// the CUDA device is -device + the local rank (IMPES_LOCAL_RANK / LOCAL_RANK, default: the rank).
// VisIt output is not reproduced (off the hot path); pressure and saturation per-cell .bin + .vts dumps are.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "config.hpp"
#include "param.hpp"
#include "pressure_solver.hpp"
#include "saturation_solver.hpp"
#include "solver_context.hpp"
#include "utilities.hpp"

using namespace std;
using namespace mfem;

namespace {
int env_int(const char* a, const char* b, int dflt) {
  const char* e = getenv(a);
  if (!e && b) e = getenv(b);
  return e ? atoi(e) : dflt;
}
double seconds_since(const chrono::steady_clock::time_point& t0) {
  return chrono::duration<double>(chrono::steady_clock::now() - t0).count();
}
void minmax(const double* a, int n, double& mn, double& mx) {
  mn = mx = a[0];
  for (int i = 1; i < n; ++i) { mn = min(mn, a[i]); mx = max(mx, a[i]); }
}

// <name>_<extra>_<dim>D_<phase>_<step>.bin, float32, one value per cell [REF src/utilities.cpp:565-582,587-632]: the
// value of each element's LAST local dof (ValuesInCells semantics, see run_serial.cpp).  Every rank writes the cells of
// its slab into a part file; rank 0 assembles them once all parts of the step exist (the reference funnels the same
// data through MPI_Send / MPI_Recv to rank 0).
void output_cells(const Param& p, const double* local, long n_local_dofs, int dofs_per_cell, long cell_start, int rank,
                  int world, int ti, const string& name) {
  const string phase = p.two_phase_flow ? "2phase" : "1phase";
  const string base = string(p.outdir) + "/" + name + "_" + p.extra + "_" + d2s(p.spacedim) + "D_" + phase + "_" +
                      d2s(ti, 0, 0, 0, 6) + ".bin";
  const long ncl = n_local_dofs / dofs_per_cell;
  vector<float> cells(ncl);
  for (long c = 0; c < ncl; ++c) cells[c] = (float)local[c * dofs_per_cell + dofs_per_cell - 1];
  // the .vts twin of the dump [REF src/utilities.cpp:587-630]: the float32 values the .bin holds, as cell data
  auto write_vts = [&](const vector<float>& v) {
    Vector x((int)v.size());
    for (size_t c = 0; c < v.size(); ++c) x((int)c) = v[c];
    const string fv = base.substr(0, base.size() - 4) + ".vts";
    if (p.spacedim == 2) write_vts_scalar_cells(fv, name, p.sx, p.sy, p.nx, p.ny, x);
    else write_vts_scalar_cells(fv, name, p.sx, p.sy, p.sz, p.nx, p.ny, p.nz, x);
  };
  if (world == 1) {
    FILE* f = fopen(base.c_str(), "wb");
    MFEM_VERIFY(f, "cannot open " + base);
    fwrite(cells.data(), sizeof(float), ncl, f);
    fclose(f);
    write_vts(cells);
    return;
  }
  const string part = base + ".part" + d2s(rank);
  {
    FILE* f = fopen((part + ".tmp").c_str(), "wb");
    MFEM_VERIFY(f, "cannot open " + part);
    fwrite(&cell_start, sizeof cell_start, 1, f);
    fwrite(&ncl, sizeof ncl, 1, f);
    fwrite(cells.data(), sizeof(float), ncl, f);
    fclose(f);
    MFEM_VERIFY(rename((part + ".tmp").c_str(), part.c_str()) == 0, "cannot publish " + part);
  }
  if (rank != 0) return;
  vector<float> all(p.get_n_cells());
  for (int r = 0; r < world; ++r) {
    const string pr = base + ".part" + d2s(r);
    FILE* f = nullptr;
    for (int tries = 0; tries < 6000 && !(f = fopen(pr.c_str(), "rb")); ++tries) this_thread::sleep_for(chrono::milliseconds(100));
    MFEM_VERIFY(f, "timed out waiting for " + pr);
    long start = 0, n = 0;
    MFEM_VERIFY(fread(&start, sizeof start, 1, f) == 1 && fread(&n, sizeof n, 1, f) == 1, "bad part file");
    MFEM_VERIFY(fread(all.data() + start, sizeof(float), n, f) == (size_t)n, "short part file");
    fclose(f);
    remove(pr.c_str());
  }
  FILE* f = fopen(base.c_str(), "wb");
  MFEM_VERIFY(f, "cannot open " + base);
  fwrite(all.data(), sizeof(float), all.size(), f);
  fclose(f);
  write_vts(all);
}
}  // namespace

void run_parallel(int argc, char** argv) {
  const int myid = env_int("IMPES_RANK", "RANK", 0), num_procs = env_int("IMPES_WORLD_SIZE", "WORLD_SIZE", 1);
  const int local_rank = env_int("IMPES_LOCAL_RANK", "LOCAL_RANK", myid);
  if (argc == 1) {
    if (myid == 0) cout << "\nTo see the available options, run\n" << argv[0] << " -h\n" << endl;
    throw 1;
  }
  const auto total_time = chrono::steady_clock::now();

  Param p;
  OptionsParser args(argc, argv);
  p.add_options(args);
  args.Parse();
  if (!args.Good()) {
    if (myid == 0) args.PrintUsage(cout);
    throw 1;
  }
  if (p.info) {
    if (myid == 0) cout << p.get_info() << endl;
    throw 1;
  }
  if (myid == 0) args.PrintOptions(cout);
  p.device += local_rank;
  if (num_procs > 1 && p.schur_prec != IMPES_SCHUR_MG && myid == 0)
    cout << "note: -schur-prec 0 on several GPUs pipelines the exact Gauss-Seidel sweep rank by rank (parity mode)\n";

  p.init_arrays();
  if (myid == 0) {
    const int n = p.get_n_cells();
    double mn, mx;
    minmax(p.K_array_x, n, mn, mx);
    cout << "Kx: min " << mn << " max " << mx << "\n";
    minmax(p.K_array_y, n, mn, mx);
    cout << "Ky: min " << mn << " max " << mx << "\n";
    if (p.spacedim == 3) {
      minmax(p.K_array_z, n, mn, mx);
      cout << "Kz: min " << mn << " max " << mx << "\n";
    }
    minmax(p.phi_array, n, mn, mx);
    cout << "phi: min " << mn << " max " << mx << "\n";
    const string cmd = "mkdir -p " + string(p.outdir);
    MFEM_VERIFY(system(cmd.c_str()) == 0, "Failed to create an output directory");
  }

  Mesh* mesh = nullptr;
  if (p.spacedim == 2) mesh = new Mesh(p.nx, p.ny, Element::QUADRILATERAL, 1, p.sx, p.sy);
  else if (p.spacedim == 3) mesh = new Mesh(p.nx, p.ny, p.nz, Element::HEXAHEDRON, 1, p.sx, p.sy, p.sz);
  else MFEM_ABORT("Unknown spacedim");
  ParMesh* pmesh = new ParMesh(*mesh, myid, num_procs);
  delete mesh;

  // the library decides the slabs; the spaces of this rank hold exactly the dofs it owns
  impes_handle h = impes_host_context(p, IMPES_SUBSTEPS_PARALLEL, myid, num_procs);
  impes_dims gd;
  impes_local_dims ld;
  impes_get_dims(h, &gd);
  impes_get_local_dims(h, &ld);

  FiniteElementCollection* hdiv_coll = new RT_FECollection(p.order_v, p.spacedim);
  FiniteElementCollection* l2_coll = new L2_FECollection(p.order_p, p.spacedim);
  FiniteElementCollection* dg_coll = new DG_FECollection(p.order_s, p.spacedim);
  ParFiniteElementSpace V_space(pmesh, hdiv_coll, (int)ld.n_v), P_space(pmesh, l2_coll, (int)ld.n_p),
      S_space(pmesh, dg_coll, (int)ld.n_s);
  if (myid == 0) {
    cout << "dim(V)  = " << gd.n_v << "\n";
    cout << "dim(P)  = " << gd.n_p << "\n";
    cout << "dim(V+P)= " << gd.n_v + gd.n_p << "\n";
    cout << "dim(S)  = " << gd.n_s << "\n" << endl;
  }

  Array<int> block_offsets(3), block_trueOffsets(3);
  block_offsets[0] = 0; block_offsets[1] = V_space.GetVSize(); block_offsets[2] = P_space.GetVSize();
  block_offsets.PartialSum();
  block_trueOffsets[0] = 0; block_trueOffsets[1] = V_space.TrueVSize(); block_trueOffsets[2] = P_space.TrueVSize();
  block_trueOffsets.PartialSum();
  BlockVector x(block_offsets), trueX(block_trueOffsets);
  x = 0.0;
  trueX = 0.0;
  ParGridFunction V, P;
  V.Update(&V_space, x.GetBlock(0), 0);
  P.Update(&P_space, x.GetBlock(1), 0);
  ParGridFunction S(&S_space);
  S = 0.0;
  GridFunctionCoefficient saturation(&S);
  VectorGridFunctionCoefficient velocity(&V);

  const int nsl = (int)(gd.n_s / gd.n_cells), npl = (int)(gd.n_p / gd.n_cells);
  const long cell_start = ld.s_start / nsl;
  if (p.saturation_steps > 0) output_cells(p, S.GetData(), ld.n_s, nsl, cell_start, myid, num_procs, 0, "saturation");

  VisItDataCollection visit_local("inc-flow-parallel-local", pmesh);
  const auto global_time_loop = chrono::steady_clock::now();
  const int nt = (int)ceil(p.t_final / p.dt_global);
  if (myid == 0) cout << "Number of global time steps: " << nt << endl;
  int valid_loop = 1;
  for (int ti = 1; ti <= nt && valid_loop; ++ti) {
    ParPressureSolver(block_offsets, block_trueOffsets, *pmesh, p, V_space, P_space, saturation, x, trueX);
    V.Distribute(&trueX.GetBlock(0));
    ParSaturationSolver(p, S, velocity, ti, p.dt_global, visit_local);

    if (p.vis_steps_global > 0 && ti % p.vis_steps_global == 0) {  // pressure per cell [REF src/run_parallel.cpp:189-204]
      P.Distribute(&trueX.GetBlock(1));
      output_cells(p, P.GetData(), ld.n_p, npl, cell_start, myid, num_procs, ti, "pressure");
    }
    if (p.saturation_steps > 0 && ti % p.saturation_steps == 0)
      output_cells(p, S.GetData(), ld.n_s, nsl, cell_start, myid, num_procs, ti, "saturation");
    if (myid == 0) cout << "time step " << ti << " is done" << endl;

    // S.CheckFinite() + MPI_Allreduce(SUM) of the reference [REF src/run_parallel.cpp:220-230] (the S > 1.1 guard is
    // commented out there)
    int64_t n_bad = 0;
    double s_max = 0;
    MFEM_VERIFY(impes_guards(h, &n_bad, &s_max) == IMPES_OK, "impes_guards failed");
    if (n_bad > 0) {
      valid_loop = 0;
      if (myid == 0) cout << "\n\n\tSaturation has nonfinite values. The process stops.\n" << endl;
    }
  }
  if (myid == 0) cout << "Time of the global time loop: " << seconds_since(global_time_loop) << " sec" << endl;

  impes_host_release(p);
  delete hdiv_coll;
  delete l2_coll;
  delete dg_coll;
  delete pmesh;
  if (myid == 0) cout << "TOTAL TIME: " << seconds_since(total_time) << " sec" << endl;
}
