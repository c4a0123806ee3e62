// param.hpp -- the problem description the two solver calls read, with the reference's option surface
// (SURVEY Appendix D): same member names, same flags, same defaults [REF src/param.hpp:11-122]
// [REF src/param.cpp:8-13,26-30,35-68,115-170].  The MFEM Coefficient subclasses of the reference
// (CWVectorCoefficient, WellFunctionCoefficient, ...) are not reproduced: their arithmetic runs inside
// the CUDA kernels (csrc/kernels.cuh) and is described to them through impes_desc (to_desc()).
#pragma once
#include <string>

#include "../../include/impes_b200.h"
#include "config.hpp"

struct DarcySolverParam {
  int maxiter;
  double rtol;
  double atol;
  int print_level;
  DarcySolverParam();
  void add_options(mfem::OptionsParser& args);
};

// A well: box of half-width `radius` around `center` (2-D), or box in (x,z) extruded over
// [center.y, center.y + height] (3-D) [REF src/param.hpp:310-345].
struct WellCenter {
  double x[3];
  double& operator()(int i) { return x[i]; }
  const double& operator()(int i) const { return x[i]; }
};
struct Well {
  WellCenter center;
  double radius;
  double height;
  Well();
};

struct Param {
  int spacedim;
  int nx, ny, nz;
  double sx, sy, sz;
  double *K_array_x, *K_array_y, *K_array_z;  // one value per cell, x fastest
  double* phi_array;
  int order_v, order_p, order_s;
  double t_final, dt_global, dt_local;
  int vis_steps_global, vis_steps_local, saturation_steps;
  double K_x, K_y, K_z, phi;
  const char *K_file_x, *K_file_y, *K_file_z, *phi_file;
  const char* outdir;
  const char* extra;
  bool two_phase_flow;
  int ode_solver_type;
  Well injection, production;
  double inflow, outflow, saturation_source;
  double Sor, Swc;
  DarcySolverParam darcy;
  bool info;

  // B200 build only (not in the reference): which Schur-block preconditioner the GPU solve uses
  // (0 = the serial reference's lexicographic symmetric Gauss-Seidel, 1 = multigrid) and which GPU.
  int schur_prec;
  int device;

  Param();
  ~Param();
  void init_arrays();
  void add_options(mfem::OptionsParser& args);
  std::string get_info() const;

  int get_n_cells() const { return spacedim == 3 ? nx * ny * nz : nx * ny; }
  double get_cell_volume() const { return (spacedim == 3 ? sx * sy * sz : sx * sy) / get_n_cells(); }
  double Krw(double S) const;  // relative permeability of water [REF src/param.cpp:195-204]
  double Kro(double S) const;  // relative permeability of oil   [REF src/param.cpp:206-215]

  // Flatten into the C-ABI descriptor (include/impes_b200.h).
  impes_desc to_desc() const;

 private:
  Param(const Param&);
  Param& operator=(const Param&);
};
