#include "param.hpp"

#include <cstring>

#include "utilities.hpp"

DarcySolverParam::DarcySolverParam() : maxiter(1000), rtol(1e-6), atol(1e-10), print_level(0) {}

void DarcySolverParam::add_options(mfem::OptionsParser& args) {
  args.AddOption(&maxiter, "-darcy-maxiter", "--darcy-maxiter", "Darcy solver: max number of MINRES iterations");
  args.AddOption(&rtol, "-darcy-rtol", "--darcy-rtol", "Darcy solver: relative tolerance");
  args.AddOption(&atol, "-darcy-atol", "--darcy-atol", "Darcy solver: absolute tolerance");
  args.AddOption(&print_level, "-darcy-print-level", "--darcy-print-level", "Darcy solver: print level");
}

Well::Well() : radius(1.), height(1.) { center.x[0] = center.x[1] = center.x[2] = 0.; }

Param::Param()
    : spacedim(2), nx(100), ny(100), nz(100), sx(100.), sy(100.), sz(100.),
      K_array_x(nullptr), K_array_y(nullptr), K_array_z(nullptr), phi_array(nullptr),
      order_v(1), order_p(1), order_s(0), t_final(200), dt_global(100), dt_local(1),
      vis_steps_global(1), vis_steps_local(-1), saturation_steps(1),
      K_x(1.0), K_y(1.0), K_z(1.0), phi(1.0),
      K_file_x("no-file"), K_file_y("no-file"), K_file_z("no-file"), phi_file("no-file"),
      outdir("output"), extra(""), two_phase_flow(false), ode_solver_type(1),
      inflow(1.), outflow(-1.), saturation_source(1.), Sor(0.), Swc(0.), info(false),
      schur_prec(IMPES_SCHUR_SGS_LEX), device(0) {}

Param::~Param() {
  delete[] phi_array;
  delete[] K_array_z;
  delete[] K_array_y;
  delete[] K_array_x;
}

namespace {
double* load_cells(const char* file, double constant, int n) {
  double* a = new double[n];
  if (!std::strcmp(file, "no-file")) for (int i = 0; i < n; ++i) a[i] = constant;
  else read_binary(file, n, a);
  return a;
}
}  // namespace

// K / phi arrays: constant or from a raw binary file; porosity is floored at 1e-3 [REF src/param.cpp:78-113]
void Param::init_arrays() {
  const int n = get_n_cells();
  K_array_x = load_cells(K_file_x, K_x, n);
  K_array_y = load_cells(K_file_y, K_y, n);
  if (spacedim == 3) K_array_z = load_cells(K_file_z, K_z, n);
  phi_array = load_cells(phi_file, phi, n);
  for (int i = 0; i < n; ++i)
    if (phi_array[i] < 1e-3) phi_array[i] = 1e-3;
}

void Param::add_options(mfem::OptionsParser& args) {
  struct IntOpt { int* v; const char *s, *l, *d; };
  struct DblOpt { double* v; const char *s, *l, *d; };
  struct StrOpt { const char** v; const char *s, *l, *d; };
  const IntOpt grid_i[] = {
      {&spacedim, "-d", "--dim", "Space dimension (2 or 3)"},
      {&nx, "-nx", "--nx", "Cells along x"}, {&ny, "-ny", "--ny", "Cells along y"}, {&nz, "-nz", "--nz", "Cells along z"}};
  const DblOpt grid_d[] = {
      {&sx, "-sx", "--sx", "Domain size along x"}, {&sy, "-sy", "--sy", "Domain size along y"}, {&sz, "-sz", "--sz", "Domain size along z"}};
  const IntOpt order_i[] = {
      {&order_v, "-ov", "--orderv", "Order of the RT space (velocity)"},
      {&order_p, "-op", "--orderp", "Order of the L2 space (pressure)"},
      {&order_s, "-os", "--orders", "Order of the DG space (saturation)"}};
  const DblOpt time_d[] = {
      {&t_final, "-tf", "--t-final", "Final time (start time is 0)"},
      {&dt_global, "-dtg", "--time-step-global", "Time step of the pressure solves"},
      {&dt_local, "-dtl", "--time-step-local", "Time step of the saturation sub-steps"}};
  const IntOpt out_i[] = {
      {&vis_steps_global, "-vsg", "--vis-steps-global", "Visualize every n-th global step (<=0: never)"},
      {&vis_steps_local, "-vsl", "--vis-steps-local", "Visualize every n-th saturation sub-step (<=0: never)"},
      {&saturation_steps, "-ss", "--saturation-steps", "Write saturation as a binary file every n-th global step (<=0: never)"}};
  const DblOpt med_d[] = {
      {&K_x, "-Kx", "--Kx", "Constant permeability, x component"}, {&K_y, "-Ky", "--Ky", "Constant permeability, y component"},
      {&K_z, "-Kz", "--Kz", "Constant permeability, z component"}, {&phi, "-phi", "--phi", "Constant porosity"}};
  const StrOpt file_s[] = {
      {&K_file_x, "-K-file-x", "--K-file-x", "Binary file (float32) with per-cell permeability, x component"},
      {&K_file_y, "-K-file-y", "--K-file-y", "Binary file (float32) with per-cell permeability, y component"},
      {&K_file_z, "-K-file-z", "--K-file-z", "Binary file (float32) with per-cell permeability, z component"},
      {&phi_file, "-phi-file", "--phi-file", "Binary file (float32) with per-cell porosity"},
      {&outdir, "-outdir", "--output-directory", "Output directory"},
      {&extra, "-extra", "--extra", "Extra string in output file names"}};
  for (const IntOpt& o : grid_i) args.AddOption(o.v, o.s, o.l, o.d);
  for (const DblOpt& o : grid_d) args.AddOption(o.v, o.s, o.l, o.d);
  for (const IntOpt& o : order_i) args.AddOption(o.v, o.s, o.l, o.d);
  for (const DblOpt& o : time_d) args.AddOption(o.v, o.s, o.l, o.d);
  for (const IntOpt& o : out_i) args.AddOption(o.v, o.s, o.l, o.d);
  for (const DblOpt& o : med_d) args.AddOption(o.v, o.s, o.l, o.d);
  for (const StrOpt& o : file_s) args.AddOption(o.v, o.s, o.l, o.d);
  args.AddOption(&two_phase_flow, "-two-phase", "--two-phase-flow", "-single-phase", "--single-phase-flow",
                 "Two-phase flow (otherwise single-phase)");
  args.AddOption(&ode_solver_type, "-ode", "--ode-solver-type", "Saturation ODE solver (1=Forward Euler, 2=RK2, 3=RK3, 4=RK4, 6=RK6)");
  struct WellOpt { Well* w; const char *sx, *lx, *sy, *ly, *sz, *lz, *sr, *lr, *sh, *lh; };
  const WellOpt wells[] = {
      {&injection, "-inj-xcen", "--inj-center-x", "-inj-ycen", "--inj-center-y", "-inj-zcen", "--inj-center-z", "-inj-r", "--inj-radius", "-inj-h", "--inj-height"},
      {&production, "-pro-xcen", "--pro-center-x", "-pro-ycen", "--pro-center-y", "-pro-zcen", "--pro-center-z", "-pro-r", "--pro-radius", "-pro-h", "--pro-height"}};
  for (const WellOpt& w : wells) {
    args.AddOption(&w.w->center(0), w.sx, w.lx, "Well centre, x");
    args.AddOption(&w.w->center(1), w.sy, w.ly, "Well centre, y");
    args.AddOption(&w.w->center(2), w.sz, w.lz, "Well centre, z");
    args.AddOption(&w.w->radius, w.sr, w.lr, "Well radius");
    args.AddOption(&w.w->height, w.sh, w.lh, "Well height");
  }
  const DblOpt flow_d[] = {
      {&inflow, "-inflow", "--inflow", "Source value in the injection well"},
      {&outflow, "-outflow", "--outflow", "Source value in the production well"},
      {&saturation_source, "-sat-source", "--saturation-source", "Saturation source value"},
      {&Sor, "-Sor", "--Sor", "Irreducible oil saturation"}, {&Swc, "-Swc", "--Swc", "Connate water saturation"}};
  for (const DblOpt& o : flow_d) args.AddOption(o.v, o.s, o.l, o.d);
  darcy.add_options(args);
  args.AddOption(&info, "-info", "--info", "-no-info", "--no-info", "Print build information and exit");
  // extensions of the B200 build (absent from the reference's option table)
  args.AddOption(&schur_prec, "-schur-prec", "--schur-prec", "GPU Schur-block preconditioner: 0 = lexicographic symmetric Gauss-Seidel (serial reference), 1 = multigrid");
  args.AddOption(&device, "-device", "--device", "CUDA device ordinal");
}

std::string Param::get_info() const {
  return std::string("\nBuild                      : impes-b200 (sm_100a CUDA hot path behind the reference's solver interface)") +
         "\nC ABI version              : " + d2s(impes_abi_version()) +
         "\nMFEM / HYPRE / METIS       : not used (matrix-free GPU kernels; see DESIGN.md)\n";
}

double Param::Krw(double S) const {
  if (!two_phase_flow) return S;
  const double S_ = (S - Swc) / (1. - Sor - Swc);
  return S_ * S_;
}

double Param::Kro(double S) const {
  if (!two_phase_flow) return (1. - S);
  const double S_ = (S - Swc) / (1. - Sor - Swc);
  return (1. - S_) * (1. - S_);
}

impes_desc Param::to_desc() const {
  impes_desc d;
  impes_desc_defaults(&d);
  d.dim = spacedim; d.nx = nx; d.ny = ny; d.nz = nz;
  d.sx = sx; d.sy = sy; d.sz = sz;
  d.order_v = order_v; d.order_p = order_p; d.order_s = order_s;
  d.two_phase = two_phase_flow ? 1 : 0;
  d.ode_solver_type = ode_solver_type;
  d.Sor = Sor; d.Swc = Swc; d.mu_w = MU_W; d.mu_o = MU_O;
  const Well* w[2] = {&injection, &production};
  impes_well* o[2] = {&d.injection, &d.production};
  for (int i = 0; i < 2; ++i) {
    for (int c = 0; c < 3; ++c) o[i]->center[c] = w[i]->center(c);
    o[i]->radius = w[i]->radius; o[i]->height = w[i]->height;
  }
  d.inflow = inflow; d.outflow = outflow; d.saturation_source = saturation_source;
  d.darcy_maxiter = darcy.maxiter; d.darcy_rtol = darcy.rtol; d.darcy_atol = darcy.atol;
  d.darcy_print_level = darcy.print_level;
  d.schur_prec = schur_prec;
  d.substep_rule = IMPES_SUBSTEPS_SERIAL;
  d.device = device;
  return d;
}
