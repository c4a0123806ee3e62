// host_selftest.cpp -- CPU-only checks of the host layer (no GPU needed): Param defaults and option table,
// Krw/Kro, read_binary/write_binary round trip, descriptor flattening.  Prints one "ok <name>" per check.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <fstream>
#include <sstream>
#include <vector>

#include "dofmap.hpp"
#include "param.hpp"
#include "utilities.hpp"

#define CHECK(c) do { if (!(c)) { std::cout << "FAIL " << #c << " (" << __FILE__ << ":" << __LINE__ << ")" << std::endl; return 1; } } while (0)

int main(int argc, char** argv) {
  const std::string tmp = argc > 1 ? argv[1] : "/tmp";
  {  // defaults [REF src/param.cpp:8-13,26-30,35-68]
    Param p;
    CHECK(p.spacedim == 2 && p.nx == 100 && p.ny == 100 && p.nz == 100);
    CHECK(p.order_v == 1 && p.order_p == 1 && p.order_s == 0);
    CHECK(p.t_final == 200 && p.dt_global == 100 && p.dt_local == 1);
    CHECK(p.darcy.maxiter == 1000 && p.darcy.rtol == 1e-6 && p.darcy.atol == 1e-10 && p.darcy.print_level == 0);
    CHECK(!p.two_phase_flow && p.ode_solver_type == 1 && p.inflow == 1. && p.outflow == -1. && p.saturation_source == 1.);
    CHECK(p.injection.radius == 1. && p.production.height == 1. && p.injection.center(2) == 0.);
    CHECK(!std::strcmp(p.K_file_x, "no-file") && !std::strcmp(p.outdir, "output") && !std::strcmp(p.extra, ""));
    std::cout << "ok defaults" << std::endl;
  }
  {  // option surface
    const char* av[] = {"prog", "-d", "3", "-nx", "7", "--ny", "5", "-nz", "3", "-sx", "14", "-ov", "0", "-op", "0", "-os", "0",
                        "-two-phase", "-Sor", "0.1", "--Swc", "0.2", "-darcy-rtol", "1e-8", "-darcy-maxiter", "77",
                        "-inj-xcen", "1.5", "-pro-h", "9", "-K-file-x", "kx.bin", "-ode", "3", "-dtl", "0.25", "-no-info"};
    int ac = sizeof(av) / sizeof(av[0]);
    Param p;
    mfem::OptionsParser args(ac, const_cast<char**>(av));
    p.add_options(args);
    args.Parse();
    CHECK(args.Good());
    CHECK(p.spacedim == 3 && p.nx == 7 && p.ny == 5 && p.nz == 3 && p.sx == 14. && p.order_v == 0);
    CHECK(p.two_phase_flow && p.Sor == 0.1 && p.Swc == 0.2 && p.darcy.rtol == 1e-8 && p.darcy.maxiter == 77);
    CHECK(p.injection.center(0) == 1.5 && p.production.height == 9. && !std::strcmp(p.K_file_x, "kx.bin"));
    CHECK(p.ode_solver_type == 3 && p.dt_local == 0.25 && !p.info);
    CHECK(p.get_n_cells() == 105 && std::fabs(p.get_cell_volume() - 14. * 100. * 100. / 105.) < 1e-9);
    // rel-perm [REF src/param.cpp:195-215]
    const double S_ = (0.5 - 0.2) / (1. - 0.1 - 0.2);
    CHECK(p.Krw(0.5) == S_ * S_ && p.Kro(0.5) == (1. - S_) * (1. - S_));
    const impes_desc d = p.to_desc();
    CHECK(d.dim == 3 && d.nx == 7 && d.two_phase == 1 && d.Sor == 0.1 && d.darcy_maxiter == 77 && d.mu_w == 1.0);
    CHECK(d.injection.center[0] == 1.5 && d.production.height == 9. && d.ode_solver_type == 3);
    const char* bad[] = {"prog", "-nx", "7", "-bogus"};
    Param q;
    mfem::OptionsParser a2(4, const_cast<char**>(bad));
    q.add_options(a2);
    a2.Parse();
    CHECK(!a2.Good());
    Param s;
    CHECK(s.Krw(0.3) == 0.3 && s.Kro(0.3) == 0.7);  // single-phase
    std::cout << "ok options" << std::endl;
  }
  {  // raw binary formats [REF src/utilities.cpp:21-74]
    const int n = 11;
    std::vector<double> a(n), b(n), c(n);
    for (int i = 0; i < n; ++i) a[i] = 0.1 * i + 1.0 / 3.0;
    const std::string f32 = tmp + "/impes_selftest_f32.bin", f64 = tmp + "/impes_selftest_f64.bin";
    write_binary(f32.c_str(), n, a.data());  // always float32
    read_binary(f32.c_str(), n, b.data());
    for (int i = 0; i < n; ++i) CHECK(b[i] == (double)(float)a[i]);
    FILE* fp = std::fopen(f64.c_str(), "wb");
    CHECK(fp && std::fwrite(a.data(), sizeof(double), n, fp) == (size_t)n);
    std::fclose(fp);
    read_binary(f64.c_str(), n, c.data());  // float64 auto-detected
    for (int i = 0; i < n; ++i) CHECK(c[i] == a[i]);
    CHECK(d2s(7, 0, 0, 0, 6) == "000007" && d2s(3) == "3");
    std::remove(f32.c_str());
    std::remove(f64.c_str());
    // array helpers [REF src/utilities.cpp:14-122]
    const double v[5] = {0.5, -2.0, 3.25, 1.0, 3.0};
    double lo, hi, le[5], ge[5];
    get_minmax(v, 5, lo, hi);
    CHECK(lo == -2.0 && hi == 3.25 && get_max(v, 5) == 3.25 && std::fabs(to_radians(180.0) - 3.141592653589793) < 1e-15);
    get_LE(v, 5, 1.0, le);
    get_GE(v, 5, 1.0, ge);
    const double want_le[5] = {0.5, -2.0, 1.0, 1.0, 1.0}, want_ge[5] = {1.0, 1.0, 3.25, 1.0, 3.0};
    for (int i = 0; i < 5; ++i) CHECK(le[i] == want_le[i] && ge[i] == want_ge[i]);
    std::cout << "ok binary_io" << std::endl;
  }
  {  // .vts writers [REF src/utilities.cpp:126-472]: the 2-D per-cell file byte for byte, the shapes of the others
    auto slurp = [](const std::string& f) { std::ifstream in(f); std::stringstream ss; ss << in.rdbuf(); return ss.str(); };
    auto count = [](const std::string& s, const std::string& w) { size_t n = 0; for (size_t p = s.find(w); p != std::string::npos; p = s.find(w, p + 1)) ++n; return n; };
    const std::string f = tmp + "/impes_selftest.vts";
    mfem::Vector c2(2); c2(0) = 0.25; c2(1) = 1.5;
    write_vts_scalar_cells(f, "saturation", 4.0, 3.0, 2, 1, c2);
    const std::string want =
        "<?xml version=\"1.0\"?>\n<VTKFile type=\"StructuredGrid\" version=\"0.1\">\n"
        "  <StructuredGrid WholeExtent=\"1 3 1 2 1 1\">\n    <Piece Extent=\"1 3 1 2 1 1\">\n"
        "      <CellData Scalars=\"saturation\">\n"
        "        <DataArray type=\"Float64\" Name=\"saturation\" format=\"ascii\" NumberOfComponents=\"1\">\n"
        "0.25 1.5 \n        </DataArray>\n      </CellData>\n      <Points>\n"
        "        <DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n"
        "0 0 0.0 2 0 0.0 4 0 0.0 0 3 0.0 2 3 0.0 4 3 0.0 \n"
        "        </DataArray>\n      </Points>\n    </Piece>\n  </StructuredGrid>\n</VTKFile>\n";
    CHECK(slurp(f) == want);
    mfem::Vector v3(3 * 3 * 2), w3(3 * 3 * 2), z3(3 * 3 * 2), c3(2 * 2 * 1);
    for (int i = 0; i < 18; ++i) { v3(i) = 3.0; w3(i) = 4.0; z3(i) = 12.0; }
    c3 = 0.5;
    write_vts_vector(f, "velocity", 1.0, 1.0, 1.0, 2, 2, 1, v3, w3, z3);
    std::string t = slurp(f);
    CHECK(t.find("WholeExtent=\"1 3 1 3 1 2\"") != std::string::npos && t.find("Name=\"velocity_scalar_mag\"") != std::string::npos);
    CHECK(count(t, "3 4 12 ") == 18 && count(t, "13 ") == 18);   // components, then |v| = 13 at every vertex
    write_vts_scalar_cells(f, "pressure", 1.0, 1.0, 1.0, 2, 2, 1, c3);
    t = slurp(f);
    CHECK(t.find("<CellData Scalars=\"pressure\">") != std::string::npos && count(t, "0.5 ") >= 4);
    mfem::Vector s2(6); s2 = 2.0;
    write_vts_scalar(f, "saturation", 4.0, 3.0, 2, 1, s2);   // 2-D vertex files: y extent in the third slot
    t = slurp(f);
    CHECK(t.find("WholeExtent=\"1 3 1 1 1 2\"") != std::string::npos && t.find("<PointData Scalars=\"saturation\">") != std::string::npos);
    write_vts_vector(f, "velocity", 4.0, 3.0, 2, 1, s2, s2);
    t = slurp(f);
    CHECK(count(t, "2 2 0.0 ") == 6 && t.find("_scalar_mag") == std::string::npos);
    std::remove(f.c_str());
    std::cout << "ok vts" << std::endl;
  }
  {  // init_arrays: constants, files, porosity floor [REF src/param.cpp:78-113]
    Param p;
    p.nx = 4; p.ny = 3; p.K_x = 2.5; p.phi = 1e-5;
    p.init_arrays();
    CHECK(p.K_array_x[11] == 2.5 && p.K_array_y[0] == 1.0 && p.K_array_z == nullptr && p.phi_array[5] == 1e-3);
    std::cout << "ok init_arrays" << std::endl;
  }
  {  // MFEM numbering <-> native layout, identified geometrically from the element interface (dofmap.hpp): the shim
     // space is scrambled (random permutation, random orientation of the RT dofs, outward-pointing lower-face bases)
     // and the map has to undo exactly that
    for (int dim = 2; dim <= 3; ++dim)
      for (int order = 0; order <= 1; ++order)
        for (int rt = 0; rt <= 1; ++rt) {
          const int n[3] = {5, 4, dim == 3 ? 3 : 1};
          const double size[3] = {10.0, 6.0, 4.5};
          mfem::Mesh* mesh = dim == 2 ? new mfem::Mesh(n[0], n[1], mfem::Element::QUADRILATERAL, 1, size[0], size[1])
                                      : new mfem::Mesh(n[0], n[1], n[2], mfem::Element::HEXAHEDRON, 1, size[0], size[1], size[2]);
          mfem::FiniteElementCollection* fec = rt ? (mfem::FiniteElementCollection*)new mfem::RT_FECollection(order, dim)
                                                   : (mfem::FiniteElementCollection*)new mfem::L2_FECollection(order, dim);
          mfem::FiniteElementSpace fes(mesh, fec);
          const DofMap id = build_dofmap(fes, rt, order, dim, n, size);
          CHECK(id.identity);  // unscrambled shim space = native layout
          fes.SetScramble(17 + dim + 2 * order + rt);
          const DofMap m = build_dofmap(fes, rt, order, dim, n, size);
          CHECK(!m.identity);
          const int nd = fes.GetVSize();
          std::vector<double> nat(nd), own(nd), back(nd), again(nd);
          for (int i = 0; i < nd; ++i) nat[i] = std::sin(0.37 * i) + 2.0;
          fes.FromNative(nat.data(), own.data());   // the same field in the scrambled numbering
          m.to_native(own.data(), back.data());
          for (int i = 0; i < nd; ++i) CHECK(back[i] == nat[i]);
          m.from_native(nat.data(), again.data());
          for (int i = 0; i < nd; ++i) CHECK(again[i] == own[i]);
          delete fec;
          delete mesh;
        }
    std::cout << "ok dofmap" << std::endl;
  }
  return 0;
}
