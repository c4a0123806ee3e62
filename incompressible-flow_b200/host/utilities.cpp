#include "utilities.hpp"

#include <cmath>
#include <cstdio>
#include <sstream>
#include <string>
#include <fstream>
#include <vector>

#include "config.hpp"
#include "mfem_compat.hpp"

// Same contract as the reference: the element size is inferred from file length / n_values; float32
// values are widened to double one by one [REF src/utilities.cpp:43-51].  Unlike the reference, float64
// files are accepted too (the reference's own gcount() check rejects them) and files larger
// than 2 GiB do not overflow.
void read_binary(const char* filename, int n_values, double* values) {
  std::ifstream in(filename, std::ios::binary | std::ios::ate);
  MFEM_VERIFY(in, "File '" + std::string(filename) + "' can't be opened");
  const long long length = (long long)in.tellg();
  MFEM_VERIFY(n_values > 0 && length % n_values == 0,
              "The number of bytes in the file '" + std::string(filename) +
                  "' is not divisible by the number of elements " + d2s(n_values));
  const long long size_value = length / n_values;
  in.seekg(0, std::ios::beg);
  if (size_value == (long long)sizeof(double)) {
    in.read(reinterpret_cast<char*>(values), (std::streamsize)n_values * sizeof(double));
    MFEM_VERIFY(in.gcount() == (std::streamsize)n_values * (std::streamsize)sizeof(double),
                "The number of successfully read elements is different from the expected one");
  } else if (size_value == (long long)sizeof(float)) {
    std::vector<float> buf(n_values);
    in.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)n_values * sizeof(float));
    MFEM_VERIFY(in.gcount() == (std::streamsize)n_values * (std::streamsize)sizeof(float),
                "The number of successfully read elements is different from the expected one");
    for (int i = 0; i < n_values; ++i) values[i] = buf[i];
  } else {
    MFEM_ABORT("Unknown size of an element (" + d2s(size_value) + ") in bytes. Expected one is either sizeof(float) = " +
               d2s(sizeof(float)) + ", or sizeof(double) = " + d2s(sizeof(double)));
  }
}

void write_binary(const char* filename, int n_values, const double* values) {
  std::ofstream out(filename, std::ios::binary);
  MFEM_VERIFY(out, "File '" + std::string(filename) + "' can't be opened");
  std::vector<float> buf(n_values);
  for (int i = 0; i < n_values; ++i) buf[i] = (float)values[i];
  out.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)n_values * sizeof(float));
}

double to_radians(double x) { return x * 3.14159265358979323846 / 180.0; }
double get_max(const double* a, int n_elements) {
  double lo, hi;
  get_minmax(a, n_elements, lo, hi);
  return hi;
}
void get_minmax(const double* a, int n_elements, double& min_val, double& max_val) {
  MFEM_VERIFY(n_elements > 0, "empty array");
  min_val = max_val = a[0];
  for (int i = 1; i < n_elements; ++i) {
    if (a[i] < min_val) min_val = a[i];
    if (a[i] > max_val) max_val = a[i];
  }
}
// one clip routine behind the reference's pair: keep the values on the `keep_below` side of the bound
static void clip(const double* a, int n, double bound, bool keep_below, double* out) {
  for (int i = 0; i < n; ++i) out[i] = ((a[i] <= bound) == keep_below || a[i] == bound) ? a[i] : bound;
}
void get_LE(const double* a, int n_elements, double LE_value, double* a_LE) { clip(a, n_elements, LE_value, true, a_LE); }
void get_GE(const double* a, int n_elements, double GE_value, double* a_GE) { clip(a, n_elements, GE_value, false, a_GE); }

// ---- .vts ------------------------------------------------------------------------------------------------
// One table-driven emitter behind the six reference signatures.  A file = header, one data section (PointData
// for vertex fields, CellData for per-cell fields), the vertex coordinates (the last vertex of every direction is
// put exactly on the domain boundary, as the reference does), footer.  Numbers go through the stream's default
// formatting like the reference's `out << value << " "`.
namespace {
struct VtsGrid {
  double s[3];
  int n[3];   // cells per direction; n[2] == 0 marks a 2-D grid (one vertex layer, z printed as the literal "0.0")
  bool y_third = false;  // 2-D vertex files of the reference put the y extent in the third slot ("1 nx+1 1 1 1 ny+1")
  int verts(int d) const { return n[d] + 1; }
  long n_verts() const { return (long)verts(0) * verts(1) * verts(2); }
  long n_cells() const { return (long)n[0] * n[1] * (n[2] ? n[2] : 1); }
};
struct VtsArray {
  std::string name;
  int ncomp;
  const double* comp[3];   // an absent component is written as the literal "0.0" (2-D vectors)
  bool magnitude;          // write sqrt(sum comp^2) instead of the components
};

void vts_array(std::ofstream& out, const VtsArray& a, long count) {
  out << "        <DataArray type=\"Float64\" Name=\"" << a.name << "\" format=\"ascii\" NumberOfComponents=\""
      << (a.magnitude ? 1 : a.ncomp) << "\">\n";
  for (long i = 0; i < count; ++i) {
    if (a.magnitude) {
      double m = 0.0;
      for (int c = 0; c < a.ncomp; ++c) m += a.comp[c] ? a.comp[c][i] * a.comp[c][i] : 0.0;
      out << std::sqrt(m) << " ";
    } else {
      for (int c = 0; c < a.ncomp; ++c) {
        if (a.comp[c]) out << a.comp[c][i] << " ";
        else out << "0.0 ";
      }
    }
  }
  out << "\n        </DataArray>\n";
}

void vts_write(const std::string& filename, const VtsGrid& g, bool cell_data, const std::string& attrs,
               const std::vector<VtsArray>& arrays) {
  std::ofstream out(filename.c_str());
  MFEM_VERIFY(out, "File '" + filename + "' can't be opened");
  std::ostringstream ext;
  if (g.y_third) ext << "1 " << g.verts(0) << " 1 1 1 " << g.verts(1);
  else ext << "1 " << g.verts(0) << " 1 " << g.verts(1) << " 1 " << g.verts(2);
  const char* section = cell_data ? "CellData" : "PointData";
  out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"StructuredGrid\" version=\"0.1\">\n"
      << "  <StructuredGrid WholeExtent=\"" << ext.str() << "\">\n    <Piece Extent=\"" << ext.str() << "\">\n"
      << "      <" << section << " " << attrs << ">\n";
  for (const VtsArray& a : arrays) vts_array(out, a, cell_data ? g.n_cells() : g.n_verts());
  out << "      </" << section << ">\n      <Points>\n"
      << "        <DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
  double h[3];
  for (int d = 0; d < 3; ++d) h[d] = g.n[d] ? g.s[d] / g.n[d] : 0.0;
  for (int iz = 0; iz < g.verts(2); ++iz)
    for (int iy = 0; iy < g.verts(1); ++iy)
      for (int ix = 0; ix < g.verts(0); ++ix) {
        const int iv[3] = {ix, iy, iz};
        for (int d = 0; d < 3; ++d) {
          if (g.n[d] == 0) out << "0.0 ";
          else out << (iv[d] == g.n[d] ? g.s[d] : iv[d] * h[d]) << " ";
        }
      }
  out << "\n        </DataArray>\n      </Points>\n    </Piece>\n  </StructuredGrid>\n</VTKFile>\n";
}

void vts_vector(const std::string& filename, const std::string& solname, const VtsGrid& g, const double* x, const double* y,
                const double* z) {
  if (!g.n[2]) {  // the reference's 2-D vector file has no magnitude array
    vts_write(filename, g, false, "Vectors=\"" + solname + "\"", {VtsArray{solname, 3, {x, y, z}, false}});
    return;
  }
  const std::string mag = solname + "_scalar_mag";
  vts_write(filename, g, false, "Vectors=\"" + solname + "\" Scalars=\"" + mag + "\"",
            {VtsArray{solname, 3, {x, y, z}, false}, VtsArray{mag, 3, {x, y, z}, true}});
}
void vts_scalar(const std::string& filename, const std::string& solname, const VtsGrid& g, bool cells, const mfem::Vector& sol) {
  MFEM_VERIFY((long)sol.Size() == (cells ? g.n_cells() : g.n_verts()), "Dimension mismatch");
  vts_write(filename, g, cells, "Scalars=\"" + solname + "\"", {VtsArray{solname, 1, {sol.GetData(), nullptr, nullptr}, false}});
}
}  // namespace

void write_vts_vector(const std::string& filename, const std::string& solname, double sx, double sy, double sz, int nx,
                      int ny, int nz, const mfem::Vector& sol_x, const mfem::Vector& sol_y, const mfem::Vector& sol_z) {
  vts_vector(filename, solname, VtsGrid{{sx, sy, sz}, {nx, ny, nz}, false}, sol_x.GetData(), sol_y.GetData(), sol_z.GetData());
}
void write_vts_vector(const std::string& filename, const std::string& solname, double sx, double sy, int nx, int ny,
                      const mfem::Vector& sol_x, const mfem::Vector& sol_y) {
  vts_vector(filename, solname, VtsGrid{{sx, sy, 0.0}, {nx, ny, 0}, true}, sol_x.GetData(), sol_y.GetData(), nullptr);
}
void write_vts_scalar(const std::string& filename, const std::string& solname, double sx, double sy, double sz, int nx,
                      int ny, int nz, const mfem::Vector& sol) {
  vts_scalar(filename, solname, VtsGrid{{sx, sy, sz}, {nx, ny, nz}, false}, false, sol);
}
void write_vts_scalar(const std::string& filename, const std::string& solname, double sx, double sy, int nx, int ny,
                      const mfem::Vector& sol) {
  vts_scalar(filename, solname, VtsGrid{{sx, sy, 0.0}, {nx, ny, 0}, true}, false, sol);
}
void write_vts_scalar_cells(const std::string& filename, const std::string& solname, double sx, double sy, double sz,
                            int nx, int ny, int nz, const mfem::Vector& sol) {
  vts_scalar(filename, solname, VtsGrid{{sx, sy, sz}, {nx, ny, nz}, false}, true, sol);
}
void write_vts_scalar_cells(const std::string& filename, const std::string& solname, double sx, double sy, int nx,
                            int ny, const mfem::Vector& sol) {
  vts_scalar(filename, solname, VtsGrid{{sx, sy, 0.0}, {nx, ny, 0}, false}, true, sol);
}
