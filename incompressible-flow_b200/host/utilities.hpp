// utilities.hpp -- the input/output formats on either side of the hot path:
// raw binary cell arrays.  read_binary auto-detects float32 / float64 from the file length
// [REF src/utilities.cpp:21-58]; write_binary always writes float32 [REF src/utilities.cpp:62-74].
#pragma once
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>

void read_binary(const char* filename, int n_values, double* values);
void write_binary(const char* filename, int n_values, const double* values);

// small array helpers of the reference [REF src/utilities.hpp:64-94]: degrees -> radians, maximum, minimum and maximum,
// clipping from above (get_LE: min(a, bound)) and from below (get_GE: max(a, bound))
double to_radians(double x);
double get_max(const double* a, int n_elements);
void get_minmax(const double* a, int n_elements, double& min_val, double& max_val);
void get_LE(const double* a, int n_elements, double LE_value, double* a_LE);
void get_GE(const double* a, int n_elements, double GE_value, double* a_GE);

// VTK StructuredGrid (.vts, ascii) writers of the reference, same signatures and file layout
// [REF src/utilities.hpp:109-139] [REF src/utilities.cpp:126-472]: vertex-based vector / scalar fields
// ((nx+1)(ny+1)(nz+1) values, x fastest) and per-cell scalars (nx ny nz values).  2-D files have extent 1 in z.
namespace mfem { class Vector; }
void write_vts_vector(const std::string& filename, const std::string& solname, double sx, double sy, double sz, int nx,
                      int ny, int nz, const mfem::Vector& sol_x, const mfem::Vector& sol_y, const mfem::Vector& sol_z);
void write_vts_vector(const std::string& filename, const std::string& solname, double sx, double sy, int nx, int ny,
                      const mfem::Vector& sol_x, const mfem::Vector& sol_y);
void write_vts_scalar(const std::string& filename, const std::string& solname, double sx, double sy, double sz, int nx,
                      int ny, int nz, const mfem::Vector& sol);
void write_vts_scalar(const std::string& filename, const std::string& solname, double sx, double sy, int nx, int ny,
                      const mfem::Vector& sol);
void write_vts_scalar_cells(const std::string& filename, const std::string& solname, double sx, double sy, double sz,
                            int nx, int ny, int nz, const mfem::Vector& sol);
void write_vts_scalar_cells(const std::string& filename, const std::string& solname, double sx, double sy, int nx,
                            int ny, const mfem::Vector& sol);

// number -> string with optional zero padding (file names "..._000003.bin") [REF src/utilities.hpp:29-62]
template <typename T>
inline std::string d2s(T data, bool scientific = false, int precision = 0, bool noperiod = false, int width = 0) {
  std::ostringstream o;
  if (scientific) o.setf(std::ios::scientific);
  if (precision > 0) o.precision(precision);
  if (width > 0) o << std::setfill('0') << std::setw(width);
  if (!(o << data)) throw std::runtime_error("Bad conversion of data to string!");
  std::string res = o.str();
  if (!scientific && noperiod) {
    const std::string::size_type pos = res.find('.');
    if (pos != std::string::npos) res.erase(pos, 1);
    if (width > 0 && (int)res.size() < width) res.insert(0, width - res.size(), '0');
  }
  return res;
}
