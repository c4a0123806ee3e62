// mfem_compat.hpp -- the sliver of MFEM's interface that the IMPES solver calls touch, so that the
// reference-signature entry points (pressure_solver.hpp / saturation_solver.hpp) compile and run in an
// environment without MFEM (this image has none).  With -DIMPES_WITH_MFEM the real <mfem.hpp> is used
// instead and this file is not included.
//
// Only what crosses the solver boundary is modelled [REF src/pressure_solver.hpp:9-15]
// [REF src/saturation_solver.hpp:9-11] [REF src/run_serial.cpp:46-86]: Array<int>, Vector, BlockVector,
// a Cartesian Mesh, FiniteElementSpace (RT_k / L2_k / DG_k sizes), GridFunction and the two coefficient
// wrappers the drivers pass in.  Dof ordering in this shim is the library's native layout
// (include/impes_b200.h); MFEM's own permutation/sign convention is applied only in the real-MFEM adapter.
#pragma once
#include <cassert>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

namespace mfem {

template <class T>
class Array {
 public:
  Array() {}
  explicit Array(int n) : d_(n) {}
  int Size() const { return (int)d_.size(); }
  void SetSize(int n) { d_.resize(n); }
  T& operator[](int i) { return d_[i]; }
  const T& operator[](int i) const { return d_[i]; }
  T& Last() { return d_.back(); }
  const T& Last() const { return d_.back(); }
  T Max() const { T m = d_[0]; for (const T& v : d_) if (v > m) m = v; return m; }
  Array& operator=(const T& v) { for (T& x : d_) x = v; return *this; }
  void PartialSum() { for (size_t i = 1; i < d_.size(); ++i) d_[i] += d_[i - 1]; }
 private:
  std::vector<T> d_;
};

class Vector {
 public:
  Vector() : data_(nullptr), size_(0), own_(false) {}
  explicit Vector(int n) : data_(new double[n]), size_(n), own_(true) {}
  Vector(double* d, int n) : data_(d), size_(n), own_(false) {}
  Vector(const Vector& o) : data_(new double[o.size_]), size_(o.size_), own_(true) { std::memcpy(data_, o.data_, sizeof(double) * size_); }
  virtual ~Vector() { if (own_) delete[] data_; }
  void SetSize(int n) { if (own_) delete[] data_; data_ = new double[n]; size_ = n; own_ = true; }
  void SetDataAndSize(double* d, int n) { if (own_) delete[] data_; data_ = d; size_ = n; own_ = false; }
  int Size() const { return size_; }
  double* GetData() const { return data_; }
  double& operator()(int i) { return data_[i]; }
  const double& operator()(int i) const { return data_[i]; }
  Vector& operator=(double v) { for (int i = 0; i < size_; ++i) data_[i] = v; return *this; }
  Vector& operator=(const Vector& o) { assert(size_ == o.size_); std::memcpy(data_, o.data_, sizeof(double) * size_); return *this; }
  double Max() const { double m = data_[0]; for (int i = 1; i < size_; ++i) if (data_[i] > m) m = data_[i]; return m; }
  int CheckFinite() const { int bad = 0; for (int i = 0; i < size_; ++i) if (!(data_[i] - data_[i] == 0.0)) ++bad; return bad; }
 protected:
  double* data_;
  int size_;
  bool own_;
};

class BlockVector : public Vector {
 public:
  explicit BlockVector(const Array<int>& offsets) : Vector(offsets.Last()), off_(offsets) {
    blocks_.resize(offsets.Size() - 1);
    for (int b = 0; b + 1 < offsets.Size(); ++b) blocks_[b].SetDataAndSize(data_ + offsets[b], offsets[b + 1] - offsets[b]);
  }
  Vector& GetBlock(int b) { return blocks_[b]; }
  const Vector& GetBlock(int b) const { return blocks_[b]; }
  BlockVector& operator=(double v) { Vector::operator=(v); return *this; }
 private:
  Array<int> off_;
  std::vector<Vector> blocks_;
};

struct Element { enum Type { QUADRILATERAL = 3, HEXAHEDRON = 5 }; };

// Cartesian mesh, the only kind the reference builds [REF src/run_serial.cpp:46-53]
class Mesh {
 public:
  Mesh(int nx, int ny, Element::Type, int /*gen_edges*/, double sx, double sy) : dim_(2), nx_(nx), ny_(ny), nz_(1), sx_(sx), sy_(sy), sz_(1) { bdr_attributes.SetSize(4); for (int i = 0; i < 4; ++i) bdr_attributes[i] = i + 1; }
  Mesh(int nx, int ny, int nz, Element::Type, int /*gen_edges*/, double sx, double sy, double sz) : dim_(3), nx_(nx), ny_(ny), nz_(nz), sx_(sx), sy_(sy), sz_(sz) { bdr_attributes.SetSize(6); for (int i = 0; i < 6; ++i) bdr_attributes[i] = i + 1; }
  int Dimension() const { return dim_; }
  int GetNE() const { return nx_ * ny_ * nz_; }
  int nx() const { return nx_; } int ny() const { return ny_; } int nz() const { return nz_; }
  double sx() const { return sx_; } double sy() const { return sy_; } double sz() const { return sz_; }
  Array<int> bdr_attributes;
 private:
  int dim_, nx_, ny_, nz_;
  double sx_, sy_, sz_;
};

class FiniteElementCollection {
 public:
  enum Kind { RT, L2 };
  FiniteElementCollection(Kind k, int order, int dim) : kind(k), order(order), dim(dim) {}
  virtual ~FiniteElementCollection() {}
  Kind kind; int order, dim;
};
struct RT_FECollection : FiniteElementCollection { RT_FECollection(int p, int dim) : FiniteElementCollection(RT, p, dim) {} };
struct L2_FECollection : FiniteElementCollection { L2_FECollection(int p, int dim) : FiniteElementCollection(L2, p, dim) {} };
typedef L2_FECollection DG_FECollection;

class FiniteElementSpace {
 public:
  FiniteElementSpace(Mesh* m, const FiniteElementCollection* fec) : mesh_(m), fec_(fec) {
    const int k = fec->order, n[3] = {m->nx(), m->ny(), m->nz()}, dim = m->Dimension();
    if (fec->kind == FiniteElementCollection::L2) { long s = m->GetNE(); for (int e = 0; e < dim; ++e) s *= k + 1; size_ = (int)s; }
    else { long tot = 0; for (int d = 0; d < dim; ++d) { long s = 1; for (int e = 0; e < dim; ++e) s *= (e == d) ? n[e] * (k + 1) + 1 : n[e] * (k + 1); tot += s; } size_ = (int)tot; }
  }
  int GetVSize() const { return size_; }
  Mesh* GetMesh() const { return mesh_; }
  const FiniteElementCollection* FEColl() const { return fec_; }
 private:
  Mesh* mesh_; const FiniteElementCollection* fec_; int size_;
};

class GridFunction : public Vector {
 public:
  GridFunction() : fes_(nullptr) {}
  explicit GridFunction(FiniteElementSpace* f) : Vector(f->GetVSize()), fes_(f) {}
  void Update(FiniteElementSpace* f, Vector& v, int offset) { fes_ = f; SetDataAndSize(v.GetData() + offset, f->GetVSize()); }
  FiniteElementSpace* FESpace() const { return fes_; }
  GridFunction& operator=(double v) { Vector::operator=(v); return *this; }
 private:
  FiniteElementSpace* fes_;
};

class Coefficient { public: virtual ~Coefficient() {} };
class GridFunctionCoefficient : public Coefficient {
 public:
  explicit GridFunctionCoefficient(GridFunction* g) : gf_(g) {}
  GridFunction* GetGridFunction() const { return gf_; }
 private:
  GridFunction* gf_;
};
class VectorCoefficient { public: virtual ~VectorCoefficient() {} };
class VectorGridFunctionCoefficient : public VectorCoefficient {
 public:
  explicit VectorGridFunctionCoefficient(GridFunction* g) : gf_(g) {}
  GridFunction* GetGridFunction() const { return gf_; }
 private:
  GridFunction* gf_;
};

[[noreturn]] inline void mfem_error(const char* msg) { std::cerr << "\n\nMFEM abort: " << msg << std::endl; std::abort(); }

}  // namespace mfem

#define MFEM_ABORT(msg) ::mfem::mfem_error((std::string("") + msg).c_str())
#define MFEM_VERIFY(x, msg) do { if (!(x)) MFEM_ABORT(msg); } while (0)
