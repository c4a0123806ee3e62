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
#include <algorithm>
#include <cassert>
#include <cstdlib>
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

// ---- the part of MFEM's element-level interface that host/dofmap.cpp reads to identify dofs geometrically ----
struct IntegrationPoint { double x = 0, y = 0, z = 0; };
class IntegrationRule {
 public:
  int GetNPoints() const { return (int)p_.size(); }
  const IntegrationPoint& IntPoint(int i) const { return p_[i]; }
  std::vector<IntegrationPoint> p_;
};
class DenseMatrix {
 public:
  void SetSize(int r, int c) { r_ = r; c_ = c; d_.assign((size_t)r * c, 0.0); }
  double& operator()(int i, int j) { return d_[i + (size_t)r_ * j]; }
  const double& operator()(int i, int j) const { return d_[i + (size_t)r_ * j]; }
 private:
  int r_ = 0, c_ = 0;
  std::vector<double> d_;
};
// axis-aligned cell [lo, lo + h]
class ElementTransformation {
 public:
  void SetIntPoint(const IntegrationPoint* ip) { ip_ = ip; }
  const IntegrationPoint& GetIntPoint() const { return *ip_; }
  void Transform(const IntegrationPoint& ip, Vector& x) const {
    const double r[3] = {ip.x, ip.y, ip.z};
    for (int d = 0; d < dim; ++d) x(d) = lo[d] + h[d] * r[d];
  }
  int dim = 2;
  double lo[3] = {0, 0, 0}, h[3] = {1, 1, 1};
 private:
  const IntegrationPoint* ip_ = nullptr;
};

// Nodal RT_k / L2_k (k = 0, 1) reference element in the shim's own local dof order (deliberately NOT the native
// one: direction-major, open nodes fastest, basis functions on the lower faces pointing OUTWARD like MFEM's), so that
// host/dofmap.cpp is exercised on something it has to work out from node positions and shape functions.
class FiniteElement {
 public:
  FiniteElement(FiniteElementCollection::Kind kind, int k, int dim) : kind_(kind), k_(k), dim_(dim) {
    const double g = 0.5 / 1.7320508075688772;
    double cl[3], op[2];
    const int nc = k + 2, no = k + 1;
    if (k == 0) { cl[0] = 0; cl[1] = 1; op[0] = 0.5; } else { cl[0] = 0; cl[1] = 0.5; cl[2] = 1; op[0] = 0.5 - g; op[1] = 0.5 + g; }
    for (int i = 0; i < nc; ++i) cl_[i] = cl[i];
    for (int i = 0; i < no; ++i) op_[i] = op[i];
    if (kind == FiniteElementCollection::L2) {
      const int n3 = dim == 3 ? no : 1;
      for (int c = 0; c < n3; ++c) for (int b = 0; b < no; ++b) for (int a = 0; a < no; ++a) {
        IntegrationPoint ip; ip.x = op[a]; ip.y = op[b]; ip.z = dim == 3 ? op[c] : 0.0;
        nodes_.p_.push_back(ip); dir_.push_back(-1); idx_.push_back({a, b, c});
      }
    } else {
      for (int d = 0; d < dim; ++d) {
        const int e = (d + 1) % dim, f = (d + 2) % dim;  // transverse directions (f unused in 2-D)
        const int nf = dim == 3 ? no : 1;
        for (int n = 0; n < nc; ++n) for (int tf = 0; tf < nf; ++tf) for (int te = 0; te < no; ++te) {
          double r[3] = {0, 0, 0};
          r[d] = cl[n]; r[e] = op[te]; if (dim == 3) r[f] = op[tf];
          IntegrationPoint ip; ip.x = r[0]; ip.y = r[1]; ip.z = r[2];
          nodes_.p_.push_back(ip); dir_.push_back(d);
          int id[3] = {0, 0, 0}; id[d] = n; id[e] = te; if (dim == 3) id[f] = tf;
          idx_.push_back({id[0], id[1], id[2]});
        }
      }
    }
  }
  int GetDof() const { return nodes_.GetNPoints(); }
  const IntegrationRule& GetNodes() const { return nodes_; }
  // physical shape functions at Trans.GetIntPoint(): row k = basis function k (contravariant Piola on a box)
  void CalcVShape(ElementTransformation& T, DenseMatrix& shape) const {
    const IntegrationPoint& ip = T.GetIntPoint();
    const double r[3] = {ip.x, ip.y, ip.z};
    shape.SetSize(GetDof(), dim_);
    double vol = 1; for (int d = 0; d < dim_; ++d) vol *= T.h[d];
    for (int k = 0; k < GetDof(); ++k) {
      const int d = dir_[k];
      double v = 1.0;
      for (int e = 0; e < dim_; ++e) v *= (e == d) ? lag(cl_, k_ + 2, idx_[k].i[e], r[e]) : lag(op_, k_ + 1, idx_[k].i[e], r[e]);
      shape(k, d) = local_sign(k) * v * T.h[d] / vol;
    }
  }
  // shim-only accessors (the space needs them to number dofs); not part of MFEM's interface
  int dof_dir(int k) const { return dir_[k]; }
  const int* dof_idx(int k) const { return idx_[k].i; }
  double local_sign(int k) const { return (kind_ == FiniteElementCollection::RT && idx_[k].i[dir_[k]] == 0) ? -1.0 : 1.0; }
 private:
  struct I3 { int i[3]; };
  static double lag(const double* nodes, int nn, int a, double x) {
    double v = 1.0;
    for (int m = 0; m < nn; ++m) if (m != a) v *= (x - nodes[m]) / (nodes[a] - nodes[m]);
    return v;
  }
  FiniteElementCollection::Kind kind_; int k_, dim_;
  double cl_[3], op_[2];
  IntegrationRule nodes_;
  std::vector<int> dir_;
  std::vector<I3> idx_;
};

class FiniteElementSpace {
 public:
  FiniteElementSpace(Mesh* m, const FiniteElementCollection* fec) : mesh_(m), fec_(fec), fe_(fec->kind, fec->order, m->Dimension()) {
    const int k = fec->order, n[3] = {m->nx(), m->ny(), m->nz()}, dim = m->Dimension();
    if (fec->kind == FiniteElementCollection::L2) { long s = m->GetNE(); for (int e = 0; e < dim; ++e) s *= k + 1; size_ = (int)s; }
    else { long tot = 0; for (int d = 0; d < dim; ++d) { long s = 1; for (int e = 0; e < dim; ++e) s *= (e == d) ? n[e] * (k + 1) + 1 : n[e] * (k + 1); tot += s; } size_ = (int)tot; }
    if (const char* e = std::getenv("IMPES_COMPAT_SCRAMBLE")) SetScramble((unsigned)std::atoi(e));
  }
  virtual ~FiniteElementSpace() {}
  virtual int GetVSize() const { return size_; }
  int GetNE() const { return mesh_->GetNE(); }
  Mesh* GetMesh() const { return mesh_; }
  const FiniteElementCollection* FEColl() const { return fec_; }
  const FiniteElement* GetFE(int) const { return &fe_; }
  ElementTransformation* GetElementTransformation(int e) {
    const int n[3] = {mesh_->nx(), mesh_->ny(), mesh_->nz()};
    const double s[3] = {mesh_->sx(), mesh_->sy(), mesh_->sz()};
    const int c[3] = {e % n[0], (e / n[0]) % n[1], e / (n[0] * n[1])};
    tr_.dim = mesh_->Dimension();
    for (int d = 0; d < tr_.dim; ++d) { tr_.h[d] = s[d] / n[d]; tr_.lo[d] = c[d] * tr_.h[d]; }
    return &tr_;
  }
  // MFEM's convention: a negative entry -1-j means global dof j seen with the opposite orientation
  void GetElementDofs(int e, Array<int>& dofs) const {
    const int nd = fe_.GetDof();
    dofs.SetSize(nd);
    for (int k = 0; k < nd; ++k) {
      const long nat = native_index(e, k);
      const int g = perm_.empty() ? (int)nat : perm_[nat];
      const double sglob = gsign_.empty() ? 1.0 : (double)gsign_[nat];
      dofs[k] = (sglob == fe_.local_sign(k)) ? g : -1 - g;
    }
  }
  // Numbering of this shim space: seed 0 = the library's native layout with every dof along +d (what the
  // run_serial-style driver uses); otherwise a random permutation with random global orientations of the RT dofs,
  // which the adapter's dof map must undo (tests/test_host.py, host_selftest).
  void SetScramble(unsigned seed) {
    perm_.clear(); gsign_.clear();
    if (!seed) return;
    unsigned long long st = 0x9E3779B97F4A7C15ull * (seed + 1) + (fec_->kind == FiniteElementCollection::RT ? 7 : 11) + 1000003ull * fec_->order;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    perm_.resize(size_);
    for (int i = 0; i < size_; ++i) perm_[i] = i;
    for (int i = size_ - 1; i > 0; --i) std::swap(perm_[i], perm_[(int)(rnd() % (unsigned long long)(i + 1))]);
    gsign_.assign(size_, 1);
    if (fec_->kind == FiniteElementCollection::RT) for (int i = 0; i < size_; ++i) gsign_[i] = (rnd() & 1) ? 1 : -1;
  }
  bool Scrambled() const { return !perm_.empty(); }
  // shim-only: value conversion used by tests to build vectors in this space's numbering from native ones
  void FromNative(const double* nat, double* own) const {
    for (int i = 0; i < size_; ++i) own[perm_.empty() ? i : perm_[i]] = (gsign_.empty() ? 1.0 : gsign_[i]) * nat[i];
  }
 private:
  // native (include/impes_b200.h) index of local dof k of element e
  long native_index(int e, int k) const {
    const int kk = fec_->order, dim = mesh_->Dimension(), n[3] = {mesh_->nx(), mesh_->ny(), mesh_->nz()};
    const int c[3] = {e % n[0], (e / n[0]) % n[1], e / (n[0] * n[1])};
    const int* id = fe_.dof_idx(k);
    if (fec_->kind == FiniteElementCollection::L2) {
      const int no = kk + 1;
      long loc = id[0] + no * (id[1] + (dim == 3 ? no * id[2] : 0));
      long npl = 1; for (int d = 0; d < dim; ++d) npl *= no;
      return (long)e * npl + loc;
    }
    const int d = fe_.dof_dir(k);
    long off = 0;
    for (int dd = 0; dd < d; ++dd) { long sz = 1; for (int ee = 0; ee < dim; ++ee) sz *= (ee == dd) ? n[ee] * (kk + 1) + 1 : n[ee] * (kk + 1); off += sz; }
    long g[3] = {0, 0, 0}, ext[3] = {1, 1, 1};
    for (int ee = 0; ee < dim; ++ee) { g[ee] = (long)c[ee] * (kk + 1) + id[ee]; ext[ee] = (ee == d) ? n[ee] * (kk + 1) + 1 : n[ee] * (kk + 1); }
    return off + g[0] + ext[0] * (g[1] + (dim == 3 ? ext[1] * g[2] : 0));
  }
  Mesh* mesh_; const FiniteElementCollection* fec_; int size_;
  FiniteElement fe_;
  ElementTransformation tr_;
  std::vector<int> perm_;
  std::vector<signed char> gsign_;
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

// ---- the MPI-side types of the reference's signatures [REF src/pressure_solver.hpp:18-25]
// [REF src/saturation_solver.hpp:14-16] on ONE rank: in this shim a "parallel" object is its serial base (the
// multi-GPU decomposition lives inside the library: impes_*_local, host/run_parallel.cpp).
class ParMesh : public Mesh {
 public:
  explicit ParMesh(const Mesh& m, int rank = 0, int nranks = 1) : Mesh(m), rank_(rank), nranks_(nranks) {}
  int GetMyRank() const { return rank_; }
  int GetNRanks() const { return nranks_; }
 private:
  int rank_, nranks_;
};
class ParFiniteElementSpace : public FiniteElementSpace {
 public:
  ParFiniteElementSpace(ParMesh* m, const FiniteElementCollection* fec) : FiniteElementSpace(m, fec), pmesh_(m), local_(-1) {}
  // shim-only: the dofs this rank holds (its slab, impes_get_local_dims) when the mesh is split over several ranks
  ParFiniteElementSpace(ParMesh* m, const FiniteElementCollection* fec, int local_size) : FiniteElementSpace(m, fec), pmesh_(m), local_(local_size) {}
  int GetVSize() const override { return local_ >= 0 ? local_ : FiniteElementSpace::GetVSize(); }
  int GetTrueVSize() const { return GetVSize(); }
  int TrueVSize() const { return GetVSize(); }
  ParMesh* GetParMesh() const { return pmesh_; }
 private:
  ParMesh* pmesh_;
  int local_;
};
class ParGridFunction : public GridFunction {
 public:
  ParGridFunction() {}
  explicit ParGridFunction(ParFiniteElementSpace* f) : GridFunction(f) {}
  void Distribute(const Vector* tv) { std::memcpy(GetData(), tv->GetData(), sizeof(double) * Size()); }  // true dofs -> local dofs
  ParGridFunction& operator=(double v) { Vector::operator=(v); return *this; }
};
class VisItDataCollection {
 public:
  VisItDataCollection(const std::string& name, Mesh*) : name_(name) {}
  void SetPrefixPath(const std::string&) {}
  void RegisterField(const std::string&, GridFunction*) {}
  void SetCycle(int) {}
  void SetTime(double) {}
  void Save() {}
 private:
  std::string name_;
};

[[noreturn]] inline void mfem_error(const char* msg) { std::cerr << "\n\nMFEM abort: " << msg << std::endl; std::abort(); }

}  // namespace mfem

#define MFEM_ABORT(msg) ::mfem::mfem_error((std::string("") + msg).c_str())
#define MFEM_VERIFY(x, msg) do { if (!(x)) MFEM_ABORT(msg); } while (0)
