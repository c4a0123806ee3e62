#include "solver_context.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>

#include "param.hpp"

namespace {
struct Entry {
  impes_handle h = nullptr;
  impes_desc d;
  const double *kx = nullptr, *ky = nullptr, *kz = nullptr, *phi = nullptr;
  double probe[8] = {0};  // a few values of the K / phi arrays: catches a reload into the same buffers
  bool comm_ready = false;
  std::map<const mfem::FiniteElementSpace*, DofMap> maps;
};
struct Registry {
  std::map<const Param*, Entry> handles;
  ~Registry() {
    for (auto& kv : handles) impes_destroy(kv.second.h);
  }
};
Registry& registry() {
  static Registry r;
  return r;
}
void fill_probe(const Param& p, double* probe) {
  const long n = p.get_n_cells();
  const double* a[4] = {p.K_array_x, p.K_array_y, p.K_array_z, p.phi_array};
  for (int i = 0; i < 4; ++i) {
    probe[2 * i] = a[i] ? a[i][0] + 3.0 * a[i][n / 2] : 0.0;
    probe[2 * i + 1] = a[i] ? a[i][n - 1] - 5.0 * a[i][n / 3] : 0.0;
  }
}
}  // namespace

impes_handle impes_host_context(const Param& param, int substep_rule, int rank, int world_size) {
  Registry& r = registry();
  MFEM_VERIFY(param.K_array_x && param.K_array_y && param.phi_array, "Param::init_arrays() has not been called");
  impes_desc d = param.to_desc();
  d.substep_rule = substep_rule;
  d.rank = rank;
  d.world_size = world_size;
  double probe[8];
  fill_probe(param, probe);
  auto it = r.handles.find(&param);
  if (it != r.handles.end()) {
    Entry& e = it->second;
    const bool same = !std::memcmp(&e.d, &d, sizeof d) && e.kx == param.K_array_x && e.ky == param.K_array_y &&
                      e.kz == param.K_array_z && e.phi == param.phi_array && !std::memcmp(e.probe, probe, sizeof probe);
    if (same) return e.h;
    impes_destroy(e.h);  // Param changed since the handle was built: start over, like the stateless reference
    r.handles.erase(it);
  }
  Entry e;
  const int rc = impes_create(&d, param.K_array_x, param.K_array_y, param.K_array_z, param.phi_array, &e.h);
  if (rc != IMPES_OK) MFEM_ABORT("impes_create failed: " + std::string(impes_last_error(nullptr)));
  e.d = d;
  e.kx = param.K_array_x; e.ky = param.K_array_y; e.kz = param.K_array_z; e.phi = param.phi_array;
  std::memcpy(e.probe, probe, sizeof probe);
  r.handles[&param] = e;
  return e.h;
}

void impes_host_release(const Param& param) {
  Registry& r = registry();
  auto it = r.handles.find(&param);
  if (it == r.handles.end()) return;
  impes_destroy(it->second.h);
  r.handles.erase(it);
}

const DofMap& impes_host_dofmap(const Param& param, mfem::FiniteElementSpace& fes, bool rt, int order) {
  Registry& r = registry();
  auto it = r.handles.find(&param);
  MFEM_VERIFY(it != r.handles.end(), "impes_host_dofmap: no context for this Param");
  auto& maps = it->second.maps;
  auto mi = maps.find(&fes);
  if (mi != maps.end() && (int)mi->second.nat.size() == fes.GetVSize()) return mi->second;
  const int n[3] = {param.nx, param.ny, param.nz};
  const double size[3] = {param.sx, param.sy, param.sz};
  maps[&fes] = build_dofmap(fes, rt, order, param.spacedim, n, size);
  return maps[&fes];
}

void impes_host_comm_bootstrap(const Param& param, int rank, int world_size) {
  if (world_size <= 1) return;
  Registry& r = registry();
  auto it = r.handles.find(&param);
  MFEM_VERIFY(it != r.handles.end(), "impes_host_comm_bootstrap: no context for this Param");
  Entry& e = it->second;
  if (e.comm_ready) return;
  const char* path = std::getenv("IMPES_NCCL_ID_FILE");
  MFEM_VERIFY(path, "multi-process run: set IMPES_NCCL_ID_FILE to a path all ranks can read");
  unsigned char id[128];
  if (rank == 0) {
    MFEM_VERIFY(impes_comm_unique_id(id) == IMPES_OK, "impes_comm_unique_id failed");
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    MFEM_VERIFY(f && std::fwrite(id, 1, 128, f) == 128, "cannot write the NCCL id file");
    std::fclose(f);
    MFEM_VERIFY(std::rename(tmp.c_str(), path) == 0, "cannot publish the NCCL id file");
  } else {
    size_t got = 0;
    for (int tries = 0; tries < 6000 && got != 128; ++tries) {  // up to 10 minutes
      if (FILE* f = std::fopen(path, "rb")) {
        got = std::fread(id, 1, 128, f);
        std::fclose(f);
      }
      if (got != 128) std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
    MFEM_VERIFY(got == 128, "timed out waiting for the NCCL id file");
  }
  if (impes_comm_init(e.h, id) != IMPES_OK) MFEM_ABORT("impes_comm_init failed: " + std::string(impes_last_error(e.h)));
  e.comm_ready = true;
}
