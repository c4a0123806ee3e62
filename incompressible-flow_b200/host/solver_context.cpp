#include "solver_context.hpp"

#include <map>
#include <string>

#include "param.hpp"

namespace {
struct Registry {
  std::map<const Param*, impes_handle> handles;
  ~Registry() {
    for (auto& kv : handles) impes_destroy(kv.second);
  }
};
Registry& registry() {
  static Registry r;
  return r;
}
}  // namespace

impes_handle impes_host_context(const Param& param) {
  Registry& r = registry();
  auto it = r.handles.find(&param);
  if (it != r.handles.end()) return it->second;
  MFEM_VERIFY(param.K_array_x && param.K_array_y && param.phi_array, "Param::init_arrays() has not been called");
  const impes_desc d = param.to_desc();
  impes_handle h = nullptr;
  const int rc = impes_create(&d, param.K_array_x, param.K_array_y, param.K_array_z, param.phi_array, &h);
  if (rc != IMPES_OK) MFEM_ABORT("impes_create failed: " + std::string(impes_last_error(nullptr)));
  r.handles[&param] = h;
  return h;
}

void impes_host_release(const Param& param) {
  Registry& r = registry();
  auto it = r.handles.find(&param);
  if (it == r.handles.end()) return;
  impes_destroy(it->second);
  r.handles.erase(it);
}
