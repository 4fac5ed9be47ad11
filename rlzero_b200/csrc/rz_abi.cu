// rlzero_b200 -- ABI bookkeeping: version, struct sizes, last-error string.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "rz_common.cuh"

static thread_local char rz_error_buf[512] = "";

void rz_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(rz_error_buf, sizeof(rz_error_buf), fmt, ap);
  va_end(ap);
}

bool rz_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RZ_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

unsigned long long* rz_probe_buffer = nullptr;
extern "C" int rz_debug_set_probe(void* device_buffer) {
  rz_probe_buffer = (unsigned long long*)device_buffer;
  return 0;
}

extern "C" int rz_abi_version(void) { return RZ_ABI_VERSION; }
extern "C" const char* rz_last_error(void) { return rz_error_buf; }
extern "C" int rz_sizeof_tree_desc(void) { return (int)sizeof(rz_tree_desc); }
extern "C" int rz_sizeof_traj_desc(void) { return (int)sizeof(rz_traj_desc); }

namespace rz {
void set_error_tmap(const char* what, int code) { rz_set_error("%s (%d)", what, code); }
}  // namespace rz
