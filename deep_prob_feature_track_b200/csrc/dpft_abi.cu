// C-ABI plumbing: version, per-thread error message.
#include "dpft.h"
#include "dpft_host.h"

namespace dpft {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace dpft

extern "C" int dpft_abi_version(void) { return DPFT_ABI_VERSION; }
extern "C" const char* dpft_last_error(void) { return dpft::g_err; }
