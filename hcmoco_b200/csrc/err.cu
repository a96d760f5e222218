// thread-local error string + ABI version for the C-ABI (include/hcmoco.h)
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void hcm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {
const char* hcm_last_error(void) { return g_err; }
int hcm_abi_version(void) { return 1; }
}
