// Error reporting + ABI version for the C boundary (include/extdm_b200.h).
#include "common.cuh"
#include "../../include/extdm_b200.h"
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";

void extdm_set_error(const char* msg, const char* file, int line) {
  const char* base = strrchr(file, '/');
  snprintf(g_err, sizeof g_err, "%s (%s:%d)", msg, base ? base + 1 : file, line);
}

extern "C" const char* extdm_last_error(void) { return g_err; }
extern "C" int extdm_abi_version(void) { return 5; }
extern "C" int extdm_sizeof_gemm(void) { return static_cast<int>(sizeof(ExtdmGemm)); }
