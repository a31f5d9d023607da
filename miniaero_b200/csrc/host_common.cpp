#include "host_common.h"

#include "miniaero_b200.h"

namespace {
thread_local std::string g_last_error;
}

int ma_set_error(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

extern "C" {
const char *ma_last_error(void) { return g_last_error.c_str(); }
int ma_abi_version(void) { return MA_ABI_VERSION; }
}
