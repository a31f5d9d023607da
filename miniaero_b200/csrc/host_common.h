// Shared host-side helpers of the C-ABI implementation (error channel).
#pragma once
#include <string>

// Records `msg` as the calling thread's last error and returns `code` (a negative ma_status).
int ma_set_error(int code, const std::string &msg);
