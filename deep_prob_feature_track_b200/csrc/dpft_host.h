// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cstdarg>
#include <cstdio>

namespace dpft {

// Record a message for dpft_last_error() on this thread and hand the code back.
int set_error(int code, const char* fmt, ...);

}  // namespace dpft
