// tracer_b200/csrc/host/error.cpp
#include "error.h"

#include <cstdarg>
#include <cstdio>

namespace trq {

static thread_local char g_err[512] = "";

int fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return status;
}

const char* last_error() { return g_err; }

}  // namespace trq
