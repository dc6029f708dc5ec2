// tracer_b200/csrc/host/error.h -- thread-local "last error" text behind trq_last_error_string().
#pragma once

namespace trq {
int fail(int status, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
const char* last_error();
}  // namespace trq
