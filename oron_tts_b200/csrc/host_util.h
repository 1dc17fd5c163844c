// Host-side helpers shared by the translation units of liboron_b200.so.
#pragma once
#include <atomic>
#include <cstdint>

namespace oron {
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
int num_sms();
}  // namespace oron
