// Library-level entry points: version, thread-local error string, launch counter.
#include "common.cuh"

namespace estd {

std::atomic<unsigned long long> g_launches{0};

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace estd

extern "C" int estd_version(void) { return ESTD_VERSION; }
extern "C" const char* estd_last_error(void) { return estd::error_buffer(); }
extern "C" unsigned long long estd_launch_count(void) { return estd::g_launches.load(); }
