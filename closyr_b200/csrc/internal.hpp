// internal.hpp -- declarations shared between the host-only and the CUDA translation units.
#pragma once
#include <cstddef>
#include <string>

namespace closyr {

extern thread_local std::string g_last_error;
void set_error(const std::string &s);

// page-locked host memory (ctx.cu); pinned_alloc returns nullptr when no CUDA driver is present
void *pinned_alloc(size_t bytes);
void pinned_free(void *p);

}  // namespace closyr
