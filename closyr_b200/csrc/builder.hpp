// builder.hpp -- the population builder behind closyr_builder_* (flat struct-of-arrays buffer, optionally
// page-locked), shared by host_api.cpp and ga.cpp.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/closyr_b200.h"
#include "compiler.hpp"
#include "internal.hpp"

namespace closyr {

// growable host buffer, optionally page-locked so the H2D copy of a generation is a single DMA
// `alloc_pinned` / `free_pinned` are replaceable so that a test can make the page-locked allocator fail.
template <typename T> struct Buf {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = false;          // page-locked memory is wanted for the next allocation
    bool p_is_pinned = false;     // how the CURRENT block was allocated (decides how it is released)
    void *(*alloc_pinned)(size_t) = pinned_alloc;
    void (*free_pinned)(void *) = pinned_free;
    ~Buf() { release(); }
    void release_block(T *q, bool was_pinned)
    {
        if (!q) return;
        if (was_pinned) free_pinned(q);
        else free(q);
    }
    void release()
    {
        release_block(p, p_is_pinned);
        p = nullptr; n = cap = 0; p_is_pinned = false;
    }
    bool reserve(size_t want)
    {
        if (want <= cap) return true;
        size_t nc = cap ? cap : 1024;
        while (nc < want) nc *= 2;
        bool q_pinned = pinned;
        T *q = pinned ? (T *)alloc_pinned(nc * sizeof(T)) : (T *)malloc(nc * sizeof(T));
        if (!q && pinned) {                        // no CUDA driver, or the page-lock limit is reached: pageable
            pinned = false;
            q_pinned = false;
            q = (T *)malloc(nc * sizeof(T));
        }
        if (!q) return false;
        if (p) memcpy(q, p, n * sizeof(T));
        release_block(p, p_is_pinned);             // the OLD block goes back to the allocator it came from
        p = q; cap = nc; p_is_pinned = q_pinned;
        return true;
    }
    bool append(const T *src, size_t k)
    {
        if (!reserve(n + k)) return false;
        memcpy(p + n, src, k * sizeof(T));
        n += k;
        return true;
    }
    bool push(T v) { return append(&v, 1); }
};

}  // namespace closyr

struct closyr_builder {
    closyr::Buf<uint32_t> code, code_off, const_off, meta, aux;
    closyr::Buf<double> consts;
    int64_t n_expr = 0;
    explicit closyr_builder(bool pinned)
    {
        code.pinned = code_off.pinned = const_off.pinned = meta.pinned = aux.pinned = consts.pinned = pinned;
    }
    bool append(const closyr::Program &pr)
    {
        if (n_expr == 0 && code_off.n == 0) { if (!code_off.push(0) || !const_off.push(0)) return false; }
        if (code.n + pr.code.size() > 0xFFFFFFF0ull || consts.n + pr.consts.size() > 0xFFFFFFF0ull) return false;
        if (!code.append(pr.code.data(), pr.code.size())) return false;
        if (!consts.append(pr.consts.data(), pr.consts.size())) return false;
        if (!code_off.push((uint32_t)code.n) || !const_off.push((uint32_t)consts.n)) return false;
        if (!meta.push(pr.meta) || !aux.push(pr.aux)) return false;
        n_expr++;
        return true;
    }
    void clear() { code.n = code_off.n = const_off.n = meta.n = aux.n = consts.n = 0; n_expr = 0; }
};

