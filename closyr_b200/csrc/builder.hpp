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
template <typename T> struct Buf {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = false;
    ~Buf() { release(); }
    void release()
    {
        if (p) { if (pinned) pinned_free(p); else free(p); }
        p = nullptr; n = cap = 0;
    }
    bool reserve(size_t want)
    {
        if (want <= cap) return true;
        size_t nc = cap ? cap : 1024;
        while (nc < want) nc *= 2;
        T *q = pinned ? (T *)pinned_alloc(nc * sizeof(T)) : (T *)malloc(nc * sizeof(T));
        if (!q && pinned) { pinned = false; q = (T *)malloc(nc * sizeof(T)); }   // no CUDA driver: pageable
        if (!q) return false;
        if (p) { memcpy(q, p, n * sizeof(T)); if (pinned) pinned_free(p); else free(p); }
        p = q; cap = nc;
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

