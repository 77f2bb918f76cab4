// host_api.cpp -- C ABI of the host-side pieces: population builder (flat SoA buffer), text front-end,
// synthetic population generator.  No GPU needed for anything in this file.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/closyr_b200.h"
#include "builder.hpp"
#include "compiler.hpp"
#include "expr.hpp"
#include "internal.hpp"

using namespace closyr;

namespace closyr {

thread_local std::string g_last_error;

void set_error(const std::string &s) { g_last_error = s; }

}  // namespace closyr

extern "C" {

const char *closyr_version(void) { return "closyr_b200 0.2 (ISA v1.1, sm_100a)"; }

// ---- self test of the builder's buffer (tests/test_abi.py): the page-locked allocator fails on the second growth; the
// block allocated first must go back to the page-locked deallocator, the data must survive, later blocks are pageable
namespace {
int g_fake_pinned_allocs = 0, g_fake_pinned_frees = 0;
void *fake_pinned_alloc(size_t bytes) { return ++g_fake_pinned_allocs == 1 ? malloc(bytes) : nullptr; }
void fake_pinned_free(void *p) { g_fake_pinned_frees++; free(p); }
}  // namespace

int closyr_selftest_pinned_fallback(void)
{
    g_fake_pinned_allocs = g_fake_pinned_frees = 0;
    {
        closyr::Buf<uint32_t> b;
        b.pinned = true;
        b.alloc_pinned = fake_pinned_alloc;
        b.free_pinned = fake_pinned_free;
        for (uint32_t i = 0; i < 5000; i++)
            if (!b.push(i * 7u)) return -1;
        if (g_fake_pinned_allocs != 2 || g_fake_pinned_frees != 1) return -2;      // first block pinned, freed as pinned
        if (b.pinned || b.p_is_pinned) return -3;                                   // the current block is pageable
        for (uint32_t i = 0; i < 5000; i++)
            if (b.p[i] != i * 7u) return -4;
    }
    return g_fake_pinned_frees == 1 ? 0 : -5;                                       // the destructor used free()
}

int closyr_builder_create(uint32_t flags, closyr_builder **out)
{
    if (!out) return CLOSYR_E_ARG;
    *out = new closyr_builder((flags & CLOSYR_BUILDER_PINNED) != 0);
    return CLOSYR_OK;
}

void closyr_builder_destroy(closyr_builder *b) { delete b; }

void closyr_builder_clear(closyr_builder *b)
{
    if (b) b->clear();
}

int closyr_builder_add(closyr_builder *b, const char *text)
{
    if (!b || !text) return CLOSYR_E_ARG;
    std::string err;
    NodeP n = parse_expr(text, &err);
    if (!n) { set_error("parse error: " + err); return CLOSYR_E_PARSE; }
    Program pr;
    compile(*n, &pr);
    if (!b->append(pr)) { set_error("out of memory"); return CLOSYR_E_NOMEM; }
    return CLOSYR_OK;
}

int closyr_builder_add_many(closyr_builder *b, const char *const *texts, int64_t n, int nthreads)
{
    if (!b || (!texts && n > 0) || n < 0) return CLOSYR_E_ARG;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    const int64_t BLOCK = 1 << 15;
    std::vector<Program> progs;
    std::vector<char> bad;
    for (int64_t base = 0; base < n; base += BLOCK) {
        int64_t cnt = std::min<int64_t>(BLOCK, n - base);
        progs.assign((size_t)cnt, Program());
        bad.assign((size_t)cnt, 0);
        std::atomic<int64_t> next(0);
        auto work = [&]() {
            for (;;) {
                int64_t i0 = next.fetch_add(256);
                if (i0 >= cnt) break;
                for (int64_t i = i0; i < std::min(cnt, i0 + 256); i++) {
                    NodeP e = parse_expr(texts[base + i], nullptr);
                    if (!e) { bad[(size_t)i] = 1; continue; }
                    compile(*e, &progs[(size_t)i]);
                }
            }
        };
        std::vector<std::thread> th;
        int nt = (int)std::min<int64_t>(nthreads, (cnt + 255) / 256);
        for (int t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
        for (int64_t i = 0; i < cnt; i++) {
            if (bad[(size_t)i]) {
                set_error("parse error in expression " + std::to_string(base + i));
                return CLOSYR_E_PARSE;
            }
            if (!b->append(progs[(size_t)i])) { set_error("out of memory"); return CLOSYR_E_NOMEM; }
        }
    }
    return CLOSYR_OK;
}

int closyr_builder_view(const closyr_builder *b, closyr_population *out)
{
    if (!b || !out) return CLOSYR_E_ARG;
    static const uint32_t zero_off[1] = { 0 };
    out->n_expr = b->n_expr;
    out->code = b->code.p;
    out->code_off = b->n_expr ? b->code_off.p : zero_off;
    out->consts = b->consts.p;
    out->const_off = b->n_expr ? b->const_off.p : zero_off;
    out->meta = b->meta.p;
    out->aux = b->aux.p;
    return CLOSYR_OK;
}

int closyr_to_fullform(const char *text, char **out)
{
    if (!text || !out) return CLOSYR_E_ARG;
    std::string err;
    NodeP n = parse_expr(text, &err);
    if (!n) { set_error("parse error: " + err); return CLOSYR_E_PARSE; }
    std::string s = to_fullform(*n);
    *out = (char *)malloc(s.size() + 1);
    if (!*out) return CLOSYR_E_NOMEM;
    memcpy(*out, s.c_str(), s.size() + 1);
    return CLOSYR_OK;
}

int closyr_leaf_count(const char *text, int32_t *out)
{
    if (!text || !out) return CLOSYR_E_ARG;
    std::string err;
    NodeP n = parse_expr(text, &err);
    if (!n) { set_error("parse error: " + err); return CLOSYR_E_PARSE; }
    *out = leaf_count(*n);
    return CLOSYR_OK;
}

void closyr_free(void *p) { free(p); }

// ------------------------------------------------------------------------------------------------
// synthetic populations.  splitmix64, one stream per expression index, so the
// population is identical for any thread count and in any host language.
// ------------------------------------------------------------------------------------------------
namespace {

struct Rng {
    uint64_t s;
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int below(int n) { return (int)(uni() * n); }
};

NodeP gen_leaf(Rng &r, bool allow_number)
{
    double u = r.uni();
    if (u < 0.60 || !allow_number) return make(Kind::X);
    if (u < 0.85) {
        switch (r.below(7)) {
        case 0: return make_rat(1, 2);
        case 1: return make_rat(1, 10);
        case 2: return make_rat(1, 100);
        case 3: return make_int(2);
        case 4: return make_int(10);
        case 5: return make_int(100);
        default: return make_int(-1);
        }
    }
    static const double REALS[3] = { 1.1, 0.9, 0.1 };
    return make_real(REALS[r.below(3)]);
}

// cumulative weights: n-ary, power-by-constant, sin/cos, exp/log, asin/acos, tan/cot/sec/csc
const double MIX[2][6] = {
    { 0.55, 0.70, 0.84, 0.92, 0.96, 1.00 },      // "GA mix"
    { 0.35, 0.45, 0.70, 0.85, 0.92, 1.00 },      // "transcendental-heavy"
};

NodeP gen_tree(Rng &r, int budget, int mix)
{
    if (budget <= 1) return gen_leaf(r, true);
    if (budget == 2) return make_unary(r.below(2) ? Kind::Sin : Kind::Cos, make(Kind::X));
    double u = r.uni();
    const double *w = MIX[mix ? 1 : 0];
    if (u < w[0]) {
        int arity = (budget >= 5 && r.uni() < 0.3) ? 3 : 2;
        int rest = budget - 1;
        std::vector<int> parts((size_t)arity, 1);
        for (int i = 0; i < rest - arity; i++) parts[(size_t)r.below(arity)]++;
        NodeP n = make(r.below(2) ? Kind::Plus : Kind::Times);
        std::vector<NodeP> nums, xs, comps;
        for (int i = 0; i < arity; i++) {
            NodeP a = parts[(size_t)i] <= 1 ? gen_leaf(r, nums.empty()) : gen_tree(r, parts[(size_t)i], mix);
            if (a->is_number()) nums.push_back(std::move(a));
            else if (a->kind == Kind::X) xs.push_back(std::move(a));
            else comps.push_back(std::move(a));
        }
        // Symja's canonical order puts numbers first, then the symbol, then compound terms
        for (auto &a : nums) n->args.push_back(std::move(a));
        for (auto &a : xs) n->args.push_back(std::move(a));
        for (auto &a : comps) n->args.push_back(std::move(a));
        return n;
    }
    if (u < w[1]) {
        int which = r.below(4);
        NodeP e = which == 0 ? make_int(2) : which == 1 ? make_int(-1) : which == 2 ? make_int(-2) : make_rat(1, 2);
        int eb = which == 3 ? 3 : 1;
        return make_power(gen_tree(r, std::max(1, budget - 1 - eb), mix), std::move(e));
    }
    if (u < w[2]) return make_unary(r.below(2) ? Kind::Sin : Kind::Cos, gen_tree(r, budget - 1, mix));
    if (u < w[3]) {
        if (r.below(2)) return make_power(make(Kind::E), gen_tree(r, std::max(1, budget - 2), mix));
        return make_unary(Kind::Log, gen_tree(r, budget - 1, mix));
    }
    if (u < w[4]) return make_unary(r.below(2) ? Kind::ArcSin : Kind::ArcCos, gen_tree(r, budget - 1, mix));
    static const Kind T[4] = { Kind::Tan, Kind::Cot, Kind::Sec, Kind::Csc };
    return make_unary(T[r.below(4)], gen_tree(r, budget - 1, mix));
}

NodeP gen_expr(uint64_t seed, int64_t index, int max_leafs, int mix)
{
    Rng r{ seed * 0x9E3779B97F4A7C15ull + (uint64_t)index * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull };
    r.next();
    int lo = std::max(1, max_leafs / 2);
    for (int attempt = 0; attempt < 16; attempt++) {
        int target = lo + r.below(std::max(1, max_leafs - lo + 1));
        target = std::max(1, target - attempt * (max_leafs / 16));
        NodeP e = gen_tree(r, target, mix);
        if (leaf_count(*e) <= max_leafs) return e;
    }
    return make(Kind::X);
}

}  // namespace

int closyr_generate_population(closyr_builder *b, uint64_t seed, int64_t n_expr, int32_t max_leafs, int32_t mix,
                               char **out_texts)
{
    if (!b || n_expr < 0 || max_leafs < 1) return CLOSYR_E_ARG;
    int nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    const int64_t BLOCK = 1 << 15;
    std::vector<Program> progs;
    for (int64_t base = 0; base < n_expr; base += BLOCK) {
        int64_t cnt = std::min<int64_t>(BLOCK, n_expr - base);
        progs.assign((size_t)cnt, Program());
        std::atomic<int64_t> next(0);
        std::atomic<int> oom(0);
        auto work = [&]() {
            for (;;) {
                int64_t i0 = next.fetch_add(256);
                if (i0 >= cnt) break;
                for (int64_t i = i0; i < std::min(cnt, i0 + 256); i++) {
                    NodeP e = gen_expr(seed, base + i, max_leafs, mix);
                    compile(*e, &progs[(size_t)i]);
                    if (out_texts) {
                        std::string s = to_fullform(*e);
                        char *t = (char *)malloc(s.size() + 1);
                        if (!t) { oom = 1; continue; }
                        memcpy(t, s.c_str(), s.size() + 1);
                        out_texts[base + i] = t;
                    }
                }
            }
        };
        std::vector<std::thread> th;
        int nt = (int)std::min<int64_t>(nthreads, (cnt + 255) / 256);
        for (int t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
        if (oom) { set_error("out of memory"); return CLOSYR_E_NOMEM; }
        for (int64_t i = 0; i < cnt; i++)
            if (!b->append(progs[(size_t)i])) { set_error("out of memory"); return CLOSYR_E_NOMEM; }
    }
    return CLOSYR_OK;
}

}  // extern "C"
