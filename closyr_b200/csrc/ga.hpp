// ga.hpp -- stand-alone GA driver: the host loop that feeds the scorer
// when there is no JVM.  Mirrors the reference's loop so that "ms per generation" means the same thing:
//   src/closyr/ga.clj:25-88          with-score / compete / evolve (shuffle, compete in pairs, 4:1 mutation:crossover)
//   src/closyr/ops/modify.clj:77-288 modify (5 operator classes), crossover, apply-modifications, mutations-sampler
//   src/closyr/ops/initialize.clj:17-665 the initial population (x) and the 96 mutation operators, in order
//   src/closyr/ops/common.clj:127-177 valid-expr-or-default, should-modify-leaf / -branch / -ast-head
//   src/closyr/util/prng.clj:16-59   one java.util.Random, rand = nextFloat, rand-int = (int (* n (rand)))
// The expression rewrites work on the same trees the compiler lowers (expr.hpp); Symja's `.eval` is
// approximated by the canonicalising constructors plus like-term / like-factor merging, so leaf counts can
// differ from Symja's for some shapes (documented in DESIGN.md); the value of every operator's result is
// pinned against the reference's 96 golden strings (tests/test_ga.py).
#pragma once
#include <cstdint>
#include <vector>

#include "expr.hpp"

namespace closyr {

// java.util.Random: 48-bit LCG (pinned by test/closyr/util_prng_test.clj:22-25)
struct JavaRandom {
    uint64_t s = 0;
    explicit JavaRandom(int64_t seed = 888) { set_seed(seed); }
    void set_seed(int64_t seed) { s = ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
    int32_t next(int bits)
    {
        s = (s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int32_t)(uint32_t)(s >> (48 - bits));
    }
    float next_float() { return (float)next(24) / (float)(1 << 24); }
    int64_t next_long() { int64_t hi = next(32); int64_t lo = next(32); return (int64_t)((uint64_t)hi << 32) + lo; }
    int32_t next_int(int32_t bound)
    {
        int32_t r = next(31);
        const int32_t m = bound - 1;
        if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        for (int32_t u = r; (int32_t)((uint32_t)u - (uint32_t)(r = u % bound) + (uint32_t)m) < 0; u = next(31)) {}
        return r;
    }
    double rand() { return (double)next_float(); }                               // prng.clj:24-29
    int rand_int(int64_t n) { return (int)((double)n * (double)next_float()); }   // prng.clj:32-37
    template <typename T> void shuffle(std::vector<T> &v)                         // Collections.shuffle(list, rng)
    {
        for (int64_t i = (int64_t)v.size(); i > 1; i--) std::swap(v[(size_t)i - 1], v[(size_t)next_int((int32_t)i)]);
    }
};

struct MutationInfo { const char *op; const char *label; };
int mutation_count();
MutationInfo mutation_info(int i);

// Applies mutation i (initialize.clj order).  rng == nullptr: every should-modify-* draw says yes (the
// reference's own test stubs rand to 0.0, ops_modify_test.clj:754-756).  Returns nullptr when the operator
// fails the way Symja throws (1/0).
NodeP apply_mutation(int i, const Node &expr, JavaRandom *rng);
NodeP derivative(const Node &e);
// the tree rebuilt bottom-up through the evaluating constructors (the GA's stand-in for `.eval`)
NodeP canonical_form(const Node &n);

// modify.clj:142-196; returns the new expression, or a copy of e1 when the result is too big or unchanged
NodeP crossover(int max_leafs, const Node &e1, const Node &e2, JavaRandom &rng, bool *changed);
// ops.clj:143-172 + modify.clj:205-256: draws the number of modifications and applies them
NodeP mutate(int max_leafs, const Node &winner, JavaRandom &rng, int *mods_applied);

}  // namespace closyr
