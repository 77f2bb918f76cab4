// compiler.hpp -- lowers an expression tree to the postfix bytecode of include/closyr_b200.h.
// North star subsystem (1).  The Clojure mirror is clj/closyr/gpu/compile.clj; both must emit identical
// words and constants for the same tree (tests/golden/bytecode_golden.json pins the encoding).
#pragma once
#include <cstdint>
#include <vector>

#include "expr.hpp"

namespace closyr {

struct Program {
    std::vector<uint32_t> code;   // padded with CL_END to a multiple of 4 words
    std::vector<double> consts;   // padded with 0.0 to a multiple of 2
    uint32_t meta = 0;
    uint32_t aux = 0;
    int leafs = 0;
    int nslots = 0;
    int n_insn = 0;               // instructions before the END
};

// Never fails: structural overflows set CL_F_UNSUPPORTED in meta and emit a bare END.
void compile(const Node &root, Program *out);

}  // namespace closyr
