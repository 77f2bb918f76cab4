"""Population builder: expression text -> postfix bytecode -> one flat struct-of-arrays buffer.

Host-side mirror of north-star subsystem (1).  In the Clojure integration the compiler walks the Symja
IExpr (clj/closyr/gpu/compile.clj); here the same compiler (closyr_b200/csrc/compiler.cpp) is fed Symja
text: `IExpr.fullFormString()` or the printed form the reference logs.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def to_fullform(expr_text):
    """Canonical FullForm text of an expression (what `->phenotype` would hold, common.clj:135-146)."""
    out = C.c_void_p()
    L.check(L.lib().closyr_to_fullform(expr_text.encode(), C.byref(out)))
    try:
        return C.string_at(out).decode()
    finally:
        L.lib().closyr_free(out)


def leaf_count(expr_text):
    """Symja `.leafCount` of the expression (reference src/closyr/ops.clj:127)."""
    out = C.c_int32()
    L.check(L.lib().closyr_leaf_count(expr_text.encode(), C.byref(out)))
    return out.value


class PopulationBuilder:
    def __init__(self, pinned=False):
        self._h = C.c_void_p()
        L.check(L.lib().closyr_builder_create(L.BUILDER_PINNED if pinned else 0, C.byref(self._h)))

    def close(self):
        if self._h:
            L.lib().closyr_builder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        L.lib().closyr_builder_clear(self._h)

    def add(self, expr_text):
        L.check(L.lib().closyr_builder_add(self._h, expr_text.encode()))
        return self

    def add_many(self, expr_texts, nthreads=0):
        n = len(expr_texts)
        arr = (C.c_char_p * n)(*[e.encode() if isinstance(e, str) else e for e in expr_texts])
        L.check(L.lib().closyr_builder_add_many(self._h, arr, n, nthreads))
        return self

    def generate(self, seed, n_expr, max_leafs=40, mix=0, want_texts=False):
        """Appends a seeded synthetic population; returns the FullForm texts if asked."""
        texts = None
        if want_texts:
            texts = (C.c_void_p * n_expr)()
        L.check(L.lib().closyr_generate_population(self._h, seed, n_expr, max_leafs, mix, texts))
        if not want_texts:
            return None
        out = []
        for p in texts:
            out.append(C.string_at(p).decode())
            L.lib().closyr_free(p)
        return out

    def view(self):
        """The closyr_population struct (pointers stay valid until the builder is modified or closed)."""
        pop = L.Population()
        L.check(L.lib().closyr_builder_view(self._h, C.byref(pop)))
        return pop

    @property
    def n_expr(self):
        return self.view().n_expr

    def arrays(self):
        """Copies of the SoA arrays as numpy arrays (for tests and golden files)."""
        v = self.view()
        n = v.n_expr
        if n == 0:
            z = np.zeros(0, dtype=np.uint32)
            return dict(code=z, code_off=np.zeros(1, np.uint32), consts=np.zeros(0), const_off=np.zeros(1, np.uint32),
                        meta=z, aux=z)
        code_off = np.ctypeslib.as_array(v.code_off, shape=(n + 1,)).copy()
        const_off = np.ctypeslib.as_array(v.const_off, shape=(n + 1,)).copy()
        nc, nk = int(code_off[-1]), int(const_off[-1])
        return dict(
            code=np.ctypeslib.as_array(v.code, shape=(nc,)).copy() if nc else np.zeros(0, np.uint32),
            code_off=code_off,
            consts=np.ctypeslib.as_array(v.consts, shape=(nk,)).copy() if nk else np.zeros(0),
            const_off=const_off,
            meta=np.ctypeslib.as_array(v.meta, shape=(n,)).copy(),
            aux=np.ctypeslib.as_array(v.aux, shape=(n,)).copy(),
        )

    def nbytes(self):
        a = self.arrays()
        return int(sum(x.nbytes for x in a.values()))


OPNAMES = {
    0: "END", 1: "LDX", 2: "LDC", 3: "LDZ", 4: "LDSPEC", 5: "ADDX", 6: "ADDC", 7: "MULX", 8: "MULC", 9: "NEG",
    10: "RECIP", 11: "SQR", 12: "SQRT", 13: "RSQRT", 14: "POWI", 15: "POWC", 16: "EXP", 17: "LOG", 18: "SIN",
    19: "COS", 20: "TAN", 21: "ASIN", 22: "ACOS", 23: "ABS", 24: "ST", 25: "ADDS", 26: "MULS", 27: "POWS",
    28: "ST0", 29: "ST1", 30: "ST2", 31: "ST3", 32: "ADDS0", 33: "ADDS1", 34: "ADDS2", 35: "ADDS3",
    36: "MULS0", 37: "MULS1", 38: "MULS2", 39: "MULS3",
}


def disassemble(arrays, i, expand=False):
    """Human-readable listing of program i (used by tests and docs).  ISA v1.1 modifiers are shown as
    `push k:` (slot[k] = acc first) and `f(x)` (the operand is x); expand=True lists the equivalent v1
    instruction sequence instead (ST0+k ; LDX ; f)."""
    c0, c1 = int(arrays["code_off"][i]), int(arrays["code_off"][i + 1])
    k0 = int(arrays["const_off"][i])
    out = []
    for w in arrays["code"][c0:c1]:
        w = int(w)
        op, a, b = w & 0xFF, (w >> 8) & 0xFF, w >> 16
        name = OPNAMES.get(op, "?%d" % op)
        if name == "END":
            out.append("END")
            break
        if name in ("LDC", "ADDC", "MULC", "POWC", "LDZ"):
            text = "%s %r" % (name, float(arrays["consts"][k0 + b]))
        elif name == "POWI":
            text = "POWI %d" % (b - 65536 if b >= 32768 else b)
        elif name in ("ST", "ADDS", "MULS", "POWS", "LDSPEC"):
            text = "%s %d" % (name, a)
        else:
            text = name
        modifiers = op in (1, 2) or 9 <= op <= 23
        push = modifiers and (a & 2)
        srcx = modifiers and 9 <= op <= 23 and (a & 1)
        if expand:
            if push:
                out.append("ST%d" % ((a >> 2) & 3))
            if srcx:
                out.append("LDX")
            out.append(text)
        else:
            if srcx:
                head, _, rest = text.partition(" ")
                text = head + "(x)" + ((" " + rest) if rest else "")
            out.append(("push %d: " % ((a >> 2) & 3) if push else "") + text)
    return out
