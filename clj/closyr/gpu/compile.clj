(ns closyr.gpu.compile
  "Lowers a canonical Symja IExpr (what `ops-common/->phenotype` holds, src/closyr/ops/common.clj:135-146)
   to the postfix bytecode of include/closyr_b200.h (ISA v1.1).

   This is the Clojure mirror of closyr_b200/csrc/compiler.cpp.  It MUST emit the same instruction words,
   the same constant pool (first-use order, de-duplicated by bit pattern) and the same meta/aux words for
   the same tree; tests/golden/bytecode_golden.json holds the expected encoding for 126 expressions
   (the 96 mutation outputs of test/closyr/ops_modify_test.clj among them).  It cannot be executed in the
   build environment of this repository (no JVM); `closyr.gpu.compile-test` (clj/test/closyr/gpu/compile_test.clj) replays the golden file."
  (:import
    (org.matheclipse.core.interfaces IAST IExpr IFraction IInteger INum ISymbol IComplex)
    (org.matheclipse.core.expression F S)
    (java.util ArrayList)))

(set! *warn-on-reflection* true)

;; opcodes (include/closyr_b200.h)
(def ^:const END 0) (def ^:const LDX 1) (def ^:const LDC 2) (def ^:const LDZ 3) (def ^:const LDSPEC 4)
(def ^:const ADDX 5) (def ^:const ADDC 6) (def ^:const MULX 7) (def ^:const MULC 8) (def ^:const NEG 9)
(def ^:const RECIP 10) (def ^:const SQR 11) (def ^:const SQRT 12) (def ^:const RSQRT 13) (def ^:const POWI 14)
(def ^:const POWC 15) (def ^:const EXP 16) (def ^:const LOG 17) (def ^:const SIN 18) (def ^:const COS 19)
(def ^:const TAN 20) (def ^:const ASIN 21) (def ^:const ACOS 22) (def ^:const ABS 23)
(def ^:const ST 24) (def ^:const ADDS 25) (def ^:const MULS 26) (def ^:const POWS 27)
(def ^:const ST0 28) (def ^:const ADDS0 32) (def ^:const MULS0 36)
(def ^:const REG-SLOTS 4) (def ^:const MAX-SLOTS 64) (def ^:const MAX-CODE 2048) (def ^:const MAX-CONSTS 1024)
(def ^:const F-CONST 1) (def ^:const F-UNSUPPORTED 2) (def ^:const F-NIL 4)

(defn- insn ^long [^long op ^long a ^long b]
  (bit-or op (bit-shift-left a 8) (bit-shift-left (bit-and b 0xffff) 16)))

(defn- leaf-value
  "Numeric value of a real leaf the way the kernels see it, or nil.  Integer: correctly rounded double;
   Rational: double(num)/double(den); Real: itself; E, Pi: Math/E, Math/PI; +-Infinity: IEEE."
  [^IExpr e]
  (cond
    (instance? IInteger e) (.doubleValue ^IInteger e)
    (instance? IFraction e) (/ (.doubleValue (.toBigNumerator ^IFraction e))
                               (.doubleValue (.toBigDenominator ^IFraction e)))
    (instance? INum e) (.doubleValue ^INum e)
    (.equals e S/E) Math/E
    (.equals e S/Pi) Math/PI
    (.isInfinity e) Double/POSITIVE_INFINITY
    (.isNegativeInfinity e) Double/NEGATIVE_INFINITY
    :else nil))

(defn- leaf?
  [^IExpr e ^ISymbol x]
  (or (.equals e x) (some? (leaf-value e))))

(defn- pow-form
  "[:int n] | :sqrt | :rsqrt | :exp | :const | :general   (same order of tests as compiler.cpp pow_form)"
  [^IAST p ^ISymbol x]
  (let [b (.arg1 p) e (.arg2 p)]
    (cond
      (and (instance? IInteger e) (let [n (.toLongDefault ^IInteger e 0)] (and (not= n 0) (<= -32 n 32))))
      [:int (.toLongDefault ^IInteger e 0)]
      (.equals e F/C1D2) :sqrt
      (.equals e F/CN1D2) :rsqrt
      (.equals b S/E) :exp
      (and (some? (leaf-value e)) (not (.equals e x))) :const
      :else :general)))

(declare need unary-ops)

(defn- need-nary [^IAST n ^ISymbol x]
  (let [k (.argSize n)
        arg (fn [i] (.get n (int (inc i))))
        l0 (leaf? (arg 0) x)
        l1 (and (> k 1) (leaf? (arg 1) x))
        [cur idx] (cond
                    (and l0 (> k 1) (not l1)) [(need (arg 1) x) 2]
                    l0 [0 1]
                    (and (> k 1) (not l1)) (let [n0 (need (arg 0) x) n1 (need (arg 1) x)]
                                             [(if (> n1 n0) (max n1 (inc n0)) (max n0 (inc n1))) 2])
                    :else [(need (arg 0) x) 1])]
    (reduce (fn [c i] (if (leaf? (arg i) x) c (max c (inc (need (arg i) x))))) cur (range idx k))))

(defn- need
  "Spill slots needed to evaluate e (Sethi-Ullman for the first two operands of Plus/Times only)."
  [^IExpr e ^ISymbol x]
  (cond
    (or (.isPlus e) (.isTimes e)) (need-nary e x)
    (.isPower e) (let [f (pow-form e x)]
                   (case (if (vector? f) :int f)
                     :exp (need (.arg2 ^IAST e) x)
                     :general (max (need (.arg1 ^IAST e) x) (inc (need (.arg2 ^IAST e) x)))
                     (need (.arg1 ^IAST e) x)))
    ;; only the heads `gen!` compiles: an unknown head is one LDSPEC and needs nothing (compiler.cpp need())
    (and (.isAST e) (= 1 (.argSize ^IAST e)) (unary-ops (str (.head ^IAST e)))) (need (.arg1 ^IAST e) x)
    :else 0))

(defn- emitter []
  {:code (ArrayList.) :consts (ArrayList.) :overflow (volatile! false) :max-slot (volatile! -1)})

(def ^:const M-SRCX 1) (def ^:const M-PUSH 2)

(defn- emit!
  "Appends one instruction, applying the ISA v1.1 fusions exactly as Emitter::emit in compiler.cpp does:
   `ST0+k ; LDX|LDC` becomes the load with M-PUSH and slot k, `LDX ; f` becomes f with M-SRCX."
  [{:keys [^ArrayList code overflow]} op a b]
  (if (> (+ (.size code) 8) MAX-CODE)
    (vreset! overflow true)
    (let [n    (.size code)
          last (if (pos? n) (long (.get code (dec n))) 0)
          lop  (bit-and last 0xff)
          la   (bit-and (bit-shift-right last 8) 0xff)]
      (cond
        (and (pos? n) (or (= op LDX) (= op LDC)) (>= lop ST0) (< lop ADDS0))
        (.set code (dec n) (insn op (bit-or M-PUSH (bit-shift-left (- lop ST0) 2)) b))

        (and (pos? n) (>= op NEG) (<= op ABS) (= lop LDX))
        (.set code (dec n) (insn op (bit-or la M-SRCX) b))

        :else (.add code (insn op a b))))))

(defn- konst! ^long [{:keys [^ArrayList consts overflow]} ^double v]
  (let [bits (Double/doubleToRawLongBits v)
        i (loop [i 0] (cond (= i (.size consts)) -1
                            (= bits (Double/doubleToRawLongBits (double (.get consts i)))) i
                            :else (recur (inc i))))]
    (cond (>= i 0) i
          (> (+ (.size consts) 4) MAX-CONSTS) (do (vreset! overflow true) 0)
          :else (do (.add consts v) (dec (.size consts))))))

(defn- slot-op! [{:keys [overflow max-slot] :as em} generic fast0 s]
  (if (>= s MAX-SLOTS)
    (vreset! overflow true)
    (do (vswap! max-slot max s)
        (if (and (< s REG-SLOTS) (>= fast0 0)) (emit! em (+ fast0 s) s 0) (emit! em generic s 0)))))

(declare gen!)

(defn- load-leaf! [em ^IExpr e ^ISymbol x]
  (if (.equals e x) (emit! em LDX 0 0) (emit! em LDC 0 (konst! em (leaf-value e)))))

(defn- apply-leaf! [em plus? ^IExpr e ^ISymbol x]
  (if (.equals e x)
    (emit! em (if plus? ADDX MULX) 0 0)
    (emit! em (if plus? ADDC MULC) 0 (konst! em (leaf-value e)))))

(defn- apply-slot! [em plus? s]
  (if plus? (slot-op! em ADDS ADDS0 s) (slot-op! em MULS MULS0 s)))

(defn- nary! [em ^IAST n ^ISymbol x s]
  (let [plus? (.isPlus n)
        k (.argSize n)
        arg (fn [i] (.get n (int (inc i))))
        l0 (leaf? (arg 0) x)
        l1 (and (> k 1) (leaf? (arg 1) x))
        idx (cond
              (and l0 (> k 1) (not l1)) (do (gen! em (arg 1) x s) (apply-leaf! em plus? (arg 0) x) 2)
              l0 (do (load-leaf! em (arg 0) x) 1)
              (and (> k 1) (not l1)) (let [first (if (> (need (arg 1) x) (need (arg 0) x)) 1 0)]
                                       (gen! em (arg first) x s)
                                       (slot-op! em ST ST0 s)
                                       (gen! em (arg (- 1 first)) x (inc s))
                                       (apply-slot! em plus? s)
                                       2)
              :else (do (gen! em (arg 0) x s) 1))]
    (doseq [i (range idx k)]
      (if (leaf? (arg i) x)
        (apply-leaf! em plus? (arg i) x)
        (do (slot-op! em ST ST0 s) (gen! em (arg i) x (inc s)) (apply-slot! em plus? s))))))

(def ^:private unary-ops
  {"Sin" [SIN] "Cos" [COS] "Tan" [TAN] "Cot" [TAN RECIP] "Sec" [COS RECIP] "Csc" [SIN RECIP]
   "ArcSin" [ASIN] "ArcCos" [ACOS] "Log" [LOG] "Abs" [ABS]})

(defn- gen! [em ^IExpr e ^ISymbol x s]
  (when-not @(:overflow em)
    (cond
      (leaf? e x) (load-leaf! em e x)
      (instance? IComplex e) (let [re (.doubleValue (.getRealPart ^IComplex e))
                                   im (.doubleValue (.getImaginaryPart ^IComplex e))
                                   ^ArrayList cs (:consts em)
                                   i (konst! em re)
                                   i (if (= (inc i) (.size cs)) i (do (.add cs re) (dec (.size cs))))]
                               (.add cs im)
                               (emit! em LDZ 0 i))
      (.isComplexInfinity e) (emit! em LDSPEC 0 0)
      (.isIndeterminate e) (emit! em LDSPEC 1 0)
      (or (.isPlus e) (.isTimes e)) (nary! em e x s)
      (.isPower e) (let [^IAST p e f (pow-form p x)]
                     (cond
                       (vector? f) (let [n (long (second f))]
                                     (gen! em (.arg1 p) x s)
                                     (cond (= n 2) (emit! em SQR 0 0)
                                           (= n -1) (emit! em RECIP 0 0)
                                           (not= n 1) (emit! em POWI 0 n)))
                       (= f :sqrt) (do (gen! em (.arg1 p) x s) (emit! em SQRT 0 0))
                       (= f :rsqrt) (do (gen! em (.arg1 p) x s) (emit! em RSQRT 0 0))
                       (= f :exp) (do (gen! em (.arg2 p) x s) (emit! em EXP 0 0))
                       (= f :const) (do (gen! em (.arg1 p) x s) (emit! em POWC 0 (konst! em (leaf-value (.arg2 p)))))
                       :else (do (gen! em (.arg1 p) x s) (slot-op! em ST ST0 s)
                                 (gen! em (.arg2 p) x (inc s)) (slot-op! em POWS -1 s))))
      (and (.isAST e) (= 1 (.argSize ^IAST e)) (unary-ops (str (.head ^IAST e))))
      (do (gen! em (.arg1 ^IAST e) x s) (doseq [op (unary-ops (str (.head ^IAST e)))] (emit! em op 0 0)))
      :else (emit! em LDSPEC 1 0))))                                  ; unknown head: never a real number

(defn compile-expr
  "IExpr -> {:code int[] :consts double[] :meta int :aux int}.  Never throws for structural overflow:
   sets F-UNSUPPORTED and emits a bare END (the kernel scores it as min-score)."
  [^IExpr expr ^ISymbol x]
  (let [em (emitter)
        _ (gen! em expr x 0)
        ^ArrayList code (:code em) ^ArrayList consts (:consts em)
        leafs (.leafCount expr)
        const? (.isFree expr x)
        overflow? @(:overflow em)
        aux (if (and const? (.isAST expr) (pos? (.argSize ^IAST expr)) (not overflow?)
                     (< (.argSize ^IAST expr) 65536) (<= (+ (.size consts) (.argSize ^IAST expr) 4) MAX-CONSTS))
              (let [off (.size consts) argc (.argSize ^IAST expr)]
                (doseq [i (range argc)]
                  (let [^IExpr a (.get ^IAST expr (int (inc i)))]
                    (.add consts (if (.isReal a) (leaf-value a) Double/POSITIVE_INFINITY))))   ; eval.clj:109
                (bit-or (bit-shift-left argc 16) off))
              0)
        flags (cond-> 0 const? (bit-or F-CONST) (.isIndeterminate expr) (bit-or F-NIL) overflow? (bit-or F-UNSUPPORTED))]
    (when overflow? (.clear code) (.clear consts))
    (.add code (insn END 0 0))
    (while (pos? (mod (.size code) 4)) (.add code (insn END 0 0)))
    (while (pos? (mod (.size consts) 2)) (.add consts 0.0))
    {:code   (int-array (map unchecked-int code))
     :consts (double-array consts)
     :meta   (unchecked-int (bit-or (min leafs 65535) (bit-shift-left flags 16)
                                    (bit-shift-left (if overflow? 0 (inc @(:max-slot em))) 24)))
     :aux    (unchecked-int (if overflow? 0 aux))}))
