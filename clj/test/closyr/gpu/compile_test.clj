(ns closyr.gpu.compile-test
  "Replays tests/golden/bytecode_golden.edn (126 expressions, written by the C++ compiler through
   tests/golden/make_bytecode_golden.py) against the Clojure mirror of the compiler: instruction words, constant
   pool bit patterns, meta and aux words must be identical.  Needs a JVM with closyr's own dependencies (Symja);
   it cannot run in the build image of this repository.

   Run from a closyr checkout with clj/ on the source path and clj/test on the test path:
     lein test :only closyr.gpu.compile-test"
  (:require
    [clojure.edn :as edn]
    [clojure.java.io :as io]
    [clojure.test :refer [deftest is testing]]
    [closyr.gpu.compile :as gc]
    [closyr.ops.common :as ops-common])
  (:import
    (org.matheclipse.core.eval ExprEvaluator)
    (org.matheclipse.core.interfaces IExpr)))

(def ^:private golden-path
  (or (System/getenv "CLOSYR_B200_GOLDEN") "tests/golden/bytecode_golden.edn"))

(defn- canonical
  "What `->phenotype` holds for a printed expression: parse, then `.eval` (src/closyr/ops/common.clj:135-146)."
  ^IExpr [^String text]
  (let [util (ExprEvaluator. (ops-common/new-eval-engine) true 0)]
    (.eval util (.parse util text))))

(defn- u32 [x] (bit-and (long x) 0xffffffff))

(defn- bits-hex [^double d] (format "%016x" (Double/doubleToRawLongBits d)))

(deftest bytecode-matches-the-golden-file
  (let [programs (edn/read-string (slurp (io/file golden-path)))]
    (is (>= (count programs) 120))
    (doseq [{:keys [text code consts meta aux]} programs]
      (testing text
        ;; the variable of the golden corpus is printed as x; closyr's own symbol is (F/Dummy "x") = ops-common/sym-x
        (let [expr (.subs (canonical text) (.parse (ExprEvaluator.) "x") ops-common/sym-x)
              got  (gc/compile-expr expr ops-common/sym-x)]
          (is (= (mapv u32 code) (mapv u32 (:code got))))
          (is (= consts (mapv bits-hex (:consts got))))
          (is (= (u32 meta) (u32 (:meta got))))
          (is (= (u32 aux) (u32 (:aux got)))))))))
