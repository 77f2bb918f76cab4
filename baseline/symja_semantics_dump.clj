;; baseline/symja_semantics_dump.clj -- asks Symja the questions the value algebra of DESIGN.md section 3 answers by
;; assumption ("parity unpinned": poles, complex intermediates, overflow, the real/complex collapse), and prints
;; every disagreement with the oracle.  Needs a JVM and closyr's classpath; cannot run in this repository's image.
;;
;;   lein run -m clojure.main /path/to/baseline/symja_semantics_dump.clj /path/to/baseline/semantics_cases.edn
;;
;; For each case the reference's own entry point is used: ops-eval/eval-vec-pheno on the phenotype
;; (src/closyr/ops/eval.clj:99-112), which yields a double per point, ##Inf where the element is not a real number,
;; nil for an Indeterminate result, or throws for a constant that is not real.  The oracle's answer to the same
;; call is in the :eval-vec-pheno key of the case.  The raw Symja result is printed next to every disagreement.
(require '[clojure.edn :as edn]
         '[closyr.ops.common :as ops-common]
         '[closyr.ops.eval :as ops-eval])
(import '(org.matheclipse.core.eval ExprEvaluator))

(defn close-enough? [a b]
  (or (= a b)
      (and (Double/isNaN a) (Double/isNaN b))
      (<= (Math/abs (- a b)) (* 1e-9 (max 1.0 (Math/abs b))))))

(let [{:keys [xs cases]} (edn/read-string {:readers {}} (slurp (first *command-line-args*)))
      x        ops-common/sym-x
      parser   (ExprEvaluator.)
      x-text   (.parse parser "x")
      run-args {:input-xs-list  (ops-common/exprs->exprs-list (ops-common/doubles->exprs xs))
                :input-xs-count (count xs)}
      bad      (atom 0)]
  (doseq [{:keys [expr eval-vec-pheno]} cases]
    (let [pheno (ops-common/->phenotype x (.subs (.parse parser expr) x-text x) nil)
          got   (try (ops-eval/eval-vec-pheno pheno run-args) (catch Exception _ :throws))
          same? (cond
                  (or (keyword? got) (keyword? eval-vec-pheno) (nil? got) (nil? eval-vec-pheno)) (= got eval-vec-pheno)
                  :else (and (= (count got) (count eval-vec-pheno))
                             (every? true? (map close-enough? (map double got) (map double eval-vec-pheno)))))]
      (when-not same?
        (swap! bad inc)
        (println "DISAGREE" expr)
        (println "  symja :" got)
        (println "  oracle:" eval-vec-pheno)
        (println "  raw   :" (str (ops-eval/eval-phenotype-on-expr-args pheno (:input-xs-list run-args)))))))
  (println (count cases) "cases," @bad "disagreements")
  (shutdown-agents))
