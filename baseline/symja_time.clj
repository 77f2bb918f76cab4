;; baseline/symja_time.clj -- times the REFERENCE's own scoring path (Symja under pmap on the JVM) on the inputs
;; bench.py uses, so that the "cpu_baseline" line of the bench can be replaced by the true reference number
;; wherever a JVM and closyr's dependencies are available.  It cannot run in this repository's build image
;; (no JVM, no jars, no network).
;;
;; Usage, from a checkout of ogeagla/closyr (it needs closyr's classpath, nothing of this repository):
;;   python scripts/dump_population.py c3 /tmp/pop_c3.txt        # in THIS repository: the seeded population as text
;;   lein run -m clojure.main /path/to/baseline/symja_time.clj /tmp/pop_c3.txt 4096 40 3
;; arguments: population file (one expression per line, Symja FullForm text), points, max-leafs, repetitions.
;;
;; What is timed is exactly what ga/evolve does to an unscored population (src/closyr/ga.clj:62-65):
;;   (doall (pmap (partial ops/score-fn run-args run-config) phenos))
;; with run-args built by the reference's own marshalling (src/closyr/symbolic_regression.clj:355-383).
(require '[closyr.ops :as ops]
         '[closyr.ops.common :as ops-common]
         '[closyr.symbolic-regression :as sr]
         '[clojure.string :as str])
(import '(org.matheclipse.core.eval ExprEvaluator))

(let [[pop-file n-str max-leafs-str reps-str] *command-line-args*
      n          (Long/parseLong (or n-str "4096"))
      max-leafs  (Long/parseLong (or max-leafs-str "40"))
      reps       (Long/parseLong (or reps-str "3"))
      ;; the synthetic points of SURVEY.md section 8d (bench.py synth_points)
      xs         (mapv #(* (/ Math/PI 15.0) 49.0 (/ (double %) (max 1 (dec n)))) (range n))
      ys         (mapv #(+ 2.0 (/ (* 1.5 %) Math/PI) (Math/cos %)) xs)
      x          ops-common/sym-x
      parser     (ExprEvaluator.)
      x-text     (.parse parser "x")
      phenos     (->> (str/split-lines (slurp pop-file))
                      (remove str/blank?)
                      (mapv (fn [line]
                              ;; the file holds the variable as `x`; closyr's own symbol is (F/Dummy "x")
                              (ops-common/->phenotype x (.subs (.parse parser line) x-text x) nil))))
      ;; ->run-args insists on :iters and a population count/phenos (symbolic_regression.clj:355-383); note that the
      ;; reference's schema caps points at 10000 and the population at 100000 (spec.clj:140-145, row N3)
      run-args   (sr/->run-args {:input-xs-exprs     (ops-common/doubles->exprs xs)
                                 :input-xs-vec       xs
                                 :input-ys-vec       ys
                                 :iters              1
                                 :input-phenos-count (count phenos)
                                 :max-leafs          max-leafs})
      run-config {:max-leafs max-leafs}
      score      (partial ops/score-fn run-args run-config)
      cores      (.availableProcessors (Runtime/getRuntime))]
  (println "phenotypes" (count phenos) "points" n "max-leafs" max-leafs "cores" cores)
  (dotimes [_ 1] (doall (pmap score (take (min 200 (count phenos)) phenos))))            ; warm-up (JIT)
  (dotimes [r reps]
    (let [t0 (System/nanoTime)
          scores (doall (pmap score phenos))
          dt (/ (- (System/nanoTime) t0) 1e9)]
      (println (format "{\"impl\": \"symja\", \"rep\": %d, \"seconds\": %.3f, \"value\": %.6e, \"unit\": \"expr-points/s\", \"cores\": %d, \"best\": %s}"
                       r dt (/ (* (double (count phenos)) n) dt) cores (apply max scores)))))
  (shutdown-agents))
