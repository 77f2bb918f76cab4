(ns closyr.gpu.ops
  "Drop-in scoring entry points backed by the B200 path.

   `score-fn` has the signature and the error behaviour of closyr.ops/score-fn (src/closyr/ops.clj:119-136):
   [run-args run-config pheno] -> number, every failure -> ops/min-score.  `score-batch` is the batched
   form a two-phase `evolve` uses (generate all children, one call, attach scores): `compete` scores children
   inline today (src/closyr/ga.clj:32-46), which is one device round trip per child."
  (:require
    [closyr.gpu.compile :as gc]
    [closyr.gpu.ffi :as ffi]
    [closyr.ops :as ops]
    [closyr.ops.common :as ops-common]
    [closyr.util.log :as log]))

(defn ->gpu-run-args
  "Extends the map built by `->run-args` (src/closyr/symbolic_regression.clj:355-383) with a device context
   that holds xs / ys; call once per run."
  [{:keys [input-xs-vec input-ys-vec] :as run-args} {:keys [gpus shard-points?] :or {gpus 1}}]
  (let [ctx (ffi/create-context gpus (if shard-points? 1 0))]
    (ffi/set-points! ctx (double-array input-xs-vec) (double-array input-ys-vec))
    (assoc run-args :gpu-ctx ctx)))

(defn score-batch
  "Scores every phenotype of `phenos` in one device call; returns a vector of numbers in the same order."
  [{:keys [gpu-ctx]} {:keys [max-leafs]} phenos]
  (try
    (let [programs (mapv (fn [{:keys [expr sym]}] (gc/compile-expr expr sym)) phenos)]
      (vec (ffi/score-batch! gpu-ctx programs (or max-leafs ops/default-max-leafs))))
    (catch Exception e
      (log/error "Err in gpu score-batch: " (.getMessage e))
      (vec (repeat (count phenos) ops/min-score)))))

(defn score-fn
  "Same contract as ops/score-fn; a batch of one."
  [run-args run-config pheno]
  (first (score-batch run-args run-config [pheno])))

(defn with-scores
  "Two-phase replacement of ga/with-score (src/closyr/ga.clj:25-29) for a whole population: one device call for
   every phenotype that has no :score yet.  Matched by POSITION, so nothing depends on how mutation-fn / crossover-fn
   hand out :id values."
  [run-args run-config pop]
  (let [pop    (vec pop)
        idx    (vec (keep-indexed (fn [i p] (when-not (:score p) i)) pop))
        scores (if (seq idx) (score-batch run-args run-config (mapv pop idx)) [])]
    (reduce (fn [acc [i s]] (assoc-in acc [i :score] s)) pop (map vector idx scores))))
