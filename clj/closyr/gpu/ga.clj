(ns closyr.gpu.ga
  "Two-phase `evolve` (row N1 of the scope table): what closyr.ga/evolve (src/closyr/ga.clj:58-88) does, with the
   scoring pulled out of `compete` so that a generation costs ONE device call instead of one per child.

   Phase 1  score whoever has no score (generation 0: everybody), shuffle, walk the pairs exactly as `compete`
            does (src/closyr/ga.clj:32-46) but only CREATE the child of a winning e1.
   Phase 2  score all children with `closyr.gpu.ops/with-scores`, put them back into their pairs.

   `score-fn` is a pure function of (expr, xs, ys, max-leafs) apart from the stats counters (ops.clj:66-69, 110),
   and the draw order of the shared generator is already scheduling-dependent under `pmap`, so the result is a
   valid run of the same GA.  The executable specification of this loop is closyr_b200/csrc/ga.cpp
   (closyr_ga_evolve), which is tested against the reference's invariants in tests/test_ga.py."
  (:refer-clojure :exclude [rand rand-int rand-nth shuffle])
  (:require
    [closyr.gpu.ops :as gpu-ops]
    [closyr.util.log :as log]
    [closyr.util.prng :refer :all]))

(def ^:private new-phen-modifier-sampler
  ;; 4 / 5 chance of mutation instead of crossover (src/closyr/ga.clj:20-22)
  [true true true true false])

(defn- propose
  "`compete` without the scoring: [pair-score [e1 e2-or-unscored-child]]."
  [{:keys [mutation-fn crossover-fn]} [{e1-score :score :as e1} {e2-score :score :as e2}]]
  (if (nil? e2)
    [e1-score [e1]]
    (let [new-e-fn (if (rand-nth new-phen-modifier-sampler) mutation-fn crossover-fn)
          next-e   (if (>= (double e1-score) (double e2-score))
                     (dissoc (new-e-fn e1 e2) :score)
                     e2)]
      [(+ (double e1-score) (double e2-score)) [e1 next-e]])))

(defn evolve
  "Same contract as closyr.ga/evolve: takes and returns the GA map {:pop :score-fn :mutation-fn :crossover-fn},
   adds :pop-scores.  `run-args` must come from closyr.gpu.ops/->gpu-run-args."
  [run-args run-config {:keys [pop] :as config}]
  (try
    (let [scored     (gpu-ops/with-scores run-args run-config pop)          ; phase 0: generation 0, or nothing to do
          pairs      (partition-all 2 (shuffle scored))
          proposals  (mapv (partial propose config) pairs)                   ; host: mutation / crossover only
          new-pop    (gpu-ops/with-scores run-args run-config                ; ONE device call for all children
                                          (vec (mapcat second proposals)))]
      (merge config {:pop new-pop :pop-scores (mapv first proposals)}))
    (catch Exception e
      (log/error "Err in gpu evolve: " e)
      (throw e))))
