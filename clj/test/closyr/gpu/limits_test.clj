(ns closyr.gpu.limits-test
  "Row N3: the schema / CLI bounds are parameters, the GPU options reach run-config.
   Run inside a closyr checkout with clj/patches/0001 applied and clj/ on the source path:
   lein test :only closyr.gpu.limits-test   (start the JVM with the -D properties below to see the raised bounds)."
  (:require
    [clojure.test :refer :all]
    [clojure.tools.cli :as cli]
    [closyr.core :as core]
    [closyr.util.spec :as specs]
    [malli.core :as m]))

(deftest bounds-default-to-the-reference-values
  (when-not (System/getProperty "closyr.max-points")
    (is (= 10000 specs/max-points-count))
    (is (= 100000 specs/max-population-count))
    (is (= 100 specs/max-cli-points-count))
    (is (= 10000 specs/max-cli-population-count))))

(deftest bounds-follow-the-system-properties
  (when-let [p (System/getProperty "closyr.max-points")]
    (is (= (Long/parseLong p) specs/max-points-count))
    (is (m/validate @#'specs/PointsCount (Long/parseLong p)))
    (is (not (m/validate @#'specs/PointsCount (inc (Long/parseLong p)))))))

(deftest population-schema-uses-the-bound
  (is (m/validate @#'specs/PopulationCount specs/max-population-count))
  (is (not (m/validate @#'specs/PopulationCount (inc specs/max-population-count)))))

(deftest cli-accepts-the-gpu-options
  (let [{:keys [options errors]} (cli/parse-opts ["-t" "-p" "25" "-G" "8" "--shard-points"] @#'core/cli-options)]
    (is (empty? errors))
    (is (= 8 (:gpus options)))
    (is (true? (:shard-points? options))))
  (let [{:keys [options errors]} (cli/parse-opts ["-t"] @#'core/cli-options)]
    (is (empty? errors))
    (is (= 0 (:gpus options)))
    (is (false? (:shard-points? options))))
  (is (seq (:errors (cli/parse-opts ["-G" "9"] @#'core/cli-options)))))

(deftest cli-population-bound-is-the-parameter
  (is (empty? (:errors (cli/parse-opts ["-p" (str specs/max-cli-population-count)] @#'core/cli-options))))
  (is (seq (:errors (cli/parse-opts ["-p" (str (inc specs/max-cli-population-count))] @#'core/cli-options)))))

(deftest run-config-schema-accepts-the-gpu-keys
  (let [schema (m/schema @#'specs/SolverRunConfig)
        keys*  (set (map first (m/children schema)))]
    (is (contains? keys* :gpus))
    (is (contains? keys* :shard-points?))))
