(ns closyr.gpu.ffi
  "Java FFM (Panama, JDK >= 22) binding of include/closyr_b200.h.  One context per run: it owns the CUDA
   streams and device buffers and keeps xs/ys resident across generations (replaces the IExpr[] marshalling
   of src/closyr/ops/common.clj:30-76 and `->run-args`, src/closyr/symbolic_regression.clj:355-383)."
  (:import
    (java.lang.foreign Arena FunctionDescriptor Linker MemorySegment SymbolLookup ValueLayout)
    (java.lang.invoke MethodHandle)))

(set! *warn-on-reflection* true)

(def ^:private ^Linker linker (Linker/nativeLinker))
(defonce ^:private lookup* (delay (SymbolLookup/libraryLookup "libclosyr_b200.so" (Arena/global))))

(defn- ^MethodHandle handle [^String sym ^FunctionDescriptor fd]
  (.downcallHandle linker (.orElseThrow (.find ^SymbolLookup @lookup* sym)) fd (make-array java.lang.foreign.Linker$Option 0)))

(def ^:private I ValueLayout/JAVA_INT)
(def ^:private L ValueLayout/JAVA_LONG)
(def ^:private P ValueLayout/ADDRESS)
(defn- fd [ret & args] (FunctionDescriptor/of ret (into-array java.lang.foreign.MemoryLayout args)))

(def ^:private h-ctx-create   (delay (handle "closyr_ctx_create" (fd I P I I P))))
(def ^:private h-ctx-destroy  (delay (handle "closyr_ctx_destroy" (FunctionDescriptor/ofVoid (into-array java.lang.foreign.MemoryLayout [P])))))
(def ^:private h-set-points   (delay (handle "closyr_set_points" (fd I P P P L))))
(def ^:private h-score-batch  (delay (handle "closyr_score_batch" (fd I P P I P P))))
(def ^:private h-eval-batch   (delay (handle "closyr_eval_batch" (fd I P P P L P P))))
(def ^:private h-rank         (delay (handle "closyr_rank_resident" (fd I P L P))))
(def ^:private h-last-error   (delay (handle "closyr_last_error" (fd P P))))

(defn- check! [^MemorySegment ctx rc]
  (when-not (zero? (long rc))
    (let [^MemorySegment msg (.invokeWithArguments ^MethodHandle @h-last-error [ctx])]
      (throw (ex-info (str "closyr_b200: " (.getString (.reinterpret msg 4096) 0)) {:code rc})))))

(defn create-context
  "n-gpus devices 0..n-1 of this box; flags 0 = population sharding, 1 = point sharding."
  [n-gpus flags]
  (with-open [a (Arena/ofConfined)]
    (let [out (.allocate a P)
          rc (.invokeWithArguments ^MethodHandle @h-ctx-create [MemorySegment/NULL (int n-gpus) (int flags) out])]
      (check! MemorySegment/NULL rc)
      (.get out P 0))))

(defn destroy-context [^MemorySegment ctx]
  (.invokeWithArguments ^MethodHandle @h-ctx-destroy [ctx]))

(defn set-points!
  "Copies xs / ys (double arrays) to every device of the context."
  [^MemorySegment ctx ^doubles xs ^doubles ys]
  (with-open [a (Arena/ofConfined)]
    (let [sx (.allocateFrom a ValueLayout/JAVA_DOUBLE xs)
          sy (.allocateFrom a ValueLayout/JAVA_DOUBLE ys)]
      (check! ctx (.invokeWithArguments ^MethodHandle @h-set-points [ctx sx sy (long (alength xs))])))))

(defn population-segment
  "Packs compiled programs (closyr.gpu.compile/compile-expr maps) into the flat struct-of-arrays buffer and the
   closyr_population struct that points into it.  The arena owns all of it."
  [^Arena a programs]
  (let [n (count programs)
        ncode (reduce + (map #(alength ^ints (:code %)) programs))
        nk (reduce + (map #(alength ^doubles (:consts %)) programs))
        code (.allocate a (* 4 (max 1 ncode)) 16) consts (.allocate a (* 8 (max 1 nk)) 16)
        code-off (.allocate a (* 4 (inc n)) 4) const-off (.allocate a (* 4 (inc n)) 4)
        meta (.allocate a (* 4 (max 1 n)) 4) aux (.allocate a (* 4 (max 1 n)) 4)
        pop (.allocate a 56 8)]
    (loop [i 0 co 0 ko 0 ps (seq programs)]
      (.setAtIndex code-off I (long i) (int co))
      (.setAtIndex const-off I (long i) (int ko))
      (when ps
        (let [{:keys [^ints code-arr ^doubles consts-arr] :as p} (assoc (first ps) :code-arr (:code (first ps)) :consts-arr (:consts (first ps)))]
          (MemorySegment/copy code-arr 0 code I (* 4 (long co)) (alength code-arr))
          (MemorySegment/copy consts-arr 0 consts ValueLayout/JAVA_DOUBLE (* 8 (long ko)) (alength consts-arr))
          (.setAtIndex meta I (long i) (int (:meta p)))
          (.setAtIndex aux I (long i) (int (:aux p)))
          (recur (inc i) (+ co (alength code-arr)) (+ ko (alength consts-arr)) (next ps)))))
    (.set pop L 0 (long n))
    (.set pop P 8 code) (.set pop P 16 code-off) (.set pop P 24 consts)
    (.set pop P 32 const-off) (.set pop P 40 meta) (.set pop P 48 aux)
    pop))

(defn score-batch!
  "Scores a whole generation; returns a double-array of n scores (failures are min-score, as ops/score-fn)."
  [^MemorySegment ctx programs max-leafs]
  (with-open [a (Arena/ofConfined)]
    (let [n (count programs)
          pop (population-segment a programs)
          out (.allocate a (* 8 (max 1 n)) 8)
          rc (.invokeWithArguments ^MethodHandle @h-score-batch [ctx pop (int max-leafs) out MemorySegment/NULL])]
      (check! ctx rc)
      (.toArray (.asSlice out 0 (* 8 n)) ValueLayout/JAVA_DOUBLE))))
