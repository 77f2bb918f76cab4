"""Writes the seeded synthetic population of a bench workload as Symja FullForm text, one expression per line, for
baseline/symja_time.clj (the reference timed on a JVM):  python scripts/dump_population.py c3 out.txt [count]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import closyr_b200 as cb  # noqa: E402
from bench import SEED, WORKLOADS  # noqa: E402

name, out = sys.argv[1], sys.argv[2]
pop, n, max_leafs, mix, desc = WORKLOADS[name]
count = int(sys.argv[3]) if len(sys.argv) > 3 else pop
b = cb.PopulationBuilder()
texts = b.generate(SEED, count, max_leafs, mix, want_texts=True)
with open(out, "w") as f:
    f.write("\n".join(texts) + "\n")
print("%d expressions of %s -> %s (points %d, max-leafs %d)" % (len(texts), desc, out, n, max_leafs))
