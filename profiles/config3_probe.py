"""BASELINE.json configs[2] (performance/nof_sequence_filter): nucleotideMutationProfile(distance, querySequence)
-> count() on 10 M full-length rows, plus single-position filters. Prints wall time per query through the
host API (expression text in, cardinality out) and the filter cardinalities."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lapis_silo_b200 import abi, host_api
rows = int(os.environ.get("ROWS", "10000000"))
synthetic = host_api.Synthetic(bench.GENOME_LENGTH, bench.REFERENCE_SEED, bench.GENERATIONS)
sizes = host_api.dense_chunk_sizes(rows)
ctx = abi.Context(0)
table = host_api.HostTable(ctx, sizes)
table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_column(rows, 0, len(sizes), 16))
synthetic.release_column()
query = synthetic.sequence(synthetic.num_sequences - 1)

def timed(label, expression, n=10):
    flt = table.filter(expression); flt.close()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        flt = table.filter(expression)
        cardinality = flt.cardinality
        flt.close()
    torch.cuda.synchronize()
    print(f"{label:52s} {(time.perf_counter() - t) / n * 1e3:8.3f} ms   |filter| = {cardinality}")

for distance in (0, 5, 50, 200):
    timed(f"mutationProfile(distance={distance}, querySequence)", f"(profile main {distance} seq {query})")
timed("3-of-5 single-position tests", "(n-of 3 0 (has-mut main 241) (has-mut main 3037) (has-mut main 14408) (sym-eq main 23403 G) (sym-eq main 100 A))")
timed("symbolEquals at one position", "(sym-eq main 23403 G)", 50)
timed("hasMutation at one position", "(has-mut main 14408)", 50)
timed("5-symbol OR at one position", "(or (sym-eq main 77 A) (sym-eq main 77 C) (sym-eq main 77 G) (sym-eq main 77 T) (sym-eq main 77 -))", 50)

# where a profile query spends its time on the host
import time as _time
expression = f"(profile main 5 seq {query})"
for _ in range(3):
    t0 = _time.perf_counter()
    flt = table.filter(expression)
    t1 = _time.perf_counter()
    profile = table.last_query_profile()
    flt.close()
print("profile query: total %.2f ms; rewrite+compile+lower %.2f ms; silo_gpu_filter_eval %.2f ms; parse+rest %.2f ms" % (
    (t1 - t0) * 1e3, profile["compile_us"] / 1e3, profile["filter_us"] / 1e3,
    (t1 - t0) * 1e3 - profile["compile_us"] / 1e3 - profile["filter_us"] / 1e3))
