"""Where an end-to-end step spends its time (run on the GPU box; prints a table to stdout)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
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
ancestor = next(e for e in range(synthetic.num_sequences) if synthetic.generation(e) == 2)
table.register_bitmap("lineage", synthetic.lineage_bitmap(ancestor, rows, 0, len(sizes)))
expression = f"(and {host_api.date_ranges_expression(rows, bench.SPAN_DAYS, bench.FROM_DAY, bench.TO_DAY, 0, len(sizes))} (bitmap lineage))"

def timed(label, fn, n=20):
    fn(); fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    print(f"{label:48s} {(time.perf_counter() - t) / n * 1e3:8.3f} ms")
    return out

flt = timed("filter(expression): parse+compile+H2D+kernel+D2H", lambda: table.filter(expression))
timed("mutation_counts(filter): kernels + D2H counts", lambda: table.mutation_counts("main", flt))
counts = table.mutation_counts("main", flt)
timed("mutation_rows_from_counts: host thresholding", lambda: table.mutation_rows_from_counts("main", counts, 0.05))
timed("mutations(): whole MutationsNode", lambda: table.mutations(["main"], expression, 0.05))
acc = {}
for _ in range(20):
    table.mutations(["main"], expression, 0.05)
    for key, value in table.last_query_profile().items():
        acc[key] = acc.get(key, 0.0) + value / 20
print("C++ phases of mutations() [us]:", {k: round(v, 1) for k, v in acc.items()}, "sum", round(sum(acc.values()), 1))
timed("mutation_counts(None): full filter path", lambda: table.mutation_counts("main", None))
s = table.stats()
print("stats:", s.containers, s.algorithmic_bytes, s.counts_kernel_bytes, s.last_counts_kernel_ms, s.last_total_ms, s.timed_calls)
