import os, sys
sys.path.insert(0, '/root/repo')
import bench
from lapis_silo_b200 import abi, host_api
rows = 10_000_000
synthetic = host_api.Synthetic(bench.GENOME_LENGTH, bench.REFERENCE_SEED, bench.GENERATIONS)
sizes = host_api.dense_chunk_sizes(rows)
ctx = abi.Context(0)
table = host_api.HostTable(ctx, sizes)
table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_column(rows, 0, len(sizes), 16))
synthetic.release_column()
ancestor = next(e for e in range(synthetic.num_sequences) if synthetic.generation(e) == 2)
table.register_bitmap("lineage", synthetic.lineage_bitmap(ancestor, rows, 0, len(sizes)))
expression = f"(and {host_api.date_ranges_expression(rows, bench.SPAN_DAYS, bench.FROM_DAY, bench.TO_DAY, 0, len(sizes))} (bitmap lineage))"
import time
for _ in range(200):
    table.mutations_columns(["main"], expression, 0.05)
t=time.perf_counter()
for _ in range(200):
    table.mutations_columns(["main"], expression, 0.05)
print("e2e us", (time.perf_counter()-t)/200*1e6)
