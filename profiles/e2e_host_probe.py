"""Where the host side of an end-to-end Mutations query goes (run on the GPU box): the C++ phases the host layer
records (silo_host_last_query_profile), the ctypes call as a whole, and the Python unpacking of the record batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
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
N = 300
for _ in range(50):
    table.mutations_columns(["main"], expression, 0.05)
acc = {}
started = time.perf_counter()
for _ in range(N):
    table.mutations_columns(["main"], expression, 0.05)
    for key, value in table.last_query_profile().items():
        acc[key] = acc.get(key, 0.0) + value / N
whole = (time.perf_counter() - started) / N * 1e6
print("C++ phases [us]:", {k: round(v, 1) for k, v in acc.items()}, "sum", round(sum(acc.values()), 1))
started = time.perf_counter()
for _ in range(N):
    table.mutations_columns(["main"], expression, 0.05)
print("mutations_columns, whole [us]:", round((time.perf_counter() - started) / N * 1e6, 1), "(with the profile read:", round(whole, 1), ")")
lib = host_api.lib()
names = (C.c_char_p * 1)(b"main")
n_rows, n_names, needed = table._packed_out
encoded = expression.encode()
started = time.perf_counter()
for _ in range(N):
    lib.silo_host_mutations_packed(table._h, encoded, names, 1, 0.05, table._packed.ctypes.data, table._packed.nbytes, n_rows, n_names, needed)
print("the ctypes call alone [us]:", round((time.perf_counter() - started) / N * 1e6, 1))
started = time.perf_counter()
for _ in range(N):
    expression.encode()
print("expression.encode [us]:", round((time.perf_counter() - started) / N * 1e6, 2))
batch = table._packed[:needed.value]
started = time.perf_counter()
for _ in range(N):
    host_api._unpack_record_batch(batch.copy(), int(n_rows.value), int(n_names.value))
print("unpack [us]:", round((time.perf_counter() - started) / N * 1e6, 1))
