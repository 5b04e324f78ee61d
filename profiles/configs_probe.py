"""Supplementary measurements for the BASELINE.json configs that are not the bench line (run on the GPU box):
  configs[2]  performance/nof_sequence_filter: nucleotideMutationProfile(distance, querySequence) -> count()
  configs[4]  co-occurrence (BitmapAggregationNode) over six positions under the config-2 filter
on the 10 M-row config-2 table through the host API (expression text in, result out, everything inside the
timer), with the oracle port timed single-threaded on a bounded 2-chunk sample of the same generator beside
it (rows/s so that the two are comparable). Prints one line per query."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import types
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
config2_filter = f"(and {host_api.date_ranges_expression(rows, bench.SPAN_DAYS, bench.FROM_DAY, bench.TO_DAY, 0, len(sizes))} (bitmap lineage))"

sample_chunks = 2  # (the reference's Threshold is O(n k): distance 200 takes ~9 s per chunk on one core)
oracle_table, oracle_filter = bench.build_oracle_sample(sample_chunks)
oracle_rows = sample_chunks * 65536
from oracle import oracle as O
evolved, _parents = O.gen_evolved(oracle_table.columns[0][2], seed=42, generations=bench.GENERATIONS)


def timed(fn, n, warm=True):
    if warm:
        fn()
    torch.cuda.synchronize()
    begin = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - begin) / n, out


def report(label, device_seconds, device_result, oracle_seconds, oracle_result):
    print(f"{label:44s} device {device_seconds * 1e3:9.3f} ms ({rows / device_seconds:10.3e} rows/s, result {device_result}) | "
          f"oracle 1 core {oracle_seconds * 1e3:9.1f} ms on {oracle_rows} rows ({oracle_rows / oracle_seconds:10.3e} rows/s, result {oracle_result}) | "
          f"ratio {rows / device_seconds / (oracle_rows / oracle_seconds):8.1f}x", flush=True)


def count(t, expression):
    flt = t.filter(expression)
    cardinality = flt.cardinality
    flt.close()
    return cardinality


device_query = synthetic.sequence(synthetic.num_sequences - 1)
oracle_query = evolved[-1]
for distance in (0, 5, 50, 200):
    d_seconds, d_result = timed(lambda: count(table, f"(profile main {distance} seq {device_query})"), 5)
    o_seconds, o_result = timed(lambda: count(oracle_table, f"(profile main {distance} seq {oracle_query})"), 1, warm=False)
    report(f"config 3: mutationProfile(distance={distance})", d_seconds, d_result, o_seconds, o_result)

nof = "(n-of 3 0 (has-mut main 241) (has-mut main 3037) (has-mut main 14408) (sym-eq main 23403 G) (sym-eq main 100 A))"
d_seconds, d_result = timed(lambda: count(table, nof), 20)
o_seconds, o_result = timed(lambda: count(oracle_table, nof), 3)
report("config 3: 3-of-5 single-position tests", d_seconds, d_result, o_seconds, o_result)

dimensions = [("position", "main", p) for p in (5, 10, 20, 30, 40, 50)]
for label, d_filter, o_filter in (("no filter", None, None), ("config-2 filter", config2_filter, oracle_filter)):
    d_seconds, d_result = timed(lambda: table.bitmap_aggregation(dimensions, d_filter), 10)
    o_seconds, o_result = timed(lambda: oracle_table.bitmap_aggregation(dimensions, o_filter), 2)
    report(f"config 5: co-occurrence, 6 positions, {label}", d_seconds, f"{len(d_result)} combinations", o_seconds, f"{len(o_result)} combinations")
