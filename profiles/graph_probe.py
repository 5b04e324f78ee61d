"""Does capturing the device-resident step into a CUDA graph shorten it? (launch gaps between the
dependent kernels of a query). Prints eager vs graph time per step."""
import os, sys
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
ancestor = next(e for e in range(synthetic.num_sequences) if synthetic.generation(e) == 2)
table.register_bitmap("lineage", synthetic.lineage_bitmap(ancestor, rows, 0, len(sizes)))
date = host_api.date_ranges_expression(rows, bench.SPAN_DAYS, bench.FROM_DAY, bench.TO_DAY, 0, len(sizes))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
counts = torch.zeros(16 * bench.GENOME_LENGTH, dtype=torch.int32, device="cuda")
prepared = table.prepare(f"(and {date} (bitmap lineage))")
def step():
    prepared.run_async(stream.cuda_stream); table.mutation_counts_async(0, prepared, counts.data_ptr(), stream.cuda_stream)
def timed(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(n): fn()
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
print(f"eager: {timed(step, 20):.1f} us/step")
ref = counts.clone()
STEPS = 20
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=stream):
    for _ in range(STEPS): step()
print(f"graph of {STEPS} steps: {timed(graph.replay, 5) / STEPS:.1f} us/step")
assert torch.equal(ref, counts)
