"""K1 timing probe: prints the dominant kernel's time / achieved GB/s for the bench workload."""
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
for label, expression in (("config2 date&lineage", f"(and {date} (bitmap lineage))"), ("lineage only (all chunks)", "(bitmap lineage)"),
                          ("not lineage (dense, all chunks)", "(not (bitmap lineage))")):
    prepared = table.prepare(expression)
    for _ in range(3):
        prepared.run_async(stream.cuda_stream); table.mutation_counts_async(0, prepared, counts.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize(); table.stats()
    for _ in range(10):
        prepared.run_async(stream.cuda_stream); table.mutation_counts_async(0, prepared, counts.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize(); s = table.stats()
    print(f"{label:34s} |filter|={prepared.cardinality():9d} K1 {s.last_counts_kernel_ms*1e3:8.1f} us  {s.counts_kernel_bytes/1e6:8.1f} MB "
          f"{s.counts_kernel_bytes/s.last_counts_kernel_ms/1e6:8.1f} GB/s  whole {s.last_total_ms*1e3:8.1f} us")
