"""ncu target for configs[2]: ONE nucleotideMutationProfile(distance 5, querySequence) filter on the 10 M-row bench table
(after one warm-up query)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for _ in range(2):
    flt = table.filter(f"(profile main 5 seq {query})")
    print(flt.cardinality)
    flt.close()
