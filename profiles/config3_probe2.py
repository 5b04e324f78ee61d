"""configs[2] probe: nucleotideMutationProfile(distance, querySequence) -> count() on the 10 M-row bench table through the
host API, device time per query with the sweep kernel on and off."""
import os, sys, time
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
for label, min_pieces in (("sweep", 0), ("interpreter walk", 2 ** 63)):
    table.set_option("sweep_min_pieces", min_pieces)
    for distance in (0, 5, 50, 200):
        text = f"(profile main {distance} seq {query})"
        table.filter(text).close()
        started = time.perf_counter()
        n = 5
        for _ in range(n):
            flt = table.filter(text)
            cardinality = flt.cardinality
            flt.close()
        elapsed = (time.perf_counter() - started) / n
        profile = table.last_query_profile()
        print(f"{label:18s} distance {distance:3d}: |filter| = {cardinality:8d}  {elapsed * 1e3:7.3f} ms per query (host lowering included) "
              f"lowering {sum(v for k, v in table.lower_timed(text).items() if k.endswith('_us')) / 1e3:6.3f} ms; last call: "
              f"rewrite+compile+lower {profile['compile_us'] / 1e3:.3f} ms, silo_gpu_filter_eval {profile['filter_us'] / 1e3:.3f} ms")
