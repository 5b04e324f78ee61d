"""configs[2] (nucleotideMutationProfile(distance, querySequence) -> count()) on the 10 M-row config-2 table, device
side only (no oracle leg: profiles/configs_probe.py has that): wall clock of the whole host call (expression text
in, cardinality out) and the host's own share (parse + rewrite + compile + lower, measured by
silo_host_filter_lower_timed on the same expression). Run on the GPU box; prints one line per distance."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lapis_silo_b200 import abi, host_api

GENOME_LENGTH, REFERENCE_SEED, GENERATIONS = 29903, 20200101, 5  # bench.py's table
rows = int(os.environ.get("ROWS", "10000000"))
synthetic = host_api.Synthetic(GENOME_LENGTH, REFERENCE_SEED, GENERATIONS)
sizes = host_api.dense_chunk_sizes(rows)
ctx = abi.Context(0)
table = host_api.HostTable(ctx, sizes)
table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_column(rows, 0, len(sizes), 16))
synthetic.release_column()
query = synthetic.sequence(synthetic.num_sequences - 1)
for distance in (0, 5, 50, 200):
    expression = f"(profile main {distance} seq {query})"
    times = []
    for _ in range(8):
        begin = time.perf_counter()
        flt = table.filter(expression)
        cardinality = flt.cardinality
        times.append(time.perf_counter() - begin)
        flt.close()
    lowered = min((table.lower_timed(expression) for _ in range(8)), key=lambda r: r["parse_us"] + r["rewrite_us"] + r["compile_us"] + r["lower_us"])
    host_us = lowered["parse_us"] + lowered["rewrite_us"] + lowered["compile_us"] + lowered["lower_us"]
    print(f"config 3: mutationProfile(distance={distance:3d}) query {min(times[1:]) * 1e3:7.3f} ms (median {sorted(times[1:])[3] * 1e3:7.3f}) "
          f"result {cardinality} | host lowering {host_us / 1e3:6.3f} ms (parse {lowered['parse_us']:.0f} rewrite {lowered['rewrite_us']:.0f} "
          f"compile {lowered['compile_us']:.0f} lower {lowered['lower_us']:.0f} us) | {rows / min(times[1:]):.3e} rows/s", flush=True)
table.close()
ctx.close()
