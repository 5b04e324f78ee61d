"""SILO_K1_STREAM_ONLY=4: how long the consumer warps of the container kernel wait for data.
Prints totals summed over all consumer warps (units of 64 cycles)."""
import os, sys
os.environ["SILO_K1_STREAM_ONLY"] = "4"
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
date = host_api.date_ranges_expression(rows, bench.SPAN_DAYS, bench.FROM_DAY, bench.TO_DAY, 0, len(sizes))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
counts = torch.zeros(16 * bench.GENOME_LENGTH, dtype=torch.int32, device="cuda")
for label, expression in (("config2", f"(and {date} (bitmap lineage))"), ("all chunks", "(bitmap lineage)")):
    prepared = table.prepare(expression)
    for _ in range(3):
        prepared.run_async(stream.cuda_stream); table.mutation_counts_async(0, prepared, counts.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    raw = counts.cpu().numpy().view(np.uint32).reshape(16, bench.GENOME_LENGTH)[15].astype(np.int64)
    n_ctas = int(os.environ.get("K1_CTAS", "296"))
    exits, smids, stages, ends = (raw[at:at + n_ctas] for at in (1024, 2048, 3072, 4096))
    base = min(exits.min(), ends.min())
    exits, ends = (exits - base) / 1e3, (ends - base) / 1e3
    slow = stages < stages.mean()
    print(f"{label}: per-CTA wall clock of the last launch (us after the earliest exit): producers out of work, deciles "
          f"{np.round(np.percentile(exits, [0, 10, 25, 50, 75, 90, 100]), 1)}; consumers done, deciles {np.round(np.percentile(ends, [0, 10, 25, 50, 75, 90, 100]), 1)}; "
          f"stages per CTA min {stages.min()} mean {stages.mean():.1f} max {stages.max()}, correlation(end, stages) {np.corrcoef(ends, stages)[0, 1]:.2f}; "
          f"CTAs with fewer stages than the mean end at median {np.median(ends[slow]):.1f}, the others at {np.median(ends[~slow]):.1f}; "
          f"%smid: {len(set(smids.tolist()))} distinct, {smids.min()}..{smids.max()}, at most {np.bincount(smids).max()} CTAs each")
    probe = raw[:8].astype(np.float64)
    total, waited, first, visits, warps, producer_waited, producer_total, producers = probe
    s = table.stats()
    print(f"{label:12s} K1 {s.last_counts_kernel_ms*1e3:7.1f} us | consumer warps {int(warps)}: mean lifetime {total/warps*64/1965:7.1f} us, "
          f"waiting for data {waited/warps*64/1965:6.1f} us ({100*waited/total:4.1f} %), first stage after {first/warps*64/1965:5.1f} us "
          f"({100*first/total:4.1f} %), stage visits per warp {visits/warps:6.1f} | producers {int(producers)}: lifetime "
          f"{producer_total/producers*64/1965:6.1f} us, waiting for a free stage {producer_waited/producers*64/1965:6.1f} us "
          f"({100*producer_waited/producer_total:4.1f} %)")
