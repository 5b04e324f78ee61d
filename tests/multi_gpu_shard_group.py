"""Run under torchrun (one process per GPU): the library's shard group across REAL devices and processes. Every rank
uploads its interleaved shard of a synthetic table, the handles travel through torch.distributed, every rank enqueues
the sharded Mutations query, rank 0 collects and compares with the oracle on the whole table. Started by
tests/test_multi_gpu.py (skipped with fewer than two GPUs) and by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_shard_group.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lapis_silo_b200 import abi, host_api  # noqa: E402


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")  # the transport of the handles only; no NCCL anywhere in this test
    total_rows, length = int(os.environ.get("ROWS", 6 * 65536 + 999)), 1500
    synthetic = host_api.Synthetic(genome_length=length, reference_seed=7, generations=5)
    sizes = host_api.dense_chunk_sizes(total_rows)
    first, n_chunks, stride = host_api.interleaved_shard(len(sizes), world, rank)
    ctx = abi.Context(local_rank)
    table = host_api.HostTable(ctx, host_api.shard_chunk_sizes(total_rows, first, n_chunks, stride))
    table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_column(total_rows, first, n_chunks, 4, stride))
    ancestor = next(e for e in range(synthetic.num_sequences) if synthetic.generation(e) == 2)
    table.register_bitmap("lineage", synthetic.lineage_bitmap(ancestor, total_rows, first, n_chunks, stride))
    expression = f"(and {host_api.date_ranges_expression(total_rows, 1095, 200, 800, first, n_chunks, stride)} (bitmap lineage))"

    handle = table.shard_group_create("main", rank, world)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    table.shard_group_connect(handles)
    dist.barrier()

    stream = torch.cuda.Stream()
    results = []
    queries = [(expression, 0.05), (None, 0.05), (expression, 0.0), (expression, 0.3), (None, 0.0), (expression, 0.05), (expression, 0.05),
               (expression, 0.05), (expression, 0.05), (expression, 0.2), (expression, 0.05)]
    with torch.cuda.stream(stream):
        summed = torch.zeros(16 * length, dtype=torch.int32, device="cuda")
        for index, (text, min_proportion) in enumerate(queries):
            if rank == 0 and index >= 3:  # the root's two halves as one call (a replayed graph from the third time on)
                columns, cardinality = table.sharded_query("main", text, min_proportion, summed.data_ptr())
                results.append((host_api.rows_from_columns(columns), cardinality, summed.cpu().numpy().view(np.uint32).reshape(16, length)[:5].copy()))
                continue
            table.sharded_enqueue("main", text, stream.cuda_stream)
            if rank == 0:
                columns, cardinality = table.sharded_collect("main", min_proportion, stream.cuda_stream, summed.data_ptr())
                results.append((host_api.rows_from_columns(columns), cardinality, summed.cpu().numpy().view(np.uint32).reshape(16, length)[:5].copy()))
        stream.synchronize()
    # BitmapAggregationNode over the same shards: every rank's (key, count) list to rank 0, summed per key there
    dimensions = [("position", "main", p) for p in (3, 700, 1499)]
    aggregated = []
    for text in (expression, None):
        shard_result = table.bitmap_aggregation_shard(dimensions, text)
        parts = [None] * world
        dist.all_gather_object(parts, shard_result)
        aggregated.append(table.bitmap_aggregation_merge(dimensions, parts) if rank == 0 else None)
    dist.barrier()
    if rank == 0:
        from oracle import oracle as O
        whole_sizes = sizes
        oracle_table = O.Table()
        oracle_table.set_layout(*whole_sizes)
        oracle_table.import_column("main", O.NUCLEOTIDE, synthetic.reference, synthetic.build_column(total_rows, 0, len(whole_sizes), 4))
        n_sequences = synthetic.num_sequences
        in_lineage = np.zeros(n_sequences, dtype=bool)
        in_lineage[ancestor] = True
        for e in range(ancestor + 1, n_sequences):
            in_lineage[e] = in_lineage[synthetic.parent(e)]
        oracle_table.register_bitmap("lineage", np.flatnonzero(in_lineage[np.arange(total_rows) % n_sequences]).astype(np.uint32))
        whole_expression = f"(and {host_api.date_ranges_expression(total_rows, 1095, 200, 800, 0, len(whole_sizes))} (bitmap lineage))"
        for (text, min_proportion), (rows, cardinality, counts) in zip(queries, results):
            flt = oracle_table.filter(whole_expression) if text is not None else None
            want_counts = oracle_table.mutation_counts("main", flt)
            assert cardinality == (flt.cardinality if flt is not None else total_rows), (cardinality, text is None)
            np.testing.assert_array_equal(counts, want_counts[:5])
            assert rows == oracle_table.mutation_rows("main", want_counts, min_proportion)
        for text, rows in zip((whole_expression, None), aggregated):
            assert rows == oracle_table.bitmap_aggregation(dimensions, text)
        print(f"multi-GPU shard group ok: {world} ranks, {len(queries)} queries, results identical to the oracle", flush=True)
    table.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
