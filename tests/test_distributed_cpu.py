"""The N>1 path's host logic at world_size 2 on gloo (no GPU): chunk-range partitioning, per-shard
filter expressions in global row coordinates, the u32-counts-as-i32 all-reduce, cardinality sum and
thresholding of the reduced counts (bench.py runs the same steps with NCCL).

The per-shard "device" is played by the oracle evaluating the shard's expression on the whole table:
counts are additive over rows and a shard expression only selects rows of its own chunk range, so
the shard's contribution is exactly what a rank's device table would produce (the GPU twin of this
test is tests/test_gpu_parity.py::test_synthetic_shards_sum_to_the_whole)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TOTAL_ROWS = 5 * 65536 + 4321
GENOME_LENGTH = 700
MIN_PROPORTION = 0.05


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def build_inputs():
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    synthetic = host_api.Synthetic(genome_length=GENOME_LENGTH, reference_seed=5, generations=5)
    sizes = host_api.dense_chunk_sizes(TOTAL_ROWS)
    table = O.Table()
    table.set_layout(*sizes)
    table.import_column("main", O.NUCLEOTIDE, synthetic.reference, synthetic.build_column(TOTAL_ROWS, 0, len(sizes), 2))
    n_sequences = synthetic.num_sequences
    ancestor = next(e for e in range(n_sequences) if synthetic.generation(e) == 2)
    in_lineage = np.zeros(n_sequences, dtype=bool)
    in_lineage[ancestor] = True
    for e in range(ancestor + 1, n_sequences):
        in_lineage[e] = in_lineage[synthetic.parent(e)]
    table.register_bitmap("lineage", np.flatnonzero(in_lineage[np.arange(TOTAL_ROWS) % n_sequences]).tolist())
    return host_api, synthetic, sizes, table


def shard_expression(host_api, first, n_chunks):
    return f"(and {host_api.date_ranges_expression(TOTAL_ROWS, 1095, 150, 900, first, n_chunks)} (bitmap lineage))"


def worker(rank: int, world_size: int, port: int, result_path: str) -> None:
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        host_api, synthetic, sizes, table = build_inputs()
        # uneven weights: the partition balances payload bytes, not chunk counts
        weights = [3, 1, 1, 1, 3, 1]
        bounds = host_api.partition_chunks(weights, world_size)
        assert bounds[0] == 0 and bounds[-1] == len(sizes) and len(bounds) == world_size + 1
        first, n_chunks = bounds[rank], bounds[rank + 1] - bounds[rank]
        assert n_chunks > 0

        flt = table.filter(shard_expression(host_api, first, n_chunks))
        ids = flt.ids()
        assert ((ids >> 16) >= first).all() and ((ids >> 16) < first + n_chunks).all(), "a shard selects only its own rows"
        counts = table.mutation_counts("main", flt)  # u32[16][L] of this shard
        # one extra element past 2^31 on every rank: the i32 view must still sum modulo 2^32
        padded = np.concatenate([counts.reshape(-1), np.array([0xC0000000 + rank], dtype=np.uint32)])
        reduced = torch.from_numpy(padded.view(np.int32).copy())
        dist.all_reduce(reduced)
        cardinality = torch.tensor([flt.cardinality], dtype=torch.int64)
        dist.all_reduce(cardinality)

        reduced_u32 = reduced.numpy().view(np.uint32)
        assert int(reduced_u32[-1]) == (sum(0xC0000000 + r for r in range(world_size)) & 0xFFFFFFFF)
        whole_filter = table.filter(shard_expression(host_api, 0, len(sizes)))
        assert int(cardinality.item()) == whole_filter.cardinality > 0
        whole_counts = table.mutation_counts("main", whole_filter)
        np.testing.assert_array_equal(reduced_u32[:-1].reshape(whole_counts.shape), whole_counts)
        if rank == 0:  # only rank 0 thresholds and returns rows
            rows = table.mutation_rows("main", reduced_u32[:-1].reshape(whole_counts.shape), MIN_PROPORTION)
            assert rows == table.mutations("main", shard_expression(host_api, 0, len(sizes)), MIN_PROPORTION)
            assert len(rows) > 0
            with open(result_path, "w") as out:
                out.write(f"ok {len(rows)} {int(cardinality.item())}")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_and_count_allreduce(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    result_path = str(tmp_path / "result.txt")
    mp.spawn(worker, args=(2, free_port(), result_path), nprocs=2, join=True)
    with open(result_path) as result:
        assert result.read().startswith("ok ")


def test_partition_covers_all_chunks_without_gaps():
    sys.path.insert(0, ROOT)
    from lapis_silo_b200 import host_api
    rng = np.random.default_rng(0)
    for n_chunks in (1, 2, 7, 153, 1221):
        weights = [int(w) for w in rng.integers(1, 1000, n_chunks)]
        for ranks in (1, 2, 4, 8):
            bounds = host_api.partition_chunks(weights, ranks)
            assert bounds[0] == 0 and bounds[-1] == n_chunks
            assert all(a <= b for a, b in zip(bounds, bounds[1:]))
            if n_chunks >= ranks:
                assert all(a < b for a, b in zip(bounds, bounds[1:])), "no rank may be left without chunks"
                loads = [sum(weights[a:b]) for a, b in zip(bounds, bounds[1:])]
                assert max(loads) <= sum(weights) / ranks + max(weights)


def archive_worker(rank: int, world_size: int, port: int, result_path: str) -> None:
    """every rank reads the SAME `.silo` bytes with the product's C++ loader and keeps its own chunk range
    (host/silo_loader.cpp shardOf); what the ranks hold together must be the whole table: row counts add up,
    and the stored-symbol counts of the unfiltered Mutations action -- the sums of the shard's container
    cardinalities per (symbol, position) -- all-reduce to the oracle's whole-table counts"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from lapis_silo_b200 import host_api
        from oracle import oracle as O
        from test_silo_loader import random_archive
        table, data, _, specs = random_archive(301, 0)  # seeded: the same bytes on every rank
        archive = host_api.Archive(data, specs)
        chunk_sizes = archive.chunk_sizes(0)
        bounds = host_api.partition_chunks([1] * len(chunk_sizes), world_size)
        first, n_chunks = bounds[rank], bounds[rank + 1] - bounds[rank]
        desc = archive.shard_desc(0, first, n_chunks).contents
        local = np.zeros((desc.n_symbols, desc.genome_length), dtype=np.int64)
        for i in range(desc.n_containers):
            container = desc.containers[i]
            assert first <= container.v_index < first + n_chunks
            local[container.symbol, container.position] += container.cardinality
        shard_table = host_api.HostTable.from_archive(None, data, specs, first_chunk=first, n_chunks=n_chunks)
        rows = torch.tensor([shard_table.num_rows], dtype=torch.int64)
        assert shard_table.num_rows == sum(chunk_sizes[first:first + n_chunks])
        shard_table.close()
        reduced = torch.from_numpy(local)
        dist.all_reduce(reduced)
        dist.all_reduce(rows)
        assert int(rows.item()) == table.num_rows
        whole = table.mutation_counts("c").astype(np.int64)
        local_reference = [O.NUC_SYMBOLS.index(c) for c in table.local_reference("c")]
        for symbol in range(5):  # the valid mutation symbols -ACGT; the local reference's count is derived, not stored
            positions = np.array([p for p in range(desc.genome_length) if local_reference[p] != symbol])
            np.testing.assert_array_equal(reduced.numpy()[symbol, positions], whole[symbol, positions])
        archive.close()
        if rank == 0:
            with open(result_path, "w") as out:
                out.write(f"ok {int(rows.item())}")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_load_their_chunk_ranges_from_one_archive(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    result_path = str(tmp_path / "result.txt")
    mp.spawn(archive_worker, args=(2, free_port(), result_path), nprocs=2, join=True)
    with open(result_path) as result:
        assert result.read().startswith("ok ")


def aggregation_worker(rank: int, world_size: int, port: int, result_path: str) -> None:
    """BitmapAggregationNode over a row-partitioned table (BASELINE.json configs[4]): every rank
    aggregates the rows of its chunk range (the oracle plays the device: the shard is the whole table under the
    shard's row-range filter), the (key, count) lists are all-gathered, rank 0 sums them per key and materialises the
    rows with the product's host layer (BitmapAggregationNode::mergeShards / ::materialise on a host-only table)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from lapis_silo_b200 import host_api
        from oracle import oracle as O
        total_rows = 3 * 65536 + 777
        synthetic = host_api.Synthetic(co_occurrence_sequences=20_000)
        sizes = host_api.dense_chunk_sizes(total_rows)
        whole = O.Table()
        whole.set_layout(*sizes)
        whole.import_column("main", O.NUCLEOTIDE, synthetic.reference, synthetic.build_column(total_rows, 0, len(sizes), 2))
        dimensions = [("position", "main", p - 1) for p in (5, 10, 20, 30, 40, 50)]
        bounds = host_api.partition_chunks([1] * len(sizes), world_size)
        first, n_chunks = bounds[rank], bounds[rank + 1] - bounds[rank]
        first_row = first << 16
        end_row = ((first + n_chunks - 1) << 16) + sizes[first + n_chunks - 1]
        symbol = {c: i for i, c in enumerate(O.NUC_SYMBOLS)}
        results = {}
        for name, expression in (("all", None), ("filtered", "(sym-eq main 12 A)")):
            shard_filter = f"(ranges {first_row} {end_row})" if expression is None else f"(and (ranges {first_row} {end_row}) {expression})"
            rows = whole.bitmap_aggregation(dimensions, shard_filter)
            pairs = np.array(sorted((sum(symbol[v] << (5 * (len(dimensions) - 1 - d)) for d, v in enumerate(values)), count)
                                    for *values, count in rows), dtype=np.uint64).reshape(-1, 2)
            shard = (pairs, int(sum(count for *_, count in rows)))
            gathered = [None] * world_size
            dist.all_gather_object(gathered, shard)
            if rank == 0:
                table = host_api.HostTable(None, [10])
                table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_column(10, 0, 1, 1))
                merged = table.bitmap_aggregation_merge(dimensions, gathered)
                table.close()
                assert merged == whole.bitmap_aggregation(dimensions, expression), name
                results[name] = len(merged)
        if rank == 0:
            with open(result_path, "w") as out:
                out.write(f"ok {results}")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_merge_their_shard_combinations(tmp_path):
    pytest.importorskip("torch")
    import torch.multiprocessing as mp
    result_path = str(tmp_path / "result.txt")
    mp.spawn(aggregation_worker, args=(2, free_port(), result_path), nprocs=2, join=True)
    with open(result_path) as result:
        assert result.read().startswith("ok ")
