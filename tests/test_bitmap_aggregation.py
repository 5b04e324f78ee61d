"""BitmapAggregationNode (mutation co-occurrence / groupBy over sequence positions and indexed columns).

CPU part: the oracle's restatement against the literal expectations of the reference's own scenarios
(operators/bitmap_aggregation_node.test.cpp:67-283). GPU part: the product's device path
(silo_gpu_query_combinations through the host layer's BitmapAggregationNode) against those same
literals and against the oracle on random tables."""
import numpy as np
import pytest


def make_table(rows):
    """QueryTestData of the reference scenarios: segment1 reference "ATGCN", gene1 reference "M*"."""
    from oracle import oracle as O
    t = O.Table()
    t.add_column("segment1", O.NUCLEOTIDE, "ATGCN")
    t.add_column("gene1", O.AMINO_ACID, "M*")
    for nuc, aa in rows:
        t.append_row([nuc, aa])
    t.finalize()
    return t


def base_table():  # TEST_DATA, bitmap_aggregation_node.test.cpp:33-66 (region: Europe, Europe, Asia, Europe)
    t = make_table([("ATGCN", "M*"), ("ATGCN", "C*"), ("NNNNN", "M*"), ("CATTT", "X*")])
    t.register_bitmap("region=Europe", [0, 1, 3])
    t.register_bitmap("region=Asia", [2])
    return t


def null_table():  # NULL_TEST_DATA, :190-199
    return make_table([("ATGCN", "M*"), ("ATGCN", "M*"), (None, "M*"), ("CATTT", None)])


def ambiguity_table():  # AMBIGUITY_TEST_DATA, :260-270
    return make_table([("ATGCN", "M*"), ("RTGCN", "M*"), ("RTGCN", "M*"), ("YTGCN", "M*")])


REGION = ("bitmaps", [("Europe", "region=Europe"), ("Asia", "region=Asia")], None)

# (table factory, dimensions, filter, expected) -- positions are 0-based here, 1-based in the queries
SCENARIOS = [
    (base_table, [("position", "segment1", 0), ("position", "segment1", 1)], None,
     [("A", "T", 2), ("C", "A", 1), ("N", "N", 1)]),                              # CO_OCCURRENCE_VIA_MAP_TWO_POSITIONS
    (base_table, [("position", "segment1", 0), ("position", "segment1", 1)], "(has-mut segment1 1)",
     [("C", "A", 1)]),                                                            # ..._WITH_FILTER
    (base_table, [("position", "gene1", 1)], None, [("*", 4)]),                  # ..._AMINO_ACID
    (base_table, [REGION], None, [("Asia", 1), ("Europe", 3)]),                  # INDEXED_COLUMN_SINGLE
    (base_table, [("position", "segment1", 0), REGION], None,
     [("A", "Europe", 2), ("C", "Europe", 1), ("N", "Asia", 1)]),                # MIXED_SEQUENCE_AND_INDEXED_COLUMN
    (null_table, [("position", "segment1", 0), ("position", "segment1", 1)], None,
     [("A", "T", 2), ("C", "A", 1), (None, None, 1)]),                           # CO_OCCURRENCE_NULL_TWO_NUCLEOTIDE_POSITIONS
    (null_table, [("position", "gene1", 0)], None, [("M", 3), (None, 1)]),       # CO_OCCURRENCE_NULL_AMINO_ACID
    (null_table, [("position", "segment1", 0), ("position", "gene1", 0)], None,
     [("A", "M", 2), ("C", None, 1), (None, "M", 1)]),                           # CO_OCCURRENCE_NULL_MIXED_POSITIONS
    (ambiguity_table, [("position", "segment1", 0)], None, [("A", 1), ("R", 2), ("Y", 1)]),  # CO_OCCURRENCE_AMBIGUOUS_CODES
]
OUT_OF_RANGE = "SymbolInSet<Nucleotide> position is out of bounds 6 > 5"  # ..._POSITION_OUT_OF_RANGE, :128-132


@pytest.mark.parametrize("index", range(len(SCENARIOS)))
def test_oracle_matches_the_reference_scenarios(index):
    factory, dimensions, expression, expected = SCENARIOS[index]
    assert factory().bitmap_aggregation(dimensions, expression) == expected


def test_oracle_position_out_of_range():
    from oracle import oracle as O
    with pytest.raises(O.OracleError, match=OUT_OF_RANGE):
        base_table().bitmap_aggregation([("position", "segment1", 5)])


def test_oracle_equals_a_row_by_row_group_by():
    """Independent check of the restated partition: group the generated strings row by row."""
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    reference = "".join(rng.choice(list("ACGT"), 30))
    rows = []
    for _ in range(500):
        seq = list(reference)
        for p in rng.integers(0, 30, 4):
            seq[int(p)] = "ACGTN-R"[int(rng.integers(0, 7))]
        rows.append(None if rng.random() < 0.05 else "".join(seq))
    t = O.Table()
    t.add_column("c", O.NUCLEOTIDE, reference)
    for seq in rows:
        t.append_row([seq])
    t.finalize()
    positions = [3, 11, 12, 29]
    want = {}
    for seq in rows:
        key = tuple(None if seq is None else seq[p] for p in positions)
        want[key] = want.get(key, 0) + 1
    order = {c: i for i, c in enumerate(O.NUC_SYMBOLS)}
    expected = sorted(want.items(), key=lambda item: tuple(order[v] if v is not None else 99 for v in item[0]))
    got = t.bitmap_aggregation([("position", "c", p) for p in positions])
    assert got == [key + (count,) for key, count in expected]


# ---- device ------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def ctx():
    from lapis_silo_b200 import abi
    context = abi.Context(0)
    yield context
    context.close()


@pytest.mark.gpu
@pytest.mark.parametrize("index", range(len(SCENARIOS)))
def test_device_matches_the_reference_scenarios(ctx, index):
    from test_gpu_parity import mirror
    factory, dimensions, expression, expected = SCENARIOS[index]
    oracle_table = factory()
    bitmaps = ["region=Europe", "region=Asia"] if factory is base_table else []
    for resident in (True, False):
        device_table = mirror(ctx, oracle_table, bitmaps, resident=resident)
        assert device_table.bitmap_aggregation(dimensions, expression) == expected
        device_table.close()


@pytest.mark.gpu
def test_device_position_out_of_range(ctx):
    from lapis_silo_b200 import host_api
    from test_gpu_parity import mirror
    device_table = mirror(ctx, base_table())
    with pytest.raises(host_api.HostError, match=OUT_OF_RANGE):
        device_table.bitmap_aggregation([("position", "segment1", 5)])


@pytest.mark.gpu
@pytest.mark.parametrize("seed,alphabet_id", [(301, 0), (302, 1)])
def test_device_equals_oracle_on_random_tables(ctx, seed, alphabet_id):
    """Several chunks (a partial one, single-row ones), nulls, N runs, flipped local references, every
    container kind; 1 to 7 dimensions incl. index bitmaps; filters from empty to full."""
    from test_gpu_parity import build_random, mirror
    t = build_random(seed, 1500, 50, (299, 300, 700), alphabet_id)
    rng = np.random.default_rng(seed)
    n_rows = t.num_rows
    row_ids = [(chunk << 16) | r for chunk, size in enumerate(t.chunk_sizes) for r in range(size)]
    assert len(row_ids) == n_rows
    assignment = rng.integers(0, 4, n_rows)  # group 3 = null
    names = []
    for g, name in enumerate(("val=b", "val=a", "val=c", "val-null")):
        t.register_bitmap(name, [row_ids[i] for i in range(n_rows) if assignment[i] == g])
        names.append(name)
    t.register_bitmap("lineage", [row_ids[i] for i in rng.choice(n_rows, n_rows // 3, replace=False)])
    device_table = mirror(ctx, t, names + ["lineage"], resident=(alphabet_id == 0))
    indexed = ("bitmaps", [("b", "val=b"), ("a", "val=a"), ("c", "val=c")], "val-null")
    partial = ("bitmaps", [("a", "val=a"), ("b", "val=b")], None)  # rows of the other groups are in no combination
    filters = [None, "(true)", "(false)", "(bitmap lineage)", "(not (bitmap lineage))", "(has-mut c 7)", "(ranges 3 90 196608 197000)"]
    dimension_sets = [
        [("position", "c", 0)],
        [("position", "c", 6), ("position", "c", 7)],
        [indexed],
        [("position", "c", 11), indexed, ("position", "c", 30)],
        [partial, ("position", "c", 49)],
        [("position", "c", int(p)) for p in rng.choice(50, 7, replace=False)],
    ]
    for dimensions in dimension_sets:
        for expression in filters:
            want = t.bitmap_aggregation(dimensions, expression)
            got = device_table.bitmap_aggregation(dimensions, expression)
            assert got == want, (dimensions, expression)
    device_table.close()


# ---- row-partitioned tables: executeShard on every rank, mergeShards + materialise on one (SURVEY.md 8(e), configs[4]) ----

def test_merge_of_shard_combinations_on_the_host():
    """BitmapAggregationNode::mergeShards / ::materialise without a device: keys are dimension 0 in the most
    significant bits, 5 bits per sequence position; equal keys add up, the output is ordered by key."""
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    synthetic = host_api.Synthetic(co_occurrence_sequences=50)
    table = host_api.HostTable(None, [50])
    table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_column(50, 0, 1, 1))
    symbol = {c: i for i, c in enumerate(O.NUC_SYMBOLS)}
    key = lambda a, b: (symbol[a] << 5) | symbol[b]
    dimensions = [("position", "main", 4), ("position", "main", 9)]
    shard_a = (np.array([[key("A", "C"), 3], [key("A", "T"), 1], [key("G", "C"), 7]], dtype=np.uint64), 11)
    shard_b = (np.array([[key("-", "N"), 2], [key("A", "T"), 5]], dtype=np.uint64), 7)
    shard_c = (np.zeros((0, 2), dtype=np.uint64), 0)
    assert table.bitmap_aggregation_merge(dimensions, [shard_a, shard_b, shard_c]) == [
        ("-", "N", 2), ("A", "C", 3), ("A", "T", 6), ("G", "C", 7)]
    assert table.bitmap_aggregation_merge([], [shard_a, shard_b, shard_c]) == [(18,)]  # no dimension: the cardinalities add up
    # an indexed dimension in front: 8 key bits, its values in sorted order whatever order they were given in, the null group last
    indexed = ("bitmaps", [("b", "val=b"), ("a", "val=a"), ("c", "val=c")], "val-null")
    mixed = [indexed, ("position", "main", 4)]
    key2 = lambda group, a: (group << 5) | symbol[a]
    shard_d = (np.array([[key2(0, "A"), 4], [key2(2, "T"), 1], [key2(3, "A"), 2]], dtype=np.uint64), 7)
    shard_e = (np.array([[key2(0, "A"), 1], [key2(1, "G"), 9]], dtype=np.uint64), 10)
    assert table.bitmap_aggregation_merge(mixed, [shard_d, shard_e]) == [("a", "A", 5), ("b", "G", 9), ("c", "T", 1), (None, "A", 2)]
    codes, counts = table.bitmap_aggregation_merge_columns(mixed, [shard_d, shard_e])
    assert codes.tolist() == [[0, ord("A")], [1, ord("G")], [2, ord("T")], [255, ord("A")]] and counts.tolist() == [5, 9, 1, 2]
    # the same result as arrays (what bench.py reads: no text to format and parse)
    codes, counts = table.bitmap_aggregation_merge_columns(dimensions, [shard_a, shard_b, shard_c])
    assert host_api.combination_rows_from_columns(codes, counts) == [("-", "N", 2), ("A", "C", 3), ("A", "T", 6), ("G", "C", 7)]
    table.close()


def test_co_occurrence_generator():
    """performance/sequence_generator.h:487-526: 100-nt A/C/G/T reference, Binomial(100, 0.1) substitutions per row
    (a substitution may draw the base that is already there: ~7.2 differing positions per row on average)."""
    from lapis_silo_b200 import host_api
    synthetic = host_api.Synthetic(co_occurrence_sequences=4000)
    again = host_api.Synthetic(co_occurrence_sequences=3000)
    reference = synthetic.reference
    assert len(reference) == 100 and set(reference) <= set("ACGT") and again.reference == reference
    sequences = [synthetic.sequence(i) for i in range(3000)]
    assert sequences == [again.sequence(i) for i in range(3000)]  # one stream in id order: a prefix is a prefix
    assert all(len(s) == 100 and set(s) <= set("ACGT") for s in sequences)
    differing = np.array([sum(a != b for a, b in zip(s, reference)) for s in sequences])
    assert 6.9 < differing.mean() < 7.5 and differing.max() <= 25


@pytest.mark.gpu
def test_sharded_aggregation_merges_to_the_whole(ctx):
    """Three interleaved shards of the co_occurrence_benchmark table, every shard aggregated on the device, the (key,
    count) lists merged: the rows of the oracle's aggregation over the whole table, with and without filters."""
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    total_rows = 3 * 65536 + 4321
    synthetic = host_api.Synthetic(co_occurrence_sequences=50_000)
    sizes = host_api.dense_chunk_sizes(total_rows)
    oracle_table = O.Table()
    oracle_table.set_layout(*sizes)
    oracle_table.import_column("main", O.NUCLEOTIDE, synthetic.reference, synthetic.build_column(total_rows, 0, len(sizes), 4))
    shards = []
    for rank in range(3):
        first, n_chunks, stride = host_api.interleaved_shard(len(sizes), 3, rank)
        table = host_api.HostTable(ctx, host_api.shard_chunk_sizes(total_rows, first, n_chunks, stride))
        table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_column(total_rows, first, n_chunks, 4, stride))
        shards.append(table)
    six = [("position", "main", p - 1) for p in (5, 10, 20, 30, 40, 50)]  # co_occurrence_benchmark.cpp:41 (1-based there)
    for dimensions in (six, six[:2], [("position", "main", 99)], []):
        for expression in (None, "(sym-eq main 12 A)", "(and (has-mut main 7) (not (sym-eq main 60 T)))", "(false)"):
            parts = [table.bitmap_aggregation_shard(dimensions, expression) for table in shards]
            assert all((np.diff(pairs[:, 0].astype(np.int64)) > 0).all() for pairs, _ in parts)
            merged = shards[0].bitmap_aggregation_merge(dimensions, parts)
            want = oracle_table.bitmap_aggregation(dimensions, expression)
            assert merged == want, (dimensions, expression)
            if dimensions:
                assert host_api.combination_rows_from_columns(*shards[0].bitmap_aggregation_merge_columns(dimensions, parts)) == want
                whole_of_one = shards[1].bitmap_aggregation_columns(dimensions, expression)
                assert host_api.combination_rows_from_columns(*whole_of_one) == shards[1].bitmap_aggregation(dimensions, expression)
            if dimensions == six and expression is None:
                assert sum(count for *_, count in merged) == total_rows and len(merged) > 100
    for table in shards:
        table.close()
