"""GPU parity proper: filter expressions and the Mutations action through the product's host layer
(libsilo_b200_host.so -> C ABI -> sm_100a kernels) against the CPU oracle on the same inputs, and
against the literal expectations of the reference's own unit tests. Bit-exact: row-id sets, u32
counts, and the thresholded output rows including the double proportions."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from lapis_silo_b200 import abi
    context = abi.Context(0)
    yield context
    context.close()


def mirror(ctx, oracle_table, bitmaps=(), resident=True):
    """The same table on the device: every column uploaded from the oracle's S1 export. Index bitmaps
    are made device resident once (resident) or travel with every program."""
    from lapis_silo_b200 import host_api
    table = host_api.HostTable(ctx, oracle_table.chunk_sizes)
    for name, alphabet, reference in oracle_table.columns:
        export = oracle_table.export_column(name)
        table.add_column(name, alphabet, reference, export.desc)
        export.close()
    for name in bitmaps:
        table.register_bitmap(name, oracle_table.bitmap_bytes(name), resident)
    return table


def layout_pair(ctx, *sizes):
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    oracle_table = O.Table()
    oracle_table.set_layout(*sizes)
    return oracle_table, host_api.HostTable(ctx, list(sizes))


def ids(table, expression):
    return [int(v) for v in table.filter(expression).ids()]


def lists(sets):
    return " ".join("(ids " + " ".join(map(str, s)) + ")" for s in sets)


def both(pair, expression):
    oracle_table, device_table = pair
    want = ids(oracle_table, expression)
    flt = device_table.filter(expression)
    got = [int(v) for v in flt.ids()]
    assert got == want, expression
    assert flt.cardinality == len(want), expression
    return got


# ---- the reference's operator-level vectors (in-layout inputs) -------------------------------

def test_threshold_vectors(ctx):  # filter/operators/threshold.test.cpp:40-247
    def threshold(pair, pos, neg, k, exact):
        return both(pair, f"(op-threshold {k} {int(exact)} ({lists(pos)}) ({lists(neg)}))")

    pair = layout_pair(ctx, 4)
    assert threshold(pair, [], [[1, 2, 3], [1, 3]], 1, True) == [2]
    assert threshold(pair, [], [[1, 2, 3], [1, 3]], 1, False) == [0, 2]
    pos = [[1, 2], [1, 3], [1, 2, 3]]
    assert threshold(pair, pos, [], 1, True) == []
    assert threshold(pair, pos, [], 2, True) == [2, 3]
    assert threshold(pair, pos, [], 1, False) == [1, 2, 3]
    assert threshold(pair, pos, [], 2, False) == [1, 2, 3]
    pos, neg = [[1, 2, 3], [1, 3], [1, 2, 3]], [[], [3]]
    assert [threshold(pair, pos, neg, k, True) for k in (1, 2, 3, 4)] == [[], [0], [], [2, 3]]
    assert [threshold(pair, pos, neg, k, False) for k in (1, 2, 3, 4)] == [
        [0, 1, 2, 3], [0, 1, 2, 3], [1, 2, 3], [1, 2, 3]]
    pair = layout_pair(ctx, 5)
    pos, neg = [[1, 2, 3]], [[], [3], [4], [2, 4]]
    assert [threshold(pair, pos, neg, k, True) for k in (1, 2, 3, 4)] == [[], [4], [], [0, 2, 3]]
    assert [threshold(pair, pos, neg, k, False) for k in (1, 2, 3, 4)] == [
        [0, 1, 2, 3, 4], [0, 1, 2, 3, 4], [0, 1, 2, 3], [0, 1, 2, 3]]


def test_threshold_with_ids_outside_the_layout(ctx):
    """threshold.test.cpp:249-311 feeds id 4 into a 4-row layout (row_layout.h:49-52 documents that inputs are expected
    to be a subset of the universe): complements stay inside the layout (row_layout.cpp:18-23), so the id takes part in
    the count because a child holds it. The device reproduces the reference's expectations (and the oracle's)."""
    def threshold(pair, pos, neg, k, exact):
        return both(pair, f"(op-threshold {k} {int(exact)} ({lists(pos)}) ({lists(neg)}))")

    pair = layout_pair(ctx, 4)
    neg = [[3], [4], [2, 4]]
    assert [threshold(pair, [[]], neg, k, True) for k in (1, 2, 3)] == [[4], [2, 3], [0, 1]]
    assert [threshold(pair, [[]], neg, k, False) for k in (1, 2, 3)] == [[0, 1, 2, 3, 4], [0, 1, 2, 3], [0, 1]]
    # the boolean operators carry such ids like the reference's bitmaps do
    assert both(pair, "(or (ids 1) (ids 4))") == [1, 4]
    assert both(pair, "(and (ids 1 4) (ids 4 2))") == [4]
    assert both(pair, "(not (ids 1 4))") == [0, 2, 3, 4]
    # ... and what needs per-row data for the filter's rows refuses them
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    with_rows = O.Table()
    with_rows.add_column("c", O.NUCLEOTIDE, "ACGT")
    for _ in range(4):
        with_rows.append_row(["ACGT"])
    with_rows.finalize()
    mirrored = mirror(ctx, with_rows)
    with pytest.raises(host_api.HostError, match=r"DeviceError\[-6\]"):
        mirrored.mutation_counts("c", mirrored.filter("(ids 1 4)"))
    with pytest.raises(host_api.HostError, match=r"DeviceError\[-6\]"):
        mirrored.mutations(["c"], "(ids 1 4)", 0.0)


def test_constructor_errors_match(ctx):  # threshold.cpp:30-41, intersection.cpp:26-40
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    oracle_table, device_table = layout_pair(ctx, 5)
    for expression in ("(op-threshold 2 0 ((ids 1) (ids 2)) ())", "(op-threshold 0 0 ((ids 1) (ids 2)) ())",
                       "(op-and () ())", "(op-and () ((ids 1) (ids 2)))", "(op-and ((ids 1)) ())"):
        with pytest.raises(O.OracleError) as want:
            oracle_table.filter(expression)
        with pytest.raises(host_api.HostError) as got:
            device_table.filter(expression)
        assert str(got.value) == str(want.value)


def test_intersection_complement_union_vectors(ctx):
    pair = layout_pair(ctx, 5)  # intersection.test.cpp:20-136
    assert both(pair, f"(op-and ({lists([[1, 2, 3], [1, 3], [1, 2, 3]])}) ())") == [1, 3]
    assert both(pair, f"(op-and ({lists([[1, 2, 3], [1, 3], [1, 2, 3]])}) ({lists([[], [3]])}))") == [1]
    assert both(pair, f"(op-and ({lists([[1, 2, 3]])}) ({lists([[], [3], [4], [2, 4]])}))") == [1]
    assert both(pair, f"(op-and ({lists([[]])}) ({lists([[3], [4], [2, 4]])}))") == []
    assert both(pair, "(op-not (ids 1 2 3))") == [0, 4]  # complement.test.cpp:14-72
    assert both(pair, "(op-not (ids 1))") == [0, 2, 3, 4]
    assert both(layout_pair(ctx, 3), "(op-not (ids))") == [0, 1, 2]
    assert both(layout_pair(ctx, 4), "(op-not (ids 0 1 2 3))") == []
    assert both(layout_pair(ctx), "(op-not (ids))") == []
    pair = layout_pair(ctx, 3, 2)  # row_layout.cpp:18-23: the gap between chunks stays empty
    assert both(pair, "(op-not (ids 1 65536))") == [0, 2, 65537]
    assert both(pair, "(true)") == [0, 1, 2, 65536, 65537]
    assert both(pair, "(not (true))") == []
    pair = layout_pair(ctx, 65536, 65536, 65536, 65536, 2)  # copy_on_write_bitmap.test.cpp:48-90
    m = [1, 5, 100, 65536 + 3, 3 * 65536 + 7]
    both(pair, f"(op-and ({lists([m, [5, 100, 65536 + 3, 999]])}) ())")
    both(pair, f"(op-and ({lists([m])}) ({lists([[5, 65536 + 3]])}))")
    both(pair, f"(op-or {lists([[1, 5], m, [5, 3 * 65536 + 7, 4 * 65536 + 1]])})")


def test_fast_union_staggered(ctx):  # copy_on_write_bitmap.test.cpp:130-185
    pair = layout_pair(ctx, *([65536] * 12))
    inputs = []
    for i in range(8):
        s = set()
        for key in range(i, i + 5):
            s.update(range((key << 16) + i, (key << 16) + i + 5000 + 100 * key))
            s.add((key << 16) + 60000)
        inputs.append(sorted(s))
    expected = sorted(set().union(*map(set, inputs)))
    assert both(pair, f"(op-or {lists(inputs)})") == expected


# ---- expression-level vectors ----------------------------------------------------------------

@pytest.fixture(scope="module")
def atgcn(ctx):
    from oracle import oracle as O
    t = O.Table()
    t.add_column("segment1", O.NUCLEOTIDE, "ATGCN")
    for seq in ("ATGCN", "ATGCN", "NNNNN", "CATTT", None):
        t.append_row([seq])
    t.finalize()
    return t, mirror(ctx, t)


@pytest.mark.parametrize("symbol,position,count", [
    ("A", 1, 2), ("A", 2, 1), ("A", 3, 0), ("A", 4, 0), ("A", 5, 0),
    ("C", 1, 1), ("C", 2, 0), ("C", 3, 0), ("C", 4, 2), ("C", 5, 0),
    ("G", 1, 0), ("G", 2, 0), ("G", 3, 2), ("G", 4, 0), ("G", 5, 0),
    ("T", 1, 0), ("T", 2, 2), ("T", 3, 1), ("T", 4, 1), ("T", 5, 1),
    ("N", 1, 1), ("N", 5, 3), (".", 1, 2),
])
def test_symbol_equals_counts(atgcn, symbol, position, count):  # symbol_equals.test.cpp:42-257
    assert len(both(atgcn, f"(sym-eq segment1 {position} {symbol})")) == count


def test_symbol_equals_errors_match(atgcn):
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    oracle_table, device_table = atgcn
    for expression in ("(sym-eq segment1 1000 A)", "(sym-eq segment1 0 A)", "(sym-eq nope 1 A)",
                       "(has-mut segment1 6)", "(sym-eq segment1 1 Z)"):
        with pytest.raises(O.OracleError) as want:
            oracle_table.filter(expression)
        with pytest.raises(host_api.HostError) as got:
            device_table.filter(expression)
        assert str(got.value) == str(want.value)


def test_has_mutation(ctx):  # has_mutation.test.cpp:28-60
    from oracle import oracle as O
    t = O.Table()
    t.add_column("segment1", O.NUCLEOTIDE, "ATGCN")
    t.add_column("gene1", O.AMINO_ACID, "M*")
    for nuc, aa in (("ATGCN", "M*"), ("ATGCN", "C*"), ("NNNNN", "M*"), ("CATTT", "X*")):
        t.append_row([nuc, aa])
    t.finalize()
    pair = (t, mirror(ctx, t))
    assert len(both(pair, "(has-mut segment1 1)")) == 1
    assert len(both(pair, "(has-mut gene1 1)")) == 1
    both(pair, "(maybe (has-mut gene1 1))")
    both(pair, "(not (maybe (has-mut segment1 1)))")


def test_mutation_profile_vectors(ctx):  # mutation_profile.test.cpp:24-264
    from oracle import oracle as O
    t = O.Table()
    t.add_column("segment1", O.NUCLEOTIDE, "ATGCN")
    t.add_column("gene1", O.AMINO_ACID, "M*")
    for nuc, aa in (("ATGCN", "M*"), ("CTGCN", "C*"), ("CTCCN", "M*"), ("CTCTN", "M*"), ("NNNNN", "M*"), ("RTGCN", "M*")):
        t.append_row([nuc, aa])
    t.finalize()
    pair = (t, mirror(ctx, t))
    ref, mut1, mut2, mut3, alln, amb = range(6)
    assert both(pair, "(profile segment1 0 muts)") == [ref, alln, amb]
    assert both(pair, "(profile segment1 1 muts)") == [ref, mut1, alln, amb]
    assert both(pair, "(profile segment1 2 muts)") == [ref, mut1, mut2, alln, amb]
    assert both(pair, "(profile segment1 0 muts 1 C)") == [mut1, alln]
    assert both(pair, "(profile segment1 0 seq ATGCN)") == [ref, alln, amb]
    assert both(pair, "(profile segment1 0 seq CTGCN)") == [mut1, alln]  # = sequenceId:='seq_1mut'
    assert both(pair, "(profile gene1 0 muts)") == [ref, mut2, mut3, alln, amb]
    assert both(pair, "(profile gene1 1 muts)") == [ref, mut1, mut2, mut3, alln, amb]
    assert both(pair, "(profile gene1 0 muts 1 C)") == [mut1]
    assert both(pair, "(profile gene1 0 seq M*)") == [ref, mut2, mut3, alln, amb]
    assert both(pair, "(profile gene1 0 seq C*)") == [mut1]


# ---- randomised parity: every expression kind, ragged chunks, nulls, offsets, N runs -----------

def build_random(seed, n_rows, length, flushes, alphabet_id=0):
    from test_gpu_kernels import random_table
    return random_table(seed, n_rows, length, flushes=flushes, alphabet=alphabet_id)


@pytest.mark.parametrize("seed,alphabet_id", [(101, 0), (102, 1)])
def test_random_expressions(ctx, seed, alphabet_id):
    t = build_random(seed, 1200, 45, (299, 300, 700), alphabet_id)
    rng = np.random.default_rng(seed)
    picked = sorted(set(int(v) for v in rng.integers(0, 300, 120)) | {(1 << 16)} | {(3 << 16) + int(v) for v in rng.integers(0, 499, 200)})
    t.register_bitmap("lineage", picked)
    pair = (t, mirror(ctx, t, ["lineage"], resident=(alphabet_id == 0)))
    chars = "-ACGTRYSWKMBDHVN" if alphabet_id == 0 else "-ACDEFGHIKLMNOPQRSTUVWYBJZ*X"
    leaves = []
    for _ in range(40):
        position = int(rng.integers(1, 46))
        symbol = chars[int(rng.integers(0, len(chars)))]
        leaves.append(f"(sym-eq c {position} {symbol})")
        leaves.append(f"(has-mut c {position})")
    leaves += ["(bitmap lineage)", "(ranges 5 250 65536 65537 196608 196900)", "(true)", "(false)", "(sym-eq c 3 .)"]
    for leaf in leaves:
        both(pair, leaf)
        both(pair, f"(maybe {leaf})")
        both(pair, f"(not {leaf})")
        both(pair, f"(not (maybe {leaf}))")

    def pick(n):
        return [leaves[int(i)] for i in rng.integers(0, len(leaves), n)]

    for _ in range(60):
        a, b, c, d = pick(4)
        both(pair, f"(and {a} {b})")
        both(pair, f"(or {a} {b} {c})")
        both(pair, f"(and {a} (not {b}) (or {c} (not {d})))")
        both(pair, f"(or (and {a} {b}) (not (and {c} {d})))")
        both(pair, f"(exact (or {a} (maybe {b})))")
        both(pair, f"(and (not {a}) (not {b}))")
        both(pair, f"(not (or {a} (and {b} (bitmap lineage))))")
    for _ in range(40):
        n = int(rng.integers(2, 9))
        children = " ".join(pick(n))
        k = int(rng.integers(0, n + 2))
        for exact in (0, 1):
            both(pair, f"(n-of {k} {exact} {children})")
            both(pair, f"(maybe (n-of {k} {exact} {children}))")
            both(pair, f"(and (bitmap lineage) (not (n-of {k} {exact} {children})))")
    # counter programs inside counter programs (an NOf / a wide Or / a MutationProfile as a child of an NOf): the
    # reference's Threshold takes any child; the device evaluates such children to tiles first
    wide = " ".join(f"(sym-in c {p} {''.join(rng.choice(list(chars[:-1]), 3, replace=False))})" for p in range(1, 31))
    for k, exact in ((1, 0), (2, 0), (2, 1), (3, 0)):
        a, b, c, d = pick(4)
        both(pair, f"(n-of {k} {exact} {a} (n-of 2 0 {b} {c} {d}) (not (n-of 1 1 {a} {c} {d})) {b})")
        both(pair, f"(n-of {k} {exact} (or {wide}) {a} (n-of 7 0 {wide}) (profile c 3 muts))")
        both(pair, f"(n-of {k} {exact} (and {a} (n-of 2 0 {b} {c} {d})) (not (profile c 1 muts)) {d})")
    # many children on one column: exercises the one-pass profile lowering
    for k in (1, 2, 5, 12, 30):
        children = " ".join(f"(sym-in c {p} {''.join(rng.choice(list(chars[:-1]), 3, replace=False))})" for p in range(1, 46))
        both(pair, f"(n-of {k} 0 {children})")
        both(pair, f"(n-of {k} 1 {children})")
    reference = t.columns[0][2]
    for distance in (0, 1, 3, 10, 44):
        both(pair, f"(profile c {distance} muts)")
        query = list(reference)
        for p in rng.integers(0, 45, 6):
            query[int(p)] = chars[int(rng.integers(0, len(chars)))]
        both(pair, f"(profile c {distance} seq {''.join(query)})")
        both(pair, f"(and (bitmap lineage) (profile c {distance} seq {''.join(query)}))")


@pytest.mark.parametrize("seed,alphabet_id", [(201, 0), (202, 1)])
def test_mutations_action_rows(ctx, seed, alphabet_id):
    from lapis_silo_b200 import host_api
    t = build_random(seed, 900, 60, (99, 130, 131), alphabet_id)
    rng = np.random.default_rng(seed)
    t.register_bitmap("lineage", sorted({int(v) for v in rng.integers(0, 100, 60)} | {(3 << 16) + int(v) for v in rng.integers(0, 700, 400)}))
    oracle_table, device_table = t, mirror(ctx, t, ["lineage"], resident=(alphabet_id == 1))
    filters = [None, "(true)", "(false)", "(bitmap lineage)", "(not (bitmap lineage))", "(has-mut c 7)",
               "(and (bitmap lineage) (not (sym-eq c 12 N)))", "(ranges 3 90 196608 197000)", "(profile c 4 muts)"]
    for expression in filters:
        for min_proportion in (0.0, 1e-9, 0.05, 1.0 / 3.0, 0.3, 0.5, 0.999, 1.0):
            want = oracle_table.mutations("c", expression, min_proportion)
            got = device_table.mutations(["c"], expression, min_proportion)
            assert got == want, (expression, min_proportion)
            # the same result as one record batch from one call (silo_host_mutations_packed)
            assert host_api.rows_from_columns(device_table.mutations_columns(["c"], expression, min_proportion)) == want
        device_table._packed = np.empty(8, dtype=np.uint8)  # too small: the result is fetched with a second call
        assert host_api.rows_from_columns(device_table.mutations_columns(["c"], expression, 0.0)) == oracle_table.mutations("c", expression, 0.0)
        flt_o = oracle_table.filter(expression) if expression else None
        flt_d = device_table.filter(expression) if expression else None
        np.testing.assert_array_equal(device_table.mutation_counts("c", flt_d), oracle_table.mutation_counts("c", flt_o))


def test_mutations_query_graph_replay(ctx):
    """The fused Mutations query replays a CUDA graph per query SHAPE from the second occurrence in a
    row on (csrc/mutations.cu silo_gpu_query_mutation_hits). Queries of one shape with different
    content, shapes that alternate, more shapes than the cache holds and a hits buffer that regrows
    must all keep giving the oracle's rows."""
    t = build_random(401, 900, 60, (99, 130, 131), 0)
    rng = np.random.default_rng(401)
    t.register_bitmap("lineage", sorted({int(v) for v in rng.integers(0, 100, 60)} | {(3 << 16) + int(v) for v in rng.integers(0, 700, 400)}))
    oracle_table, device_table = t, mirror(ctx, t, ["lineage"], resident=True)

    def check(expression, min_proportion=0.05):
        assert device_table.mutations(["c"], expression, min_proportion) == oracle_table.mutations("c", expression, min_proportion), expression

    # one shape, the content changes with every call (eager, capture, then replays)
    for position in list(range(1, 40)) * 2:
        check(f"(and (bitmap lineage) (not (sym-eq c {position} N)))")
    # the same query many times
    for _ in range(6):
        check("(has-mut c 7)")
    # shapes alternating, each repeated so that it gets a graph; 12 shapes > the 8 cached ones
    shapes = [" ".join(f"(has-mut c {p})" for p in range(1, n + 2)) for n in range(12)]
    for _ in range(3):
        for n, children in enumerate(shapes):
            for _ in range(3):
                check(f"(or {children} (ranges {n} {80 + n}))")
    # thresholds and proportions are part of the shape (kernel parameters)
    for min_proportion in (0.0, 0.3, 0.0, 0.0, 0.3, 0.3, 1.0, 1.0):
        check("(profile c 4 muts)", min_proportion)
        check("(bitmap lineage)", min_proportion)
    # no program / the full-filter path next to graph replays
    for expression in (None, "(true)", "(false)", "(bitmap lineage)", "(bitmap lineage)", "(bitmap lineage)", None):
        check(expression)


def test_mutations_two_columns_and_union_all_vector(ctx):
    from oracle import oracle as O
    t = O.Table()  # operators/union_all_node.test.cpp:184-193
    t.add_column("main", O.NUCLEOTIDE, "A")
    t.append_row(["T"])
    t.finalize()
    device_table = mirror(ctx, t)
    rows = device_table.mutations(["main"], None, 0.0)
    assert [(r["mutationTo"], r["proportion"]) for r in rows] == [("T", 1.0)]
    t2 = O.Table()
    t2.add_column("segment1", O.NUCLEOTIDE, "ATGCN")
    t2.add_column("gene1", O.AMINO_ACID, "M*")
    for nuc, aa in (("ATGCN", "M*"), ("CTGCN", "C*"), ("CTCCN", None), (None, "M*"), ("NNNNN", "XX"), ("RTGCN", "M-")):
        t2.append_row([nuc, aa])
    t2.finalize()
    d2 = mirror(ctx, t2)
    for expression in (None, "(sym-eq gene1 1 M)", "(not (sym-eq segment1 1 N))"):
        want = t2.mutations("segment1", expression, 0.0) + t2.mutations("gene1", expression, 0.0)
        assert d2.mutations(["segment1", "gene1"], expression, 0.0) == want


def test_synthetic_shards_sum_to_the_whole(ctx):
    """Multi-GPU invariant on one device: per-shard counts are plain addends."""
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    total_rows = 4 * 65536 + 12345
    synthetic = host_api.Synthetic(genome_length=1500, reference_seed=9, generations=5)
    sizes = host_api.dense_chunk_sizes(total_rows)
    date_filter = lambda first, n: host_api.date_ranges_expression(total_rows, 1095, 200, 800, first, n)
    ancestor = next(i for i in range(synthetic.num_sequences) if synthetic.generation(i) == 2)

    def shard(first, n, stride=1):
        # a contiguous shard keeps global chunk ids; an interleaved one is a table of its own
        table = host_api.HostTable(
            ctx, host_api.shard_chunk_sizes(total_rows, first, n, stride), first_chunk=first if stride == 1 else 0)
        table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_column(total_rows, first, n, 4, stride))
        table.register_bitmap("lineage", synthetic.lineage_bitmap(ancestor, total_rows, first, n, stride))
        expression = f"(and {host_api.date_ranges_expression(total_rows, 1095, 200, 800, first, n, stride)} (bitmap lineage))"
        flt = table.filter(expression)
        return table, flt, table.mutation_counts("main", flt), table.mutation_counts("main")

    whole_table, whole_filter, whole_counts, whole_full = shard(0, len(sizes))
    bounds = host_api.partition_chunks([1] * len(sizes), 3)
    parts = [shard(a, b - a) for a, b in zip(bounds, bounds[1:])]
    assert sum(p[1].cardinality for p in parts) == whole_filter.cardinality > 0
    np.testing.assert_array_equal(sum(p[2].astype(np.uint64) for p in parts).astype(np.uint32), whole_counts)
    np.testing.assert_array_equal(sum(p[3].astype(np.uint64) for p in parts).astype(np.uint32), whole_full)
    # interleaved shards (chunk c on rank c % 3), what bench.py uses for N > 1: balanced under a date filter
    interleaved = [shard(*host_api.interleaved_shard(len(sizes), 3, rank)) for rank in range(3)]
    assert sum(p[1].cardinality for p in interleaved) == whole_filter.cardinality
    assert max(p[1].cardinality for p in interleaved) < 0.6 * whole_filter.cardinality
    np.testing.assert_array_equal(sum(p[2].astype(np.uint64) for p in interleaved).astype(np.uint32), whole_counts)
    np.testing.assert_array_equal(sum(p[3].astype(np.uint64) for p in interleaved).astype(np.uint32), whole_full)
    # and the whole agrees with the oracle fed the SAME packed column through its import path
    oracle_table = O.Table()
    oracle_table.set_layout(*sizes)
    oracle_table.import_column("main", O.NUCLEOTIDE, synthetic.reference, synthetic.build_column(total_rows, 0, len(sizes), 4))
    # the lineage row set, recomputed independently of the product generator
    n_sequences = synthetic.num_sequences
    in_lineage = np.zeros(n_sequences, dtype=bool)
    in_lineage[ancestor] = True
    for e in range(ancestor + 1, n_sequences):
        in_lineage[e] = in_lineage[synthetic.parent(e)]
    lineage_rows = np.flatnonzero(in_lineage[np.arange(total_rows) % n_sequences])
    oracle_table.register_bitmap("lineage", lineage_rows.tolist())  # dense table: sparse id == row number
    np.testing.assert_array_equal(whole_table.filter("(bitmap lineage)").ids(), lineage_rows.astype(np.uint32))
    expression = f"(and {date_filter(0, len(sizes))} (bitmap lineage))"
    oracle_filter = oracle_table.filter(expression)
    assert oracle_filter.cardinality == whole_filter.cardinality
    np.testing.assert_array_equal(whole_counts, oracle_table.mutation_counts("main", oracle_filter))
    np.testing.assert_array_equal(whole_full, oracle_table.mutation_counts("main"))
    rows = whole_table.mutation_rows_from_counts("main", whole_counts, 0.05)
    assert rows == oracle_table.mutation_rows("main", whole_counts, 0.05)
    # size-independent property: per position the symbol counts add up to the covered filtered rows
    assert (whole_counts.sum(axis=0) == whole_filter.cardinality).all()

    # The sharded query as the scheduler runs it (bench.py at N > 1, the ranks played by three tables on one
    # device): every shard enqueues program + filter + counts without synchronising, the counts are summed on
    # the same stream (the all-reduce), one shard runs the output pass over the sums on the device.
    import torch
    stream = torch.cuda.Stream()
    # a prepared program with the counts kernels behind it (what bench.py's device-resident loop launches)
    prepared = whole_table.prepare(expression)
    for _ in range(3):
        with torch.cuda.stream(stream):
            buffer = torch.full((16 * 1500,), -1, dtype=torch.int32, device="cuda")
            prepared.run_counts_async(0, buffer.data_ptr(), stream.cuda_stream)
            stream.synchronize()
            np.testing.assert_array_equal(buffer.cpu().numpy().view(np.uint32).reshape(16, 1500), whole_counts)
    assert prepared.cardinality() == whole_filter.cardinality
    prepared.close()
    for repeat, min_proportion in enumerate((0.05, 0.0, 0.3, 0.05)):
        with torch.cuda.stream(stream):
            buffers = [torch.zeros(16 * 1500, dtype=torch.int32, device="cuda") for _ in interleaved]
            for rank, (table, _, _, _) in enumerate(interleaved):
                first, n, stride = host_api.interleaved_shard(len(sizes), 3, rank)
                text = f"(and {host_api.date_ranges_expression(total_rows, 1095, 200, 800, first, n, stride)} (bitmap lineage))"
                table.mutations_enqueue("main", text, buffers[rank].data_ptr(), stream.cuda_stream)
            summed = buffers[0] + buffers[1] + buffers[2]
            cardinalities = []
            # (every shard collects, so that each one's filter scalars are read and reset)
            for table, _, _, _ in interleaved:
                columns, shard_cardinality = table.mutations_collect("main", min_proportion, summed.data_ptr(), stream.cuda_stream)
                cardinalities.append(shard_cardinality)
                assert host_api.rows_from_columns(columns) == oracle_table.mutations("main", expression, min_proportion)
            assert sum(cardinalities) == whole_filter.cardinality
            assert cardinalities == [p[1].cardinality for p in interleaved]

    # The same through the library's own scheduler (silo_gpu_shard_group_*): no collective, the finalize kernel of
    # every shard stores its rows into the root's gather area, the root's collect kernel sums them and runs the
    # output pass. Three shards of one process on one device and stream; more queries than gather slots, the
    # unfiltered action in between (stored cardinalities only), the summed counts checked too.
    handles = [table.shard_group_create("main", rank, 3) for rank, (table, _, _, _) in enumerate(interleaved)]
    for table, _, _, _ in interleaved:
        table.shard_group_connect(handles)
    root = interleaved[0][0]
    with torch.cuda.stream(stream):
        summed = torch.zeros(16 * 1500, dtype=torch.int32, device="cuda")
        for repeat, (filtered, min_proportion) in enumerate([(True, 0.05), (True, 0.0), (False, 0.05), (True, 0.3), (False, 0.0), (True, 0.05), (True, 0.5)]):
            for rank, (table, _, _, _) in enumerate(interleaved):
                first, n, stride = host_api.interleaved_shard(len(sizes), 3, rank)
                text = f"(and {host_api.date_ranges_expression(total_rows, 1095, 200, 800, first, n, stride)} (bitmap lineage))"
                table.sharded_enqueue("main", text if filtered else None, stream.cuda_stream)
            columns, cardinality = root.sharded_collect("main", min_proportion, stream.cuda_stream, summed.data_ptr())
            assert host_api.rows_from_columns(columns) == oracle_table.mutations("main", expression if filtered else None, min_proportion)
            assert cardinality == (whole_filter.cardinality if filtered else total_rows)
            want = whole_counts if filtered else whole_full
            np.testing.assert_array_equal(summed.cpu().numpy().view(np.uint32).reshape(16, 1500)[:5], want[:5])
        # the root's two halves as one call: its finalize kernel waits for the other shards' rows, adds its own counts
        # and runs the output pass (no store of the root's rows, no collect kernel); twice the same shape -> a replayed graph
        for filtered, min_proportion in [(True, 0.05), (False, 0.05), (True, 0.0), (True, 0.05), (True, 0.05)]:
            texts = []
            for rank, (table, _, _, _) in enumerate(interleaved):
                first, n, stride = host_api.interleaved_shard(len(sizes), 3, rank)
                texts.append(f"(and {host_api.date_ranges_expression(total_rows, 1095, 200, 800, first, n, stride)} (bitmap lineage))")
                if rank != 0:
                    table.sharded_enqueue("main", texts[rank] if filtered else None, stream.cuda_stream)
            columns, cardinality = root.sharded_query("main", texts[0] if filtered else None, min_proportion, summed.data_ptr())
            assert host_api.rows_from_columns(columns) == oracle_table.mutations("main", expression if filtered else None, min_proportion)
            assert cardinality == (whole_filter.cardinality if filtered else total_rows)
            want = whole_counts if filtered else whole_full
            np.testing.assert_array_equal(summed.cpu().numpy().view(np.uint32).reshape(16, 1500)[:5], want[:5])
        # two queries in flight before the root collects (ranks run ahead of the root)
        for _ in range(2):
            for rank, (table, _, _, _) in enumerate(interleaved):
                first, n, stride = host_api.interleaved_shard(len(sizes), 3, rank)
                text = f"(and {host_api.date_ranges_expression(total_rows, 1095, 200, 800, first, n, stride)} (bitmap lineage))"
                table.sharded_enqueue("main", text, stream.cuda_stream)
        for _ in range(2):
            columns, cardinality = root.sharded_collect("main", 0.05, stream.cuda_stream)
            assert host_api.rows_from_columns(columns) == oracle_table.mutations("main", expression, 0.05)
            assert cardinality == whole_filter.cardinality


def test_baseline_sizes_size_independent_properties(ctx):
    """BASELINE.json configs 2 and 3 at full size (10 M rows x 29,903 nt; the oracle would need minutes per
    query there), checked through properties that hold at any size: per position the symbol counts add up
    to |filter|; shards of the filter sum to the whole; k-of-n is monotone in k; a profile filter of distance
    d is contained in distance d + 1; And/Or/Not identities on the cardinalities."""
    from lapis_silo_b200 import host_api
    total_rows = 10_000_000
    length = 29903
    synthetic = host_api.Synthetic(genome_length=length, reference_seed=1, generations=5)
    sizes = host_api.dense_chunk_sizes(total_rows)
    table = host_api.HostTable(ctx, sizes)
    table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_column(total_rows, 0, len(sizes), 16))
    synthetic.release_column()
    ancestor = next(e for e in range(synthetic.num_sequences) if synthetic.generation(e) == 2)
    table.register_bitmap("lineage", synthetic.lineage_bitmap(ancestor, total_rows, 0, len(sizes)))
    date = host_api.date_ranges_expression(total_rows, 1095, 366, 546, 0, len(sizes))

    # config 2: Mutations under date range AND lineage
    flt = table.filter(f"(and {date} (bitmap lineage))")
    counts = table.mutation_counts("main", flt)
    assert flt.cardinality > 0
    assert (counts.sum(axis=0, dtype=np.uint64) == flt.cardinality).all()
    in_date = table.filter(date)
    in_lineage = table.filter("(bitmap lineage)")
    either = table.filter(f"(or {date} (bitmap lineage))")
    assert in_date.cardinality + in_lineage.cardinality == flt.cardinality + either.cardinality  # inclusion-exclusion
    assert table.filter(f"(not {date})").cardinality == total_rows - in_date.cardinality
    outside = table.filter(f"(and (not {date}) (bitmap lineage))")
    np.testing.assert_array_equal(
        counts.astype(np.uint64) + table.mutation_counts("main", outside), table.mutation_counts("main", in_lineage))
    whole = table.mutation_counts("main")
    assert (whole.sum(axis=0, dtype=np.uint64) == total_rows).all()
    rows = table.mutations(["main"], f"(and {date} (bitmap lineage))", 0.05)
    assert rows == table.mutation_rows_from_counts("main", counts, 0.05) and len(rows) > 0
    assert all(r["count"] * 20 >= r["coverage"] - 20 and r["coverage"] == flt.cardinality for r in rows)

    # config 3: nucleotideMutationProfile(distance, querySequence = last evolved sequence) -> count()
    query = synthetic.sequence(synthetic.num_sequences - 1)
    cardinalities = [table.filter(f"(profile main {distance} seq {query})").cardinality for distance in (0, 5, 50, 200)]
    assert cardinalities == sorted(cardinalities) and cardinalities[0] > 0
    exact_rows = total_rows // synthetic.num_sequences  # rows that ARE the query sequence: at least every 133rd
    assert cardinalities[0] >= exact_rows
    # k-of-n over three single-position tests is monotone in k, and 1-of-n is their union
    leaves = "(has-mut main 241) (has-mut main 3037) (has-mut main 14408)"
    k_of_n = [table.filter(f"(n-of {k} 0 {leaves})").cardinality for k in (1, 2, 3)]
    assert k_of_n == sorted(k_of_n, reverse=True)
    assert k_of_n[0] == table.filter(f"(or {leaves})").cardinality
    assert k_of_n[2] == table.filter(f"(and {leaves})").cardinality
    table.close()


def profile_brute_force(synthetic, total_rows, query_index):
    """Rows of the synthetic table (row i = evolved[i % n], no missing symbols) by their Hamming distance to
    evolved[query_index]: an independent restatement of nucleotideMutationProfile(distance, querySequence) on this
    data (mutation_profile.cpp:198-257: a position counts when the row's symbol cannot be the query's)."""
    n = synthetic.num_sequences
    sequences = np.array([np.frombuffer(synthetic.sequence(e).encode(), dtype=np.uint8) for e in range(n)])
    return (sequences != sequences[query_index]).sum(axis=1)


def test_baseline_sizes_equal_the_oracle(ctx):
    """BASELINE.json configs 2 and 3 at FULL size (10 M rows x 29,903 nt, bench.py's table), bit-compared with the
    oracle on the same table (the oracle imports the product generator's S1 column: ~5 s; one Mutations query takes
    it ~0.3 s): filtered row ids, u32 counts of all 16 symbols at every position, thresholded rows incl. proportions.
    Config 3: the profile row sets at distance 0 / 5 / 50 / 200 against the brute-force Hamming distances and, at
    distance 0 (SILO_SLOW_PARITY=1: also 5), against the oracle's Threshold DP (O(n k) whole-bitmap passes: 9 s, 59 s,
    541 s for d = 0, 5, 50 -- tests/test_oracle_profile_cpu.py pins the brute force to the DP at reduced size)."""
    import os
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    total_rows, length = 10_000_000, 29903
    synthetic = host_api.Synthetic(genome_length=length, reference_seed=20200101, generations=5)
    sizes = host_api.dense_chunk_sizes(total_rows)
    column = synthetic.build_column(total_rows, 0, len(sizes), os.cpu_count() or 8)
    table = host_api.HostTable(ctx, sizes)
    table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, column)
    oracle_table = O.Table()
    oracle_table.set_layout(*sizes)
    oracle_table.import_column("main", O.NUCLEOTIDE, synthetic.reference, column)
    synthetic.release_column()
    ancestor = next(e for e in range(synthetic.num_sequences) if synthetic.generation(e) == 2)
    lineage_bytes = synthetic.lineage_bitmap(ancestor, total_rows, 0, len(sizes))
    table.register_bitmap("lineage", lineage_bytes)
    n = synthetic.num_sequences
    in_lineage = np.zeros(n, dtype=bool)
    in_lineage[ancestor] = True
    for e in range(ancestor + 1, n):
        in_lineage[e] = in_lineage[synthetic.parent(e)]
    oracle_table.register_bitmap("lineage", np.flatnonzero(in_lineage[np.arange(total_rows) % n]).astype(np.uint32))
    assert oracle_table.bitmap_bytes("lineage") == lineage_bytes

    # config 2
    expression = f"(and {host_api.date_ranges_expression(total_rows, 1095, 366, 546, 0, len(sizes))} (bitmap lineage))"
    got, want = table.filter(expression), oracle_table.filter(expression)
    assert got.cardinality == want.cardinality == 422574
    np.testing.assert_array_equal(got.ids(), want.ids())
    want_counts = oracle_table.mutation_counts("main", want)
    np.testing.assert_array_equal(table.mutation_counts("main", got), want_counts)
    for min_proportion in (0.05, 0.0, 0.5):
        assert table.mutations(["main"], expression, min_proportion) == oracle_table.mutation_rows("main", want_counts, min_proportion)
    # ... and the unfiltered action (stored cardinalities only, mutations_node.cpp:240-266)
    np.testing.assert_array_equal(table.mutation_counts("main"), oracle_table.mutation_counts("main"))

    # config 3
    query_index = n - 1
    query = synthetic.sequence(query_index)
    distances = profile_brute_force(synthetic, total_rows, query_index)
    row_sequence = np.arange(total_rows) % n
    for distance in (0, 5, 50, 200):
        flt = table.filter(f"(profile main {distance} seq {query})")
        want_ids = np.flatnonzero((distances <= distance)[row_sequence]).astype(np.uint32)
        assert flt.cardinality == len(want_ids) > 0
        np.testing.assert_array_equal(flt.ids(), want_ids)
    for distance in (0, 5) if os.environ.get("SILO_SLOW_PARITY") else (0,):
        text = f"(profile main {distance} seq {query})"
        np.testing.assert_array_equal(table.filter(text).ids(), oracle_table.filter(text).ids())
    table.close()


@pytest.mark.parametrize("seed,alphabet_id", [(301, 0), (302, 1)])
def test_threshold_sweep_kernel(ctx, seed, alphabet_id):
    """The whole-column threshold sweep (thresholdSweepKernel) forced on for small tables: wide n-of / profile
    expressions with adding and subtracting leaves, exact and at-least, ragged chunks, nulls, N runs, every container
    kind; then the same with the sweep off (the interpreter's own container walk)."""
    t = build_random(seed, 1500, 60, (299, 300, 700, 1100), alphabet_id)
    rng = np.random.default_rng(seed)
    t.register_bitmap("lineage", sorted(set(int(v) for v in rng.integers(0, 300, 150))))
    device = mirror(ctx, t, ["lineage"])
    pair = (t, device)
    chars = "-ACGTRYSWKMBDHVN" if alphabet_id == 0 else "-ACDEFGHIKLMNOPQRSTUVWYBJZ*X"
    reference = t.columns[0][2]
    for sweep_min_pieces in (0, 2 ** 63):
        device.set_option("sweep_min_pieces", sweep_min_pieces)
        for k in (1, 2, 7, 25, 59):
            children = " ".join(f"(sym-in c {p} {''.join(rng.choice(list(chars[:-1]), 4, replace=False))})" for p in range(1, 61))
            both(pair, f"(n-of {k} 0 {children})")
            both(pair, f"(n-of {k} 1 {children})")
            both(pair, f"(and (bitmap lineage) (not (n-of {k} 0 {children})))")
        for distance in (0, 1, 4, 20, 59):
            both(pair, f"(profile c {distance} muts)")
            query = list(reference)
            for p in rng.integers(0, 60, 12):
                query[int(p)] = chars[int(rng.integers(0, len(chars)))]
            both(pair, f"(profile c {distance} seq {''.join(query)})")
            both(pair, f"(or (bitmap lineage) (profile c {distance} seq {''.join(query)}))")
    device.close()


def test_threshold_sweep_on_tree_data(ctx):
    """The sweep on synthetic tree data over several chunks (a CTA's range crosses chunk boundaries, several CTAs
    flush into one chunk): arrays, runs and bitsets; profile row sets against the brute-force Hamming distances and
    the oracle's Threshold DP."""
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    total_rows, length = 5 * 65536 + 1234, 2500
    synthetic = host_api.Synthetic(genome_length=length, reference_seed=5, generations=5)
    sizes = host_api.dense_chunk_sizes(total_rows)
    column = synthetic.build_column(total_rows, 0, len(sizes), 8)
    table = host_api.HostTable(ctx, sizes)
    table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, column)
    oracle_table = O.Table()
    oracle_table.set_layout(*sizes)
    oracle_table.import_column("main", O.NUCLEOTIDE, synthetic.reference, column)
    synthetic.release_column()
    table.set_option("sweep_min_pieces", 0)
    n = synthetic.num_sequences
    row_sequence = np.arange(total_rows) % n
    for query_index in (n - 1, n // 2):
        distances = profile_brute_force(synthetic, total_rows, query_index)
        query = synthetic.sequence(query_index)
        for distance in (0, 3, 10, 40):
            flt = table.filter(f"(profile main {distance} seq {query})")
            want = np.flatnonzero((distances <= distance)[row_sequence]).astype(np.uint32)
            np.testing.assert_array_equal(flt.ids(), want)
        for distance in (0, 3):
            text = f"(profile main {distance} seq {query})"
            np.testing.assert_array_equal(table.filter(text).ids(), oracle_table.filter(text).ids())
    table.close()


GENE_LENGTHS = {"E": 76, "M": 223, "N": 420, "ORF1a": 4401, "ORF1b": 2696, "ORF3a": 276, "ORF6": 62, "ORF7a": 122, "ORF7b": 44,
                "ORF8": 122, "ORF9b": 98, "S": 1274}  # testBaseData/exampleDataset/reference_genomes.json


def test_amino_acid_mutations_over_twelve_genes(ctx):
    """BASELINE.json configs[3] at reduced rows: AminoAcidMutations over the twelve gene columns under one filter --
    one device call (silo_gpu_query_mutation_hits_columns: the filter evaluated once, one synchronisation) -- against the
    oracle gene by gene; filtered, unfiltered, several minProportions, and a subset of the genes in another order."""
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    total_rows = 3 * 65536 + 4321
    sizes = host_api.dense_chunk_sizes(total_rows)
    table = host_api.HostTable(ctx, sizes)
    oracle_table = O.Table()
    oracle_table.set_layout(*sizes)
    genes = {}
    for index, (name, length) in enumerate(GENE_LENGTHS.items()):
        gene = host_api.Synthetic(genome_length=length, reference_seed=100 + index, generations=5, gene=True, tree_seed=42 + index, mutation_rate=0.003)
        column = gene.build_column(total_rows, 0, len(sizes), 4)
        table.add_column(name, host_api.AMINO_ACID, gene.reference, column)
        oracle_table.import_column(name, O.AMINO_ACID, gene.reference, column)
        gene.release_column()
        genes[name] = gene
    rng = np.random.default_rng(4)
    picked = np.sort(rng.choice(total_rows, total_rows // 7, replace=False)).astype(np.uint32)
    oracle_table.register_bitmap("lineage", picked)
    table.register_bitmap("lineage", oracle_table.bitmap_bytes("lineage"))
    expression = f"(and {host_api.date_ranges_expression(total_rows, 1095, 200, 800, 0, len(sizes))} (bitmap lineage))"
    names = list(GENE_LENGTHS)
    for text in (expression, None, "(sym-eq S 12 .)"):
        for min_proportion in (0.05, 0.0):
            want = [row for name in names for row in oracle_table.mutations(name, text, min_proportion)]
            assert table.mutations(names, text, min_proportion) == want
            assert len(want) > 0
    subset = ["S", "E", "ORF1a"]
    want = [row for name in subset for row in oracle_table.mutations(name, expression, 0.01)]
    for _ in range(3):  # (the third call replays the captured graph)
        assert table.mutations(subset, expression, 0.01) == want
    table.close()


def test_selection_predicates_on_device(ctx):
    """Predicates over metadata columns inside the filter program (PUSH_COMPARE over device-resident value columns:
    dictionary ids of an unindexed string column, Date32 days) against the oracle's Selection (row-by-row match on a small
    child, makeBitmap + row-by-row on a large one, selection.cpp:94-141): string equality, date ranges on unsorted columns
    with nulls and on a sorted column (RangeSelection), combined with sequence filters, and as the filter of Mutations."""
    from lapis_silo_b200 import host_api
    t = build_random(77, 3000, 40, (999, 1000, 2100))
    rng = np.random.default_rng(77)
    n_rows = t.num_rows
    places = ["basel", "bern", "geneva", "generated"]
    location = [None if rng.random() < 0.05 else places[int(rng.integers(0, 4))] for _ in range(n_rows)]
    unsorted_days = [None if rng.random() < 0.05 else int(rng.integers(18000, 19500)) for _ in range(n_rows)]
    sorted_days = np.sort(rng.integers(18000, 19500, n_rows)).astype(np.int32)
    t.add_string_column("location", location)
    t.add_date_column("date", unsorted_days)
    t.add_date_column("sampling", sorted_days)
    small = sorted(set(int(v) for v in rng.integers(0, 900, 40)))
    t.register_bitmap("small", small)
    device = mirror(ctx, t, ["small"])
    device.add_string_column("location", location)
    device.add_date_column("date", unsorted_days)
    device.add_date_column("sampling", sorted_days)
    pair = (t, device)
    leaves = ["(str-eq location basel)", "(str-eq location generated)", "(str-eq location nowhere)", "(date-between date 18500 19000)",
              "(date-between date * 18200)", "(date-between date 19300 *)", "(date-between sampling 18300 18900)", "(date-between sampling * *)",
              "(bitmap small)", "(sym-eq c 7 A)", "(has-mut c 12)", "(sym-eq c 3 N)"]
    for leaf in leaves:
        both(pair, leaf)
        both(pair, f"(not {leaf})")
    for _ in range(60):
        a, b, c = (leaves[int(i)] for i in rng.integers(0, len(leaves), 3))
        both(pair, f"(and {a} {b})")
        both(pair, f"(and {a} (not {b}) {c})")
        both(pair, f"(or {a} (and {b} {c}))")
        both(pair, f"(n-of 2 0 {a} {b} {c})")
        both(pair, f"(not (and {a} (or {b} (not {c}))))")
    # CountFilterNode as one device call; queries of one shape replay a captured graph
    for position in range(1, 40):
        for text in (f"(and (str-eq location generated) (date-between sampling 18300 18900) (sym-eq c {position} A))",
                     f"(and (str-eq location basel) (or (sym-eq c {position} A) (sym-eq c {position} C) (sym-eq c {position} -)))"):
            assert device.count(text) == t.filter(text).cardinality
    assert device.count(None) == t.num_rows
    expression = "(and (str-eq location generated) (date-between date 18200 19300) (not (sym-eq c 9 N)))"
    assert device.mutations(["c"], expression, 0.02) == t.mutations("c", expression, 0.02)
    assert "$string location IN ['generated']" in device.to_strings(expression)[2] and "$date date >= 18200" in device.to_strings(expression)[2]
    device.close()


def test_concurrent_queries_from_many_host_threads(ctx):
    """The reference answers queries from `parallel_threads` Poco workers (api/api.cpp:39-51). Eight host threads issue
    different fused queries against ONE table at the same time -- Mutations (one and two columns, several filters and
    minProportions), count(), filters with row sets, a MutationProfile --; every answer equals the oracle's. (Calls on one
    table are serialised by the table's mutex: the test pins that the shared per-table staging, graph cache and scratch
    survive arbitrary interleavings, not that the queries overlap.)"""
    import threading
    t = build_random(909, 2500, 50, (700, 701, 1800))
    rng = np.random.default_rng(909)
    t.register_bitmap("lineage", sorted(set(int(v) for v in rng.integers(0, 700, 200))))
    device = mirror(ctx, t, ["lineage"])
    reference = t.columns[0][2]
    jobs = []
    for index in range(8):
        position = 3 + 5 * index
        expression = f"(and (not (sym-eq c {position} N)) (or (bitmap lineage) (has-mut c {position + 1})))"
        if index % 4 == 0:
            want = t.mutations("c", expression, 0.03 * (index + 1))
            jobs.append((lambda e=expression, p=0.03 * (index + 1): device.mutations(["c"], e, p), want))
        elif index % 4 == 1:
            want = t.filter(expression).cardinality
            jobs.append((lambda e=expression: device.count(e), want))
        elif index % 4 == 2:
            want = [int(v) for v in t.filter(expression).ids()]
            jobs.append((lambda e=expression: [int(v) for v in device.filter(e).ids()], want))
        else:
            profile = f"(profile c {index} seq {reference})"
            want = t.mutations("c", profile, 0.0)
            jobs.append((lambda e=profile: device.mutations(["c"], e, 0.0), want))
    failures = []

    def worker(job, want):
        try:
            for _ in range(25):
                if job() != want:
                    failures.append("a concurrent query returned another result than the oracle")
                    return
        except Exception as error:  # noqa: BLE001
            failures.append(repr(error))
    threads = [threading.Thread(target=worker, args=job) for job in jobs]
    for thread in threads:
        thread.start()
    for thread in threads:
        thread.join()
    assert failures == []
    device.close()
