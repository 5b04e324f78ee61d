"""GPU parity, kernel level: the C ABI (libsilo_b200.so) against the CPU oracle on the same inputs.

Filter programs are written by hand here; tests/test_gpu_parity.py drives the same kernels through
the host-side lowering of filter expressions.
"""
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from lapis_silo_b200 import abi
    context = abi.Context(0)
    yield context
    context.close()


def upload(ctx, oracle_table, columns):
    from lapis_silo_b200 import abi
    table = abi.Table(ctx, oracle_table.chunk_sizes)
    for name in columns:
        export = oracle_table.export_column(name)
        table.upload_column(export.desc)
        export.close()
    return table


def random_table(seed, n_rows, length, flushes=(), p_mut=0.05, p_missing=0.05, alphabet=None):
    from oracle import oracle as O
    alphabet = O.NUCLEOTIDE if alphabet is None else alphabet
    chars = O.NUC_SYMBOLS if alphabet == O.NUCLEOTIDE else O.AA_SYMBOLS
    missing = "N" if alphabet == O.NUCLEOTIDE else "X"
    concrete = chars[1:5] if alphabet == O.NUCLEOTIDE else chars[1:23]
    rng = np.random.default_rng(seed)
    reference = "".join(concrete[int(i)] for i in rng.integers(0, len(concrete), length))
    ref_arr = np.frombuffer(reference.encode(), dtype=np.uint8)
    all_chars = np.frombuffer(chars.encode(), dtype=np.uint8)
    table = O.Table()
    table.add_column("c", alphabet, reference)
    # per-position mutation rates spread over orders of magnitude -> array AND bitset containers
    rates = p_mut * np.exp(rng.uniform(-3, 3, length)).clip(0, 0.7 / max(p_mut, 1e-9))
    for row in range(n_rows):
        kind = rng.random()
        if kind < 0.02:
            table.append_row([None])
        else:
            seq = ref_arr.copy()
            hit = rng.random(length) < rates
            seq[hit] = all_chars[rng.integers(0, len(all_chars), int(hit.sum()))]
            seq[rng.random(length) < p_missing] = ord(missing)
            start = int(rng.integers(0, length // 2)) if kind < 0.3 else 0
            end = int(rng.integers(start + 1, length + 1)) if kind < 0.3 else length
            table.append_row([(seq[start:end].tobytes().decode(), start)])
        if row in flushes:
            table.flush_chunk()
    table.finalize()
    return table


def program(*instrs):
    return list(instrs)


def test_library_reports_sm100(ctx):
    from lapis_silo_b200 import abi
    assert b"sm_100a" in abi.lib().silo_gpu_version()


def test_mutation_counts_tiny(ctx):
    from oracle import oracle as O
    t = O.Table()
    t.add_column("segment1", O.NUCLEOTIDE, "ATGCN")
    for seq in ("ATGCN", "ATGCN", "NNNNN", "CATTT", None):
        t.append_row([seq])
    t.finalize()
    g = upload(ctx, t, ["segment1"])
    np.testing.assert_array_equal(g.mutation_counts(0), t.mutation_counts("segment1"))
    for mask in range(32):
        words = np.zeros(1024, dtype=np.uint64)
        words[0] = mask
        selected = [i for i in range(5) if mask >> i & 1]
        t.register_bitmap("sel", selected)
        want = t.mutation_counts("segment1", t.filter("(bitmap sel)"))
        got = g.mutation_counts(0, g.filter_from_words(words))
        np.testing.assert_array_equal(got, want, err_msg=f"mask {mask}")


@pytest.mark.parametrize("seed,alphabet_id", [(11, 0), (12, 1)])
def test_mutation_counts_random_ragged_chunks(ctx, seed, alphabet_id):
    t = random_table(seed, 700, 61, flushes=(99, 130, 131, 500), alphabet=alphabet_id)
    assert t.chunk_sizes == [100, 31, 1, 369, 199]
    g = upload(ctx, t, ["c"])
    np.testing.assert_array_equal(g.mutation_counts(0), t.mutation_counts("c"))
    rng = np.random.default_rng(seed)
    for density in (0.0, 0.01, 0.3, 0.9, 1.0):
        words = np.zeros(5 * 1024, dtype=np.uint64)
        chosen = []
        for chunk, size in enumerate(t.chunk_sizes):
            for row in range(size):
                if rng.random() < density:
                    words[chunk * 1024 + row // 64] |= np.uint64(1) << np.uint64(row % 64)
                    chosen.append((chunk << 16) | row)
        t.register_bitmap("sel", chosen)
        want = t.mutation_counts("c", t.filter("(bitmap sel)"))
        flt = g.filter_from_words(words)
        assert flt.cardinality == len(chosen)
        np.testing.assert_array_equal(g.mutation_counts(0, flt), want, err_msg=f"density {density}")


def test_mutation_counts_all_container_kinds(ctx):
    """Two full chunks + a ragged one; dense random columns give bitsets, the cycled tree gives runs."""
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    reference = "".join("ACGT"[int(i)] for i in rng.integers(0, 4, 300))
    evolved, _ = O.gen_evolved(reference, seed=42, mutation_rate=0.01, generations=4)
    # one extra sequence that differs from the reference nearly everywhere at the first 8 positions
    noisy = []
    for k in range(7):
        s = list(evolved[k % len(evolved)])
        for p in range(8):
            if (k >> (p % 3)) & 1:
                s[p] = "ACGT"[(("ACGT".index(reference[p])) + 1 + k % 3) % 4]
        noisy.append("".join(s))
    t = O.Table()
    t.add_column("main", O.NUCLEOTIDE, reference)
    t.append_cycled(evolved + noisy, 65536 * 2 + 3000)
    special = list(reference)
    special[100] = "ACGT"[("ACGT".index(reference[100]) + 1) % 4]
    t.append_cycled(["".join(special)], 500)  # 500 consecutive rows sharing a diff -> a run container
    t.finalize()
    export = t.export_column("main")
    kinds = {export.desc.contents.containers[i].typecode for i in range(export.desc.contents.n_containers)}
    assert kinds == {1, 2, 3}, kinds
    from lapis_silo_b200 import abi
    g = abi.Table(ctx, t.chunk_sizes)
    g.upload_column(export.desc)
    export.close()
    np.testing.assert_array_equal(g.mutation_counts(0), t.mutation_counts("main"))
    n_chunks = len(t.chunk_sizes)
    for density in (0.001, 0.2, 0.97):
        bits = rng.random(n_chunks * 65536) < density
        for chunk, size in enumerate(t.chunk_sizes):
            bits[chunk * 65536 + size:(chunk + 1) * 65536] = False
        words = np.packbits(bits, bitorder="little").view(np.uint64)
        chosen = np.flatnonzero(bits).astype(np.uint32)
        t.register_bitmap("sel", chosen.tolist())
        want = t.mutation_counts("main", t.filter("(bitmap sel)"))
        got = g.mutation_counts(0, g.filter_from_words(words))
        np.testing.assert_array_equal(got, want, err_msg=f"density {density}")
    # a filter that selects only the middle chunk: the work list must skip the others
    words = np.zeros(n_chunks * 1024, dtype=np.uint64)
    words[1024:2048] = np.uint64(0xFFFFFFFFFFFFFFFF)
    t.register_bitmap("sel", range(65536, 2 * 65536))
    np.testing.assert_array_equal(
        g.mutation_counts(0, g.filter_from_words(words)), t.mutation_counts("main", t.filter("(bitmap sel)")))


def test_filter_from_words_rejects_rows_outside_layout(ctx):
    from lapis_silo_b200 import abi
    g = abi.Table(ctx, [5])
    words = np.zeros(1024, dtype=np.uint64)
    words[0] = 1 << 5
    with pytest.raises(abi.SiloGpuError) as error:
        g.filter_from_words(words)
    assert error.value.status == abi.SILO_E_OUT_OF_LAYOUT


def test_filter_program_leaves_and_boolean_ops(ctx):
    from lapis_silo_b200 import abi as A
    t = random_table(21, 500, 40, flushes=(199,))
    g = upload(ctx, t, ["c"])
    from oracle import oracle as O
    nuc = {c: i for i, c in enumerate(O.NUC_SYMBOLS)}

    def check(instrs, expression, blob=b"", bitmaps=()):
        got = g.filter_eval(instrs, blob, list(bitmaps))
        want = t.filter(expression)
        assert got.cardinality == want.cardinality, expression
        np.testing.assert_array_equal(got.ids(), want.ids(), err_msg=expression)

    check([(A.OP_PUSH_FULL, 0, 0, 0, 0)], "(true)")
    check([(A.OP_PUSH_EMPTY, 0, 0, 0, 0)], "(false)")
    check([(A.OP_PUSH_FULL, 0, 0, 0, 0), (A.OP_NOT, 0, 0, 0, 0)], "(false)")
    local_ref = t.local_reference("c")
    for position in (1, 7, 23, 40):
        stored = [s for s in "-ACGTRYSWKMBDHV" if s != local_ref[position - 1]]
        for symbol in stored[:4]:
            check([(A.OP_PUSH_SYMBOLS, 0, 0, position - 1, 1 << nuc[symbol])], f"(sym-eq c {position} {symbol})")
        check([(A.OP_PUSH_COVERED, 0, 0, position - 1, 0)], f"(covered c {position})")
        check([(A.OP_PUSH_COVERED, 1, 0, position - 1, 0)], f"(not-covered c {position})")
        # reference symbol = covered minus every stored non-missing symbol (symbol_in_set.cpp:179-206)
        others = sum(1 << nuc[s] for s in stored)
        check([(A.OP_PUSH_COVERED, 0, 0, position - 1, 0), (A.OP_PUSH_SYMBOLS, 0, 0, position - 1, others),
               (A.OP_ANDNOT, 0, 0, 0, 0)], f"(sym-eq c {position} {local_ref[position - 1]})")
        # missing symbol = (not covered | nothing stored) minus nulls (symbol_in_set.cpp:149-177)
        check([(A.OP_PUSH_COVERED, 1, 0, position - 1, 0), (A.OP_PUSH_NULLS, 0, 0, 0, 0), (A.OP_ANDNOT, 0, 0, 0, 0)],
              f"(sym-eq c {position} N)")
    # boundary leaves: a foreign roaring bitmap and row ranges
    rng = np.random.default_rng(3)
    chosen = sorted({int(v) for v in rng.integers(0, 200, 80)} | {65536 + int(v) for v in rng.integers(0, 300, 150)})
    t.register_bitmap("lineage", chosen)
    raw = t.bitmap_bytes("lineage")
    check([(A.OP_PUSH_BITMAP, 0, 0, 0, 0)], "(bitmap lineage)", bitmaps=[raw])
    # the same bitmap made device resident once: referenced by id, alone and next to an uploaded one
    t.register_bitmap("other", [3, 9, 65536 + 7])
    first_id = g.register_bitmap(t.bitmap_bytes("other"))
    resident_id = g.register_bitmap(raw)
    assert (first_id, resident_id) == (0, 1)
    check([(A.OP_PUSH_INDEX_BITMAP, 0, 0, resident_id, 0)], "(bitmap lineage)")
    check([(A.OP_PUSH_INDEX_BITMAP, 0, 0, resident_id, 0), (A.OP_PUSH_BITMAP, 0, 0, 0, 0), (A.OP_OR, 0, 0, 0, 0),
           (A.OP_PUSH_INDEX_BITMAP, 0, 0, first_id, 0), (A.OP_ANDNOT, 0, 0, 0, 0)],
          "(op-and ((op-or (bitmap lineage) (bitmap other))) ((bitmap other)))", bitmaps=[t.bitmap_bytes("other")])
    g.unregister_bitmap(first_id)
    with pytest.raises(A.SiloGpuError) as error:
        g.filter_eval([(A.OP_PUSH_INDEX_BITMAP, 0, 0, first_id, 0)])
    assert error.value.status == A.SILO_E_BAD_PROGRAM
    assert g.register_bitmap(raw) == first_id  # freed slots are reused
    ranges = [(10, 150), (65536 + 5, 65536 + 250)]
    blob = b"".join(struct.pack("<II", s, e) for s, e in ranges)
    check([(A.OP_PUSH_RANGES, 0, 0, len(ranges), 0)], "(ranges 10 150 65541 65786)", blob=blob)
    check([(A.OP_PUSH_RANGES, 0, 0, 1, 0)], "(ranges 100 65636)", blob=struct.pack("<II", 100, 65536 + 100))
    check([(A.OP_PUSH_BITMAP, 0, 0, 0, 0), (A.OP_PUSH_RANGES, 0, 0, 2, 0), (A.OP_AND, 0, 0, 0, 0),
           (A.OP_PUSH_SYMBOLS, 0, 0, 6, 0x7FFF & ~(1 << nuc[local_ref[6]])), (A.OP_OR, 0, 0, 0, 0), (A.OP_NOT, 0, 0, 0, 0)],
          "(op-not (op-or (op-and ((bitmap lineage) (ranges 10 150 65541 65786)) ()) (sym-in c 7 " +
          "".join(s for s in "-ACGTRYSWKMBDHV" if s != local_ref[6]) + ")))",
          blob=blob, bitmaps=[raw])


def test_filter_program_keeps_bitmap_ids_outside_layout(ctx):
    """A leaf bitmap with an id outside the row layout: the filter carries it like the reference's bitmaps do
    (include/silo_b200.h, SILO_E_OUT_OF_LAYOUT); only consumers of per-row data refuse the filter."""
    from lapis_silo_b200 import abi as A
    from oracle import oracle as O
    t = O.Table()
    t.set_layout(4)
    t.register_bitmap("bad", [1, 4])
    g = A.Table(ctx, [4])
    flt = g.filter_eval([(A.OP_PUSH_BITMAP, 0, 0, 0, 0)], bitmaps=[t.bitmap_bytes("bad")])
    assert flt.cardinality == 2
    bad_id = g.register_bitmap(t.bitmap_bytes("bad"))
    flt = g.filter_eval([(A.OP_PUSH_INDEX_BITMAP, 0, 0, bad_id, 0), (A.OP_NOT, 0, 0, 0, 0)])
    assert flt.cardinality == 4  # {0, 2, 3} and the id outside the layout, which the complement leaves alone


def test_threshold_programs(ctx):
    """Threshold vectors of filter/operators/threshold.test.cpp via THR_ADD over generic child tiles."""
    from lapis_silo_b200 import abi as A
    from oracle import oracle as O

    def run(layout, pos, neg, k, exact):
        t = O.Table()
        t.set_layout(layout)
        g = A.Table(ctx, [layout])
        names, bitmaps, instrs = [], [], [(A.OP_THR_BEGIN, int(exact), 0, k, 0)]
        for i, ids in enumerate(pos + neg):
            t.register_bitmap(f"b{i}", ids)
            bitmaps.append(t.bitmap_bytes(f"b{i}"))
            instrs += [(A.OP_PUSH_BITMAP, 0, 0, i, 0), (A.OP_THR_ADD, 1 if i >= len(pos) else 0, 0, 0, 0)]
        instrs.append((A.OP_THR_END, 0, 0, 0, 0))
        got = [int(v) for v in g.filter_eval(instrs, b"", bitmaps).ids()]
        lists = lambda sets: " ".join("(ids " + " ".join(map(str, s)) + ")" for s in sets)
        want = [int(v) for v in t.filter(f"(op-threshold {k} {int(exact)} ({lists(pos)}) ({lists(neg)}))").ids()]
        assert got == want, (pos, neg, k, exact)
        return got

    assert run(4, [], [[1, 2, 3], [1, 3]], 1, True) == [2]
    assert run(4, [], [[1, 2, 3], [1, 3]], 1, False) == [0, 2]
    assert run(4, [[1, 2], [1, 3], [1, 2, 3]], [], 2, True) == [2, 3]
    pos, neg = [[1, 2, 3], [1, 3], [1, 2, 3]], [[], [3]]
    assert [run(4, pos, neg, k, True) for k in (1, 2, 3, 4)] == [[], [0], [], [2, 3]]
    assert [run(4, pos, neg, k, False) for k in (1, 2, 3, 4)] == [[0, 1, 2, 3], [0, 1, 2, 3], [1, 2, 3], [1, 2, 3]]
    pos, neg = [[1, 2, 3]], [[], [3], [4], [2, 4]]
    assert [run(5, pos, neg, k, True) for k in (1, 2, 3, 4)] == [[], [4], [], [0, 2, 3]]


def test_threshold_fused_symbol_leaves_and_profile(ctx):
    """k-of-n over SymbolInSet children: generic tiles, fused sparse adds and the one-pass profile."""
    from lapis_silo_b200 import abi as A
    from oracle import oracle as O
    t = random_table(31, 900, 50, flushes=(399,), p_mut=0.08)
    g = upload(ctx, t, ["c"])
    nuc = {c: i for i, c in enumerate(O.NUC_SYMBOLS)}
    local_ref = t.local_reference("c")
    rng = np.random.default_rng(7)
    positions = sorted(int(p) for p in rng.choice(50, 20, replace=False))
    children = []   # (position0, symbols) with symbols never containing N
    for p in positions:
        symbols = "".join(sorted(set(rng.choice(list("-ACGT"), int(rng.integers(1, 4)))), key="-ACGT".index))
        children.append((p, symbols))
    expression_children = " ".join(f"(sym-in c {p + 1} {s})" for p, s in children)
    for k, exact in ((1, False), (3, False), (3, True), (7, False), (19, False)):
        want = t.filter(f"(n-of {k} {int(exact)} {expression_children})")
        # fused leaves: children that include the local reference = covered - stored others
        cover_positions = [p for p, s in children if local_ref[p] in s]
        instrs = [(A.OP_THR_BEGIN, int(exact), 0, k, len(cover_positions))]
        for p, s in children:
            if local_ref[p] in s:
                others = sum(1 << nuc[x] for x in "-ACGTRYSWKMBDHV" if x not in s)
                instrs.append((A.OP_THR_ADD_SYMBOLS, 1, 0, p, others))
            else:
                instrs.append((A.OP_THR_ADD_SYMBOLS, 0, 0, p, sum(1 << nuc[x] for x in s)))
        blob = struct.pack(f"<{len(cover_positions)}I", *cover_positions)
        instrs.append((A.OP_THR_ADD_COVERED, 0, 0, len(cover_positions), 0))
        instrs.append((A.OP_THR_END, 0, 0, 0, 0))
        got = g.filter_eval(instrs, blob)
        np.testing.assert_array_equal(got.ids(), want.ids(), err_msg=f"fused k={k} exact={exact}")
        # the same as ONE profile pass + the covered list
        table = np.zeros((50, 2), dtype=np.uint32)
        for p, s in children:
            if local_ref[p] in s:
                table[p, 1] = sum(1 << nuc[x] for x in "-ACGTRYSWKMBDHV" if x not in s)
            else:
                table[p, 0] = sum(1 << nuc[x] for x in s)
        blob2 = table.tobytes() + blob
        instrs2 = [(A.OP_THR_BEGIN, int(exact), 0, k, len(cover_positions)),
                   (A.OP_THR_PROFILE, 0, 0, 0, 0),
                   (A.OP_THR_ADD_COVERED, 0, 0, len(cover_positions), table.nbytes),
                   (A.OP_THR_END, 0, 0, 0, 0)]
        got2 = g.filter_eval(instrs2, blob2)
        np.testing.assert_array_equal(got2.ids(), want.ids(), err_msg=f"profile k={k} exact={exact}")
