"""The reference pins Mutations numbers only through e2e goldens whose dataset is absent here
(SURVEY.md §8c), so the oracle's Mutations output is additionally checked against an independent
brute-force recount of the raw strings (numpy, no shared code with oracle/src)."""
import numpy as np
import pytest

from oracle import oracle as O


def random_rows(rng, reference, n_rows, alphabet_chars, missing_char):
    length = len(reference)
    rows = []
    for _ in range(n_rows):
        kind = rng.random()
        if kind < 0.05:
            rows.append(None)
            continue
        start = int(rng.integers(0, length // 2)) if rng.random() < 0.5 else 0
        end = int(rng.integers(start + 1, length + 1)) if rng.random() < 0.5 else length
        seq = list(reference[start:end])
        for i in range(len(seq)):
            r = rng.random()
            if r < 0.08:
                seq[i] = alphabet_chars[int(rng.integers(0, len(alphabet_chars)))]
            elif r < 0.14:
                seq[i] = missing_char
        if rng.random() < 0.1:
            seq = [missing_char] * len(seq)
        rows.append(("".join(seq), start))
    return rows


def brute_force_counts(rows, selected, alphabet_chars, missing_char, length):
    counts = np.zeros((len(alphabet_chars), length), dtype=np.uint32)
    index = {c: i for i, c in enumerate(alphabet_chars)}
    for row_id in selected:
        row = rows[row_id]
        if row is None:
            continue
        seq, offset = row
        for i, ch in enumerate(seq):
            if ch != missing_char:
                counts[index[ch], offset + i] += 1
    return counts


@pytest.mark.parametrize("alphabet,chars,missing,seed", [
    (O.NUCLEOTIDE, O.NUC_SYMBOLS, "N", 1), (O.NUCLEOTIDE, O.NUC_SYMBOLS, "N", 2),
    (O.AMINO_ACID, O.AA_SYMBOLS, "X", 3),
])
def test_mutation_counts_match_brute_force(alphabet, chars, missing, seed):
    rng = np.random.default_rng(seed)
    concrete = chars[1:5] if alphabet == O.NUCLEOTIDE else chars[1:23]
    reference = "".join(concrete[int(i)] for i in rng.integers(0, len(concrete), 37))
    rows = random_rows(rng, reference, 400, chars, missing)
    table = O.Table()
    table.add_column("c", alphabet, reference)
    global_ids = []
    chunk, row_in_chunk = 0, 0
    for i, row in enumerate(rows):
        table.append_row([row])
        global_ids.append((chunk << 16) | row_in_chunk)
        row_in_chunk += 1
        if i in (99, 130, 131):  # ragged chunks, including a single-row chunk
            table.flush_chunk()
            chunk, row_in_chunk = chunk + 1, 0
    table.finalize()
    assert table.chunk_sizes == [100, 31, 1, 268]

    everything = table.mutation_counts("c")
    expected = brute_force_counts(rows, range(len(rows)), chars, missing, len(reference))
    np.testing.assert_array_equal(everything, expected)

    picked = sorted(int(v) for v in rng.choice(len(rows), 150, replace=False))
    table.register_bitmap("picked", [global_ids[i] for i in picked])
    flt = table.filter("(bitmap picked)")
    assert flt.cardinality == 150
    np.testing.assert_array_equal(
        table.mutation_counts("c", flt), brute_force_counts(rows, picked, chars, missing, len(reference)))

    # empty filter -> all zero (mutations_node.cpp:282: filter_cardinality > 0 guard)
    np.testing.assert_array_equal(table.mutation_counts("c", table.filter("(false)")), np.zeros_like(expected))

    # every filter form must agree with row-wise evaluation of the same predicate on the strings
    for position in (1, 5, 20, 37):
        for symbol in (concrete[0], concrete[2], missing, "-"):
            want = []
            for i, row in enumerate(rows):
                if row is None:
                    continue
                seq, offset = row
                inside = offset <= position - 1 < offset + len(seq)
                actual = seq[position - 1 - offset] if inside else missing
                if actual == symbol:
                    want.append(global_ids[i])
            got = [int(v) for v in table.filter(f"(sym-eq c {position} {symbol})").ids()]
            assert got == want, (position, symbol)


def test_thresholding_matches_reference_formula():
    # mutations_node.cpp:315-326: threshold_count = ceil(total * p) - 1 computed in double
    table = O.Table()
    table.add_column("c", O.NUCLEOTIDE, "A")
    for ch in "C" * 5 + "A" * 95:
        table.append_row([ch])
    table.finalize()
    assert [r["count"] for r in table.mutations("c", None, 0.05)] == [5]
    assert table.mutations("c", None, 0.0500001) == []
    rows = table.mutations("c", None, 0.0)
    assert [(r["mutationFrom"], r["mutationTo"], r["position"], r["count"], r["coverage"], r["proportion"]) for r in rows] == [
        ("A", "C", 1, 5, 100, 0.05)]
