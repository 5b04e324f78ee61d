"""The front half of the query compiler (parse -> rewrite -> compile -> lower to a filter program) on a
HOST-ONLY table (silo_host_table_create with no device context): lowering of the reference's expression /
operator rules checked where there is no GPU, and the one thing such a table must never do -- answer a query."""
import ctypes as C
import time

import numpy as np
import pytest

from lapis_silo_b200 import abi, host_api as H

NUC = "-ACGTRYSWKMBDHVN"
REFERENCE = "ACGTACGTAC"
LOCAL = "ACGTACGTTC"  # position 8 (0-based): the local reference differs from the global one


def host_only_table(reference=REFERENCE, local=LOCAL, null_rows=False):
    table = H.HostTable(None, [65536, 100])
    desc = abi.ColumnDesc()
    desc.struct_size = C.sizeof(abi.ColumnDesc)
    desc.n_symbols, desc.genome_length, desc.missing_symbol = 16, len(reference), 15
    local_ids = (C.c_uint8 * len(local))(*[NUC.index(c) for c in local])
    desc.local_reference = C.cast(local_ids, C.POINTER(C.c_uint8))
    null_ids = (C.c_uint32 * 1)(3)
    if null_rows:
        desc.n_null_rows = 1
        desc.null_row_ids = C.cast(null_ids, C.POINTER(C.c_uint32))
    table.add_column("main", H.NUCLEOTIDE, reference, C.pointer(desc))
    return table


def instrs(table, expression):
    lines = table.explain(expression).strip().split("\n")
    out = []
    for line in lines[1:]:
        fields = dict(part.split("=") for part in line.split())
        out.append((int(fields["op"]), int(fields["flags"]), int(fields["column"]), int(fields["a"]), int(fields["b"])))
    return out


def mask(symbols):
    return sum(1 << NUC.index(s) for s in symbols)


def test_symbol_in_set_cases():  # symbol_in_set.cpp:231-264
    table = host_only_table()
    # neither the local reference symbol nor N: one scan over the vertical index
    assert instrs(table, "(sym-eq main 1 C)") == [(3, 0, 0, 0, mask("C"))]
    # the local reference symbol: covered rows minus the rows holding another non-missing symbol
    assert instrs(table, "(sym-eq main 1 A)") == [(4, 0, 0, 0, 0), (3, 0, 0, 0, mask(NUC) & ~mask("AN")), (17, 0, 0, 0, 0)]
    # '.' is the GLOBAL reference symbol; at position 9 the local reference is T, so C is a stored symbol
    assert instrs(table, "(sym-eq main 9 .)") == [(3, 0, 0, 8, mask("A"))]
    assert instrs(table, "(sym-eq main 9 T)")[0] == (4, 0, 0, 8, 0)
    # N: not covered, or stored as N
    assert instrs(table, "(sym-eq main 2 N)") == [(4, 1, 0, 1, 0), (3, 0, 0, 1, mask("N")), (18, 0, 0, 0, 0)]
    with_nulls = host_only_table(null_rows=True)
    assert instrs(with_nulls, "(sym-eq main 2 N)")[-2:] == [(5, 0, 0, 0, 0), (17, 0, 0, 0, 0)]
    for t in (table, with_nulls):
        t.close()


def test_errors_are_the_reference_messages():
    table = host_only_table()
    with pytest.raises(H.HostError, match="position is out of bounds 11 > 10"):
        table.explain("(sym-eq main 11 A)")
    with pytest.raises(H.HostError, match="Database does not contain the Sequence with name: 'other'"):
        table.explain("(has-mut other 1)")
    with pytest.raises(H.HostError, match="querySequence length 3 does not match the reference sequence length 10"):
        table.explain("(profile main 0 seq ACG)")
    with pytest.raises(H.HostError, match="Invalid Nucleotide symbol 'Z' in querySequence"):
        table.explain("(profile main 0 seq ACGTACGTAZ)")
    with pytest.raises(H.HostError, match="mutation position 11 is out of bounds"):
        table.explain("(profile main 0 muts 11 A)")
    # number lists of the harness notation are read straight from the text
    for expression, message in (("(ranges 1 2 3)", "ranges needs START END pairs"), ("(ranges 1 x)", "expected a non-negative integer, got 'x'"),
                                ("(ids 1 -2)", "expected a non-negative integer, got '-2'"), ("(ids 1 (2))", "expected an atom"),
                                ("(ranges 1 2", "missing '\\)'"), ("(ids 123456789012345678901)", "expected a non-negative integer")):
        with pytest.raises(H.HostError, match=message):
            table.explain(expression)
    assert instrs(table, "(ranges)") == [(7, 0, 0, 0, 0)] and instrs(table, "(ids)")[0][0] == 6  # empty lists are legal
    assert instrs(table, "(ranges 3 9  70000 70010 )")[0][:4] == (7, 0, 0, 2)  # PUSH_RANGES, two ranges
    rng = np.random.default_rng(3)
    for ids in ([], [0], [5, 65535, 65536, 65600], sorted({int(v) for v in rng.integers(0, 65636, 6000)})):
        text = "(ids " + "  ".join(map(str, ids)) + " )"
        runs = H.roaring_runs(table.program_bitmap(text))
        assert [v for first, end in runs for v in range(first, end)] == ids
    table.close()


def test_host_only_table_never_answers_a_query():
    table = host_only_table()
    for call in (lambda: table.filter("(true)"), lambda: table.mutations(["main"], None, 0.05), lambda: table.mutation_counts("main"),
                 lambda: table.register_bitmap("lineage", b"\x3a\x30\0\0\0\0\0\0", True)):
        with pytest.raises(H.HostError, match=r"DeviceError\[-2\].*no CPU fallback"):
            call()
    table.close()


def test_mutation_profile_lowers_to_one_streaming_pass():
    """mutation_profile.cpp:222-247 -> Not(NOf(distance + 1 of the per-position 'definitely different' sets)); with a
    genome-sized number of leaves the counter program holds ONE profile table instead of one leaf per position, and
    the lowered program is the same however the query was phrased"""
    rng = np.random.default_rng(5)
    length = 29903
    reference = "".join("ACGT"[i] for i in rng.integers(0, 4, length))
    table = host_only_table(reference, reference)
    query = list(reference)
    mutations = {int(p): "ACGT"[("ACGT".index(reference[p]) + 1) % 4] for p in rng.integers(0, length, 12)}
    for position, symbol in mutations.items():
        query[position] = symbol
    by_sequence = table.lower_timed(f"(profile main 4 seq {''.join(query)})")
    by_mutations = table.lower_timed("(profile main 4 muts " + " ".join(f"{p + 1} {s}" for p, s in mutations.items()) + ")")
    assert by_sequence["digest"] == by_mutations["digest"]
    assert by_sequence["n_instrs"] == by_mutations["n_instrs"] and by_sequence["blob_bytes"] == by_mutations["blob_bytes"]
    program = instrs(table, "(profile main 4 muts " + " ".join(f"{p + 1} {s}" for p, s in mutations.items()) + ")")
    opcodes = [op for op, *_ in program]
    assert opcodes[0] == 32 and opcodes[-2:] == [37, 19] and opcodes.count(36) == 1  # THR_BEGIN ... THR_PROFILE ... THR_END, NOT
    assert program[0][3] == 5 and program[0][4] == len(mutations)  # k = distance + 1, bias = the 'covered minus' leaves
    assert by_sequence["blob_bytes"] >= 2 * 4 * length
    # N positions drop out of the profile, an ambiguity code narrows its position's set
    query[100], query[200] = "N", "R"
    narrowed = table.lower_timed(f"(profile main 4 seq {''.join(query)})")
    assert narrowed["digest"] != by_sequence["digest"]
    # host cost: one check + one small object per position (the messages of passing checks are never built)
    best = min(sum(table.lower_timed(f"(profile main 4 seq {''.join(query)})")[k] for k in ("parse_us", "rewrite_us", "compile_us", "lower_us"))
               for _ in range(5))
    assert best < 50_000, f"lowering a genome-wide MutationProfile took {best:.0f} us"
    table.close()


def test_to_string_formats_are_the_reference_ones():
    """the reference's own toString vectors: nof.test.cpp:13-48, or.test.cpp:299-321 and :159-178 (merged SymbolInSet);
    formats of and.cpp:32, symbol_in_set.cpp:37-51, mutation_profile.cpp:39-55, threshold.cpp:45-58,
    intersection.cpp:45-53, union.cpp:23-28, selection.cpp:75-88, common/string_utils.h:35-57 (joinWithLimit)"""
    table = host_only_table()
    parsed = lambda expression: table.to_strings(expression)[0]
    assert parsed("(n-of 2 0 (true) (true) (true))") == "[2-of:true, true, true]"
    assert parsed("(n-of 1 1 (true) (true))") == "[exactly-1-of:true, true]"
    assert parsed("(n-of 1 0 (true))") == "[1-of:true]"
    assert parsed("(n-of 0 0)") == "[0-of:]"
    assert table.to_strings("(or (true) (true))")[:2] == ("Or(true | true)", "true")
    assert table.to_strings("(or (or (true) (true)) (true))")[:2] == ("Or(Or(true | true) | true)", "true")
    # Or::rewriteSymbolInSetExpressions: one set per (sequence, position)
    merged = table.to_strings("(or (sym-eq main 3 A) (sym-eq main 3 G))")
    assert merged[0] == "Or(main:3A | main:3G)" and merged[1] == "(main:symbol at position 3 in {A, G})"
    separate = table.to_strings("(or (sym-eq main 3 A) (sym-eq main 4 G))")[1]
    assert separate == "Or((main:symbol at position 3 in {A}) | (main:symbol at position 4 in {G}))"
    assert table.to_strings("(and (sym-eq main 1 C) (not (has-mut main 2)))")[0] == "And(main:1C & !(main:1))"  # has_mutation.cpp:24-26 prints the 0-based index
    assert table.to_strings("(maybe (sym-eq main 1 R))")[:2] == ("Maybe (main:1R)", "(main:symbol at position 1 in {R, D, V, N})")
    assert table.to_strings("(exact (has-mut main 2))")[0] == "Exact (main:1)"
    assert parsed("(profile main 2 muts 1 T 3 N)") == "MutationProfile(main:distance=2,mutations(count=2))"
    assert parsed("(profile main 0 seq ACGTACGTAC)") == "MutationProfile(main:distance=0,querySequence=ACGTACGTAC...)"
    # joinWithLimit: ten items, then the count of the rest
    eleven = " ".join(["(true)"] * 13)
    assert parsed(f"(n-of 3 0 {eleven})") == "[3-of:" + ", ".join(["true"] * 10) + ", ... (3 more)]"
    assert parsed(f"(and {eleven})") == "And(" + " & ".join(["true"] * 10) + " & ... (3 more))"
    # operator trees
    compiled = lambda expression: table.to_strings(expression)[2]
    assert compiled("(true)") == "Full" and compiled("(false)") == "Empty" and compiled("(not (true))") == "Empty"
    assert compiled("(sym-eq main 1 A)") == "Intersection(non_negated: (Select[IsInCoveredRegion(0)]()) negated: (IndexScan(column 0, position 1, symbols 0x7ffd)) )"  # every symbol but A and N
    assert compiled("(sym-eq main 2 N)") == "(Select[!IsInCoveredRegion(1)]() | IndexScan(column 0, position 2, symbols 0x8000))"
    assert compiled("(not (sym-eq main 2 G))") == "!IndexScan(column 0, position 2, symbols 0x8)"
    assert compiled("(n-of 2 0 (sym-eq main 1 C) (sym-eq main 2 A) (not (sym-eq main 3 T)))") == \
        "Threshold(>=2-of non_negated: (IndexScan(column 0, position 1, symbols 0x4), IndexScan(column 0, position 2, symbols 0x2)) " \
        "negated: (IndexScan(column 0, position 3, symbols 0x10)) )"
    table.close()


def test_lowered_programs_equal_the_gpu_verified_ones():
    """tests/golden/lowering_digests.json: digests of the lowered programs of 30 expressions (with and without null rows)
    from the build whose GPU parity suite was green; a change of the host mirror must leave the device's input alone"""
    import json
    import os
    import sys
    golden_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, golden_dir)
    import make_lowering_digests
    with open(os.path.join(golden_dir, "lowering_digests.json")) as handle:
        want = json.load(handle)
    got = make_lowering_digests.lowered()
    assert len(got) == len(want) == 60
    for have, expected in zip(got, want):
        assert have == expected, expected["expression"]
