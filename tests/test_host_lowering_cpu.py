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
