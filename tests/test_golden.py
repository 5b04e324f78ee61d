"""Committed golden vectors (tests/golden/synthetic_small.json, made by tests/golden/make_golden.py):
the oracle must keep reproducing them (CPU), and the CUDA path must match them through the host layer
and the C ABI (GPU)."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

with open(os.path.join(HERE, "golden", "synthetic_small.json")) as handle:
    GOLDEN = json.load(handle)


def digest(array) -> str:
    return hashlib.sha256(np.ascontiguousarray(array).tobytes()).hexdigest()


def check_case(case, ids, counts, rows_005, rows_0):
    assert len(ids) == case["cardinality"], case["expression"]
    assert digest(np.asarray(ids, dtype=np.uint32)) == case["ids_sha256"], case["expression"]
    assert [int(v) for v in ids[:8]] == case["first_ids"]
    assert digest(np.asarray(counts, dtype=np.uint32)) == case["counts_sha256"], case["expression"]
    assert rows_005[:12] == case["rows_min_proportion_0.05"], case["expression"]
    assert len(rows_005) == case["n_rows_min_proportion_0.05"]
    assert len(rows_0) == case["n_rows_min_proportion_0"]


def test_oracle_reproduces_the_golden_vectors():
    import make_golden
    reference, table = make_golden.build()
    assert hashlib.sha256(reference.encode()).hexdigest() == GOLDEN["reference_sha256"]
    assert hashlib.sha256(table.local_reference("main").encode()).hexdigest() == GOLDEN["local_reference_sha256"]
    assert table.num_containers("main") == GOLDEN["num_containers"]
    for case in GOLDEN["cases"]:
        flt = table.filter(case["expression"])
        counts = table.mutation_counts("main", flt)
        check_case(case, flt.ids(), counts, table.mutation_rows("main", counts, 0.05), table.mutation_rows("main", counts, 0.0))


@pytest.mark.gpu
def test_device_path_matches_the_golden_vectors():
    import make_golden
    from lapis_silo_b200 import abi, host_api
    reference, oracle_table = make_golden.build()  # only as the source of the uploaded column (S1 export)
    ctx = abi.Context(0)
    try:
        table = host_api.HostTable(ctx, oracle_table.chunk_sizes)
        export = oracle_table.export_column("main")
        table.add_column("main", host_api.NUCLEOTIDE, reference, export.desc)
        export.close()
        table.register_bitmap("lineage", oracle_table.bitmap_bytes("lineage"))
        for case in GOLDEN["cases"]:
            flt = table.filter(case["expression"])
            counts = table.mutation_counts("main", flt)
            check_case(
                case, flt.ids(), counts, table.mutation_rows_from_counts("main", counts, 0.05),
                table.mutation_rows_from_counts("main", counts, 0.0))
            assert table.mutations(["main"], case["expression"], 0.05)[:12] == case["rows_min_proportion_0.05"]
        table.close()
    finally:
        ctx.close()
